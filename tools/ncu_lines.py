"""Join an ncu source-page CSV (SASS view) with nvdisasm line info: warp-stall samples per CUDA source line.

usage: ncu_lines.py <report.ncu-rep> <cubin> <kernel name substring> [min_samples] [kernel regex]
(NCU_LAUNCH_SKIP=<n> selects the n-th launch of the report instead of a kernel-name regex: template instantiations
share their base name)
"""
import csv, re, subprocess, sys, collections

rep, cubin, kname = sys.argv[1:4]
kregex = sys.argv[5] if len(sys.argv) > 5 else kname
min_s = int(sys.argv[4]) if len(sys.argv) > 4 else 30
import os
sel = ["--launch-skip", os.environ["NCU_LAUNCH_SKIP"], "--launch-count", "1"] if os.environ.get("NCU_LAUNCH_SKIP") else ["-k", "regex:" + kregex]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr): break
    data.append(r)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
cands = [i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l]
best = None
for start in cands:
    lines = []
    cur = ("?", 0)
    for l in dis[start + 1:]:
        if l.startswith(".text.") or l.startswith("\t.section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            lines.append(cur)
    if len(lines) == len(data):
        best = lines
        print("matched", dis[start])
        break
if best is None:
    sys.exit(f"no function with {len(data)} instructions among {[dis[c] for c in cands]}")
lines = best
agg = collections.Counter(); st_agg = collections.defaultdict(collections.Counter); instr = collections.Counter()
tot = 0
for i, r in enumerate(data):
    n = int(r[idx["# Samples"]] or 0)
    key = lines[i]
    agg[key] += n; tot += n
    instr[key] += int(r[idx["Instructions Executed"]] or 0)
    for s in stalls:
        v = int(r[idx[s]] or 0)
        if v: st_agg[key][s[6:]] += v
print("total samples", tot, "total warp instructions", sum(instr.values()))
src = {}
for key, n in sorted(agg.items(), key=lambda x: (x[0][0], x[0][1])):
    if n < min_s: continue
    f, ln = key
    if f not in src:
        try: src[f] = open(f"/root/repo/objectdetection_ssd_b200/csrc/{f}").read().splitlines()
        except OSError: src[f] = []
    text = src[f][ln - 1].strip()[:80] if 0 < ln <= len(src[f]) else ""
    top = ", ".join(f"{k}:{v}" for k, v in st_agg[key].most_common(3))
    print(f"{n:6d} {100*n/tot:5.1f}% inst={instr[key]:9d} {f}:{ln:<4d} {text}   [{top}]")
