"""Summarise an ncu --set full report: per kernel the headline rows of the details page plus DRAM bytes, duration and
warp instructions from the raw page.   usage: ncu_summary.py <report.ncu-rep> [kernel regex]"""
import csv, subprocess, sys

rep = sys.argv[1]
kre = ["-k", "regex:" + sys.argv[2]] if len(sys.argv) > 2 else []
KEEP = ("Memory Throughput", "DRAM Throughput", "Duration", "L2 Cache Throughput", "Compute (SM) Throughput",
        "Executed Ipc Active", "Issue Slots Busy", "Block Size", "Grid Size", "Registers Per Thread",
        "Dynamic Shared Memory Per Block", "Waves Per SM", "Theoretical Occupancy", "Achieved Occupancy")
det = subprocess.run(["ncu", "-i", rep, "--page", "details"] + kre, capture_output=True, text=True).stdout
seen = set()
for line in det.splitlines():
    st = line.strip()
    if "Context" in line and "Device" in line and "(" in line:
        name = st[:150]
        first = name not in seen
        seen.add(name)
        print("  " + name if first else None) if first else None
        show = first
        continue
    if any(st.startswith(k) for k in KEEP) and "show" in dir() and show:
        print("    " + " ".join(st.split()))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + kre, capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
units = rows[1]
col = {h: i for i, h in enumerate(hdr)}
print("\n# raw: per launch")
done = set()
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    k = r[col["Kernel Name"]][:60]
    if k in done:
        continue
    done.add(k)
    def g(name):
        return f"{r[col[name]]} {units[col[name]]}" if name in col else "?"
    print(f"{k}  dram_read {g('dram__bytes_read.sum')}  dram_write {g('dram__bytes_write.sum')}  "
          f"time {g('gpu__time_duration.sum')}  warp_inst {g('smsp__inst_executed.sum')}")
