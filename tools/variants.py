"""Developer workflow for kernel experiments: build several compile-time variants of libssdhead.so here (no GPU needed),
then time all of them - and run the GPU tests on the fastest - in ONE call on the GPU box.

    python tools/variants.py build split1="-DSSDHEAD_FOO=1" split2="-DSSDHEAD_FOO=2"     # -> variants/{base,split1,split2}.so
    gpurun -- 'python tools/variants.py run --test'                                      # on the B200

`run` copies each variant over the in-tree library, runs tools/quick_bench_step.py (whole training step, launches back
to back, exact output checksums), prints one line per variant, optionally runs `pytest -m gpu` with the fastest
non-base variant, and always restores the base library.  variants/ is scratch: *.so files are git-ignored.
"""
import json
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "variants")


def build(specs):
    from objectdetection_ssd_b200 import build as B
    os.makedirs(VDIR, exist_ok=True)
    jobs = [("base", [])] + [(name, flags.split()) for name, flags in specs]

    def one(job):
        name, flags = job
        out = os.path.join(VDIR, name + ".so")
        cmd = [B._nvcc()] + B.NVCC_FLAGS + flags + ["-Xptxas", "-v", "-o", out] + B.sources()
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"{name}: nvcc failed\n{res.stderr[-2000:]}")
        spills = sum(int(ln.split("bytes spill stores")[0].split(",")[-1]) for ln in res.stderr.splitlines() if "bytes spill stores" in ln)
        return name, out, spills

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        for name, out, spills in ex.map(one, jobs):
            print(f"built {out}  (spill stores over all kernels: {spills} B)")


def run(test, batches):
    from objectdetection_ssd_b200 import build as B
    names = sorted(f[:-3] for f in os.listdir(VDIR) if f.endswith(".so"))
    if "base" not in names:
        raise SystemExit("variants/base.so missing: run `variants.py build ...` first")
    order = ["base"] + [n for n in names if n != "base"] + ["base"]          # base twice: shows the run-to-run noise
    results = {}
    try:
        for n in order:
            shutil.copyfile(os.path.join(VDIR, n + ".so"), B.LIB_PATH)
            res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quick_bench_step.py")] + [str(b) for b in batches],
                                 capture_output=True, text=True, timeout=300)
            lines = [json.loads(ln) for ln in res.stdout.splitlines() if ln.startswith("{")]
            if not lines:
                print(f"{n}: FAILED\n{res.stderr[-1500:]}")
                continue
            results.setdefault(n, []).append(lines)
            print(n, " ".join(f"B={d['B']}: {d['step_us_med']} us" for d in lines), "checksum", lines[0]["checksum"])
        if "base" not in results:
            raise SystemExit("the base library did not run (no GPU?)")
        ref = results["base"][0][0]
        cands = {n: r[0][0]["step_us_med"] for n, r in results.items() if n != "base"}
        if cands:
            best = min(cands, key=cands.get)
            same = results[best][0][0]["checksum"] == ref["checksum"]
            print(f"fastest variant: {best} {cands[best]} us (base {ref['step_us_med']} us); outputs bit-identical to base: {same}")
            if test:
                shutil.copyfile(os.path.join(VDIR, best + ".so"), B.LIB_PATH)
                res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-m", "gpu", "-x", "-q"],
                                     capture_output=True, text=True, cwd=ROOT, timeout=900)
                print(f"pytest -m gpu with {best}:", res.stdout.strip().splitlines()[-1] if res.stdout.strip() else res.stderr[-500:])
    finally:
        shutil.copyfile(os.path.join(VDIR, "base.so"), B.LIB_PATH)


if __name__ == "__main__":
    if len(sys.argv) >= 2 and sys.argv[1] == "build":
        build([a.split("=", 1) for a in sys.argv[2:]])
    elif len(sys.argv) >= 2 and sys.argv[1] == "run":
        bs = [int(a) for a in sys.argv[2:] if a.isdigit()] or [256, 32]
        run("--test" in sys.argv, bs)
    else:
        raise SystemExit(__doc__)
