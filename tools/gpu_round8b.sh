#!/bin/bash
# N-GPU call with the final library: train (peer route, sharded self-check), stress and detect bench lines (no reference arm)
N=${1:-8}; tag=${2:-r2}
o=gpurun_out; mkdir -p $o
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --steps 1000 --warmup 20 --no-others > $o/${tag}_bench_${N}gpu_train.json 2> $o/${tag}_bench_${N}gpu.err; echo "train rc=$?"
run --steps 200 --warmup 10 --no-others --workload stress > $o/${tag}_bench_${N}gpu_stress.json 2>> $o/${tag}_bench_${N}gpu.err; echo "stress rc=$?"
run --steps 500 --warmup 10 --no-others --workload detect --batch 256 > $o/${tag}_bench_${N}gpu_detect256.json 2>> $o/${tag}_bench_${N}gpu.err; echo "detect rc=$?"
for f in train stress detect256; do python - <<PY
import json
for l in open("$o/${tag}_bench_${N}gpu_$f.json"):
    if l.startswith("{"):
        d=json.loads(l); e=d.get("e2e",{})
        print("$f", d.get("n_gpus"), "ms/step", d.get("ms_per_step"), "value", d.get("value"), "e2e", e.get("value"), "sharded_check", (d.get("sharded_check") or {}).get("ok"), "step_roofline", (d.get("step_roofline") or {}).get("frac"))
PY
done
