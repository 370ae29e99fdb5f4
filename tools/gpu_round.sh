#!/bin/bash
# One GPU call: parity tests, the bench lines of every workload, the reference arm, and the ncu launch list of the
# short bench command (run only after that command exited 0 without ncu). Outputs under gpurun_out/<tag>_*.
tag=${1:-r2}
o=gpurun_out
mkdir -p $o
python -m pytest tests -m gpu -x -q > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $o/${tag}_pytest_gpu.log
python bench.py > $o/${tag}_bench_default.json 2> $o/${tag}_bench_default.err; echo "bench rc=$?"
python bench.py --workload detect --batch 256 --no-others > $o/${tag}_bench_detect256.json 2>> $o/${tag}_bench_default.err; echo "detect256 rc=$?"
python bench.py --workload detect --batch 64 --no-others > $o/${tag}_bench_detect64.json 2>> $o/${tag}_bench_default.err; echo "detect64 rc=$?"
python bench.py --workload stress --no-others > $o/${tag}_bench_stress.json 2>> $o/${tag}_bench_default.err; echo "stress rc=$?"
python bench.py --steps 20 --warmup 3 --no-others > $o/${tag}_bench_short.json 2>> $o/${tag}_bench_default.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/${tag}_bench_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-others > $o/${tag}_ncu_bench.log 2>&1; echo "ncu rc=$?"
tail -3 $o/${tag}_pytest_gpu.log; cat $o/${tag}_bench_default.json | cut -c1-1500
