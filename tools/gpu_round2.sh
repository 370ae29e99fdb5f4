#!/bin/bash
# Two-GPU call: the sharded parity tests (skipped on a one-GPU box) and the N=2 bench lines (peer and NCCL route).
tag=${1:-r2}
o=gpurun_out
mkdir -p $o
python -m pytest tests/test_gpu_sharded.py tests/test_gpu_ctx.py -m gpu -q -rs > $o/${tag}_pytest_sharded_2gpu.log 2>&1; echo "pytest rc=$?" | tee -a $o/${tag}_pytest_sharded_2gpu.log
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 "$@"; }
run --steps 1000 --warmup 20 --no-others > $o/${tag}_bench_2gpu_peer.json 2> $o/${tag}_bench_2gpu.err; echo "peer rc=$?"
run --steps 1000 --warmup 20 --no-others --collective nccl > $o/${tag}_bench_2gpu_nccl.json 2>> $o/${tag}_bench_2gpu.err; echo "nccl rc=$?"
run --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_2gpu_reference.json 2>> $o/${tag}_bench_2gpu.err; echo "ref rc=$?"
tail -4 $o/${tag}_pytest_sharded_2gpu.log; cut -c1-600 $o/${tag}_bench_2gpu_peer.json
