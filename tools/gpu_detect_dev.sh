#!/bin/bash
# developer call: detect tests + timing of the detect path (short-list route and, for comparison, the exhaustive route)
o=gpurun_out; tag=${1:-dev}
timeout 900 python -m pytest --timeout 300 tests/test_gpu_detect_shortlist.py tests/test_gpu_detect.py tests/test_gpu_levels.py tests/test_gpu_dropin.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py tests/test_gpu_ctx.py -m gpu -x -q > $o/${tag}_pytest_detect.log 2>&1; echo "pytest rc=$?"; tail -15 $o/${tag}_pytest_detect.log
timeout 60 python tools/quick_bench_detect.py 1 8 32 128 256 2>&1 | grep -v "create priors" | tee -a $o/${tag}_quick_detect.log
echo "exhaustive route:" | tee -a $o/${tag}_quick_detect.log
SSDHEAD_DETECT_SHORTLIST=0 timeout 300 python tools/quick_bench_detect.py 1 256 2>&1 | grep -v "create priors" | tee -a $o/${tag}_quick_detect.log
