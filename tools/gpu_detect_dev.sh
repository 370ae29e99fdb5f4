#!/bin/bash
# developer call: detect tests + timing of the detect path (automatic route, then each route forced)
o=gpurun_out; tag=${1:-dev}
timeout 600 python -m pytest --timeout 300 tests/test_gpu_detect_shortlist.py tests/test_gpu_detect.py tests/test_gpu_levels.py tests/test_gpu_dropin.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py tests/test_gpu_ctx.py -m gpu -x -q 2>&1 | tail -3 | tee $o/${tag}_pytest_detect.log
for mode in 1 0; do
  echo "SSDHEAD_DETECT_SHORTLIST=$mode" | tee -a $o/${tag}_quick_detect.log
  SSDHEAD_DETECT_SHORTLIST=$mode python tools/quick_bench_detect.py 1 8 32 128 256 2>&1 | grep '"B"' | cut -c1-100 | tee -a $o/${tag}_quick_detect.log
done
