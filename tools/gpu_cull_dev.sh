#!/bin/bash
# developer call: the seeded cull of the fused match - loss tests with the cull forced for every batch, then the stress timing
SSDHEAD_MATCH_CULL_MIN=1 timeout 900 python -m pytest --timeout 300 tests/test_gpu_loss.py tests/test_gpu_golden.py tests/test_gpu_ctx.py tests/test_gpu_fullsize.py tests/test_gpu_dropin.py -m gpu -q 2>&1 | grep -E "^E  |passed|failed|Error" | head -30
timeout 900 python -m pytest --timeout 300 tests/test_gpu_loss.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
for m in 24 1000000; do echo "SSDHEAD_MATCH_CULL_MIN=$m"; SSDHEAD_MATCH_CULL_MIN=$m python tools/quick_bench_stress.py 2>&1 | grep -v "create priors" | tail -3 | cut -c1-200; done
