timeout 600 python -m pytest --timeout 300 tests/test_gpu_detect_shortlist.py tests/test_gpu_detect.py tests/test_gpu_levels.py tests/test_gpu_dropin.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py tests/test_gpu_ctx.py -m gpu -x -q 2>&1 | tail -3
python tools/quick_bench_detect.py 1 8 32 128 256 2>&1 | grep '"B"' | cut -c1-90
echo exhaustive
SSDHEAD_DETECT_SHORTLIST=0 python tools/quick_bench_detect.py 1 8 32 128 256 2>&1 | grep '"B"' | cut -c1-90
