#!/bin/bash
# developer call: time the detect path with every variants/det_*.so (built here with SSDHEAD_NVCC_EXTRA), restore the library
cp objectdetection_ssd_b200/libssdhead.so /tmp/lib_keep.so
for v in variants/det_*.so; do
  cp $v objectdetection_ssd_b200/libssdhead.so
  echo "== $v"
  timeout 120 python tools/quick_bench_detect.py "$@" 2>&1 | grep '"B"' | grep -v '"bias": [48]' | cut -c1-80
done
cp /tmp/lib_keep.so objectdetection_ssd_b200/libssdhead.so
