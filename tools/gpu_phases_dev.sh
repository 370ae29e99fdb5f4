#!/bin/bash
# developer build with phase clocks: per-phase times of the sweep CTA of image 0
o=gpurun_out; tag=${1:-dev}
cp objectdetection_ssd_b200/libssdhead.so /tmp/lib_keep.so
SSDHEAD_NVCC_EXTRA=-DSSDHEAD_PHASE_TIMES python -m objectdetection_ssd_b200.build --force > /dev/null && for v in 1; do timeout 300 python tools/detect_phases.py 1 6.0 2>&1 | grep -v "create priors" | tee -a $o/${tag}_phases.log; done
cp /tmp/lib_keep.so objectdetection_ssd_b200/libssdhead.so
