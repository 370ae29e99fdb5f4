#!/bin/bash
# developer build with phase clocks: per-CTA statistics of the detect stream kernel, phase times of the sweep CTA of image 0
o=gpurun_out; tag=${1:-dev}; shift
cp objectdetection_ssd_b200/libssdhead.so /tmp/lib_keep.so
SSDHEAD_NVCC_EXTRA="-DSSDHEAD_PHASE_TIMES $SSDHEAD_DEV_EXTRA" python -m objectdetection_ssd_b200.build --force > /dev/null && for B in "$@"; do timeout 300 python tools/stream_phases.py $B 2>&1 | grep -v "create priors" | tee -a $o/${tag}_phases.log; done
cp /tmp/lib_keep.so objectdetection_ssd_b200/libssdhead.so
