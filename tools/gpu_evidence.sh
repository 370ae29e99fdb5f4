#!/bin/bash
# Evidence call (one GPU): ncu --set full of the kernels of the benchmarked binary - the two loss kernels at batch 256
# (dense step), the detect kernels at batch 64 and 256 - each after the same command exited 0 without ncu.
o=gpurun_out; tag=${1:-r2}
python tools/prof_loss.py 256 4 > $o/${tag}_prof_loss_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ce_stream|mine_kernel" --launch-skip 4 -c 4 -o $o/${tag}_loss_b256 -f python tools/prof_loss.py 256 4 > $o/${tag}_ncu_loss.log 2>&1; echo "loss ncu rc=$?"
for B in 64 256; do
  python tools/prof_detect.py $B 6.0 > $o/${tag}_prof_detect_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:detect_ --launch-skip 2 -c 4 -o $o/${tag}_detect_b$B -f python tools/prof_detect.py $B 6.0 > $o/${tag}_ncu_detect$B.log 2>&1; echo "detect $B ncu rc=$?"
done
python -c "from objectdetection_ssd_b200 import build; print('source_hash', build.source_hash())" | tee $o/${tag}_source_hash.txt
