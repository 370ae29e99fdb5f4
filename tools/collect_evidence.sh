#!/bin/bash
# Turn the outputs of `gpu_round.sh <tag>` + `gpu_evidence.sh <tag>` (gpurun_out/<tag>_*) into the tracked files under
# profiles/r2/ (bench lines, test log, launch list, ncu summaries, per-line stall tables) and re-stamp profiles/traffic.json.
# usage: collect_evidence.sh <tag> [<old tag whose files are replaced>]
set -e
tag=$1; old=$2
o=gpurun_out; d=profiles/r2
python tools/make_traffic.py $tag > /dev/null
for f in default detect256 detect64 stress short; do cp $o/${tag}_bench_$f.json $d/${tag}_bench_$f.json; done
cp $o/${tag}_bench_launches.csv $o/${tag}_pytest_gpu.log $o/${tag}_source_hash.txt $d/
python tools/ncu_summary.py $o/${tag}_loss_b256.ncu-rep > $d/${tag}_loss_kernels_ncu_summary.txt 2>&1
(echo "## batch 64 (automatic route: exhaustive)"; python tools/ncu_summary.py $o/${tag}_detect_b64.ncu-rep; echo; echo "## batch 256 (automatic route: short list)"; python tools/ncu_summary.py $o/${tag}_detect_b256.ncu-rep) > $d/${tag}_detect_kernels_ncu_summary.txt 2>&1
tmp=$(mktemp -d); (cd $tmp && cuobjdump -xelf all $OLDPWD/objectdetection_ssd_b200/libssdhead.so > /dev/null 2>&1)
python tools/ncu_lines.py $o/${tag}_detect_b256.ncu-rep $tmp/detect.sm_100a.cubin "detect_stream_kernelILi21ELb0" 30 detect_stream > $d/${tag}_detect_stream_kernel_stall_lines.txt 2>&1
python tools/ncu_lines.py $o/${tag}_detect_b256.ncu-rep $tmp/detect.sm_100a.cubin "detect_sweep_kernelILi21ELb0" 30 detect_sweep > $d/${tag}_detect_sweep_kernel_stall_lines.txt 2>&1
python tools/ncu_lines.py $o/${tag}_loss_b256.ncu-rep $tmp/loss.sm_100a.cubin "ce_stream_kernelILi21ELb1ELb1" 30 ce_stream > $d/${tag}_ce_stream_kernel_stall_lines.txt 2>&1
python tools/ncu_lines.py $o/${tag}_loss_b256.ncu-rep $tmp/loss.sm_100a.cubin "mine_kernelILi21ELb1ELb1" 30 mine_kernel > $d/${tag}_mine_kernel_stall_lines.txt 2>&1
rm -rf $tmp
if [ -n "$old" ]; then for f in $d/${old}_*; do git rm -q --cached $f 2>/dev/null || true; rm -f $f; done; fi
ls $d | grep "^${tag}_"
