#!/bin/bash
# developer call: per-kernel durations / instructions / DRAM bytes of the detect launches (ncu, serialised)
o=gpurun_out; tag=${1:-dev}; shift
for B in "$@"; do
python tools/prof_detect.py $B 6.0 > $o/${tag}_plain_$B.log 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:detect_ --launch-skip 3 -c 3 --csv --log-file $o/${tag}_launches_$B.csv python tools/prof_detect.py $B 6.0 > $o/${tag}_ncu_$B.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('$o/${tag}_launches_$B.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]
for r in rows[hdr+1:]:
    d=dict(zip(h,r))
    print($B, d['ID'], d['Kernel Name'][:40], d['Metric Name'], d['Metric Value'], d['Grid Size'])
PY
done
