#!/bin/bash
mkdir -p gpurun_out
DDRL_NARROW_W1=0 timeout 200 python tools/micro_sac.py > gpurun_out/i_micro_nonarrow.log 2>&1; cat gpurun_out/i_micro_nonarrow.log
DDRL_NO_GRAPH=1 timeout 200 python tools/micro_sac.py > gpurun_out/i_micro_nograph.log 2>&1; cat gpurun_out/i_micro_nograph.log
DDRL_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/i_launches.csv python tools/prof_sac.py C2 4 > gpurun_out/i_prof.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/i_launches.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]
for r in rows[hdr+1:]:
    if len(r)>=len(h): print(r[h.index('Kernel Name')][:60].ljust(60), r[h.index('Grid Size')] if 'Grid Size' in h else '', r[-1])
PY
