#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02g_launches_vasnet_train_step.csv python scripts/vasnet_one_step.py > gpurun_out/r02g_ncu_step.log 2>&1
tail -3 gpurun_out/r02g_ncu_step.log
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02g_launches_vasnet_train_step.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
ks=[(r[ki], float(r[vi].replace(',',''))/ (1000.0 if r[ui]=='ns' else 1.0)) for r in rows[1:]]
n=len(ks)//3
last=ks[-n:]
print(len(ks), 'launches,', n, 'per step; last step sum us:', sum(v for _,v in last))
for k,v in last: print(f'{v:8.2f}  {k[:100]}')
P
