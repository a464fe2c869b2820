#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== capture debug"; TMPI_DEBUG_CAPTURE=1 timeout 300 python bench.py --steps 5 --warmup 3 2>&1 | tail -12 | tee gpurun_out/capture_dbg.log
echo "=== launch list (one eager step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/ncu_launch.log 2>&1; tail -3 gpurun_out/ncu_launch.log
python - <<'PY'
import csv, collections
rows=[]
with open('gpurun_out/launches.csv') as f:
    lines=[l for l in f if not l.startswith('==')]
r=csv.DictReader(lines)
tot=collections.OrderedDict(); order=[]
for row in r:
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    name=row['Kernel Name'][:70]; v=float(row['Metric Value'].replace(',',''))
    unit=row['Metric Unit']
    us = v/1000 if unit in ('ns','nsecond') else (v if unit in ('us','usecond') else v*1000)
    order.append((name,us))
tot_us=sum(u for _,u in order)
print("kernels in step: %d, serialized total %.1f us"%(len(order),tot_us))
agg=collections.defaultdict(lambda:[0,0.0])
for n,u in order: agg[n][0]+=1; agg[n][1]+=u
for n,(c,u) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]:
    print("%8.1f us %5.1f%%  x%-3d %s"%(u,100*u/tot_us,c,n))
open('gpurun_out/step_order.txt','w').write("\n".join("%9.1f  %s"%(u,n) for n,u in order))
PY
echo "=== 2-GPU part skipped on 1-GPU box" 
