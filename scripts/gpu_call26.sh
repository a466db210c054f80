#!/bin/bash
# call 26: launch list of one CoNet step (composed fp32 path) and of one DTCDR-like step
set -u
OUT=gpurun_out/c26
mkdir -p $OUT
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/conet_launches.csv \
  python bench.py --workload conet_5m --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1; echo "ncu conet rc=$?"
python - <<PY
import csv,re,collections
rows=list(csv.reader(open('$OUT/conet_launches.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]; ki=H.index('Kernel Name'); vi=H.index('Metric Value')
seq=[(re.sub(r'\(.*','',r[ki])[:70], float(r[vi].replace(',',''))/1e3) for r in rows[hdr+1:] if len(r)>vi]
# the last graph-replayed step: take the last 112 launches
last=seq[-112:]
agg=collections.defaultdict(lambda:[0,0.0])
for n,t in last: agg[n][0]+=1; agg[n][1]+=t
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(),key=lambda t:-t[1][1])[:16]: print('%-72s n=%3d %8.1f us %5.1f%% avg %.1f'%(k,v[0],v[1],100*v[1]/tot,v[1]/v[0]))
print('sum of the last 112 launches: %.1f us'%tot)
PY
