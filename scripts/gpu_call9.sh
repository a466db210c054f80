#!/bin/bash
set -u
OUT=gpurun_out/c9
mkdir -p $OUT
say() { echo "$1" | tee -a $OUT/summary.txt; }
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 ))s; }
LIBDIR=$PWD/recbole-cdr_b200/recbole_cdr_b200/lib
timeout 300 python -m pytest tests/test_gpu_lazy_tables.py tests/test_gpu_steps.py -q --timeout 300 -p no:cacheprovider > $OUT/lazy.log 2>&1; say "lazy + steps rc=$? $(el)"
for v in "" _vG _vH _vA; do
  for coop in 1 0; do
    for k in 20 200; do
      XDR_LIB=$LIBDIR/libxdr$v.so timeout 200 python bench.py --steps $k --warmup 5 --repeats 9 --no-cpu-baseline --no-e2e --no-extras --grad-mode accumulate --coop $coop \
        > $OUT/bench${v}_c${coop}_k$k.json 2> $OUT/bench${v}_c${coop}_k$k.err
      python - <<PY | tee -a $OUT/summary.txt
import json
try:
    d = json.loads(open('$OUT/bench${v}_c${coop}_k$k.json').read().strip().splitlines()[-1])
    print('lib$v accumulate coop=$coop K=$k: %.3f us/step frac %.3f (min %.3f)' % (d['ms_per_step'] * 1e3, d['roofline']['frac'], d['timing']['min_ms']/$k*1e3))
except Exception as e:
    print('lib$v K=$k: FAILED', e)
PY
    done
  done
done
say "variants done $(el)"
timeout 300 python bench.py --workload emcdr_map --steps 20 --warmup 5 > $OUT/bench_map_tc5.json 2> $OUT/bench_map_tc5.err; say "bench map tc5 rc=$? $(el)"
timeout 300 python bench.py --workload emcdr_map --steps 20 --warmup 5 --map-engine "" --no-cpu-baseline > $OUT/bench_map_composed.json 2> $OUT/bench_map_composed.err; say "bench map composed rc=$? $(el)"
XDR_SECTIONS=emcdr_map_step XDR_BENCH_TC5=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/map_launches.csv \
  python scripts/bench_new_kernels.py > $OUT/map_ncu.log 2>&1; say "map step ncu launch list rc=$? $(el)"
cut -c1-400 $OUT/bench_map_tc5.json; tail -3 $OUT/bench_map_tc5.err
cat $OUT/summary.txt
