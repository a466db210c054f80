#!/bin/bash
set -u
OUT=gpurun_out/c11
mkdir -p $OUT
say() { echo "$1" | tee -a $OUT/summary.txt; }
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 ))s; }
timeout 300 python -m pytest tests/test_gpu_lazy_tables.py -q --timeout 300 -p no:cacheprovider > $OUT/lazy.log 2>&1; say "lazy rc=$? $(el)"
timeout 300 python -m pytest tests/test_gpu_engines.py -q --timeout 300 -p no:cacheprovider -k "tc5" > $OUT/tc5_tests.log 2>&1; say "tc5 tests rc=$? $(el)"
timeout 300 python bench.py --workload emcdr_map --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_map_tc5.json 2> $OUT/bench_map_tc5.err; say "bench map tc5 rc=$? $(el)"
XDR_SECTIONS=emcdr_map_step XDR_BENCH_TC5=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc5_mlp -s 6 -c 2 -o $OUT/tc5_mlp \
  python scripts/bench_new_kernels.py > $OUT/map_ncu.log 2>&1; say "ncu full tc5 rc=$? $(el)"
python -c "
import json
d=json.loads(open('$OUT/bench_map_tc5.json').read().strip().splitlines()[-1]); print('map tc5: %.2f us/step'%(d['ms_per_step']*1e3))"
cat $OUT/summary.txt
