#!/bin/bash
# call 24: bf16 plane split on the conversion instruction (cvt.rn.bf16x2.f32): tcgen05 engines re-validated and re-timed
set -u
OUT=gpurun_out/c24
mkdir -p $OUT
say() { echo "$1" | tee -a $OUT/summary.txt; }
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 ))s; }
timeout 600 python -m pytest tests/test_gpu_engines.py tests/test_gpu_tc5_dense.py -q --timeout 300 -p no:cacheprovider > $OUT/engines.log 2>&1; say "engine tests rc=$? $(el)"
tail -3 $OUT/engines.log
timeout 300 python bench.py --workload emcdr_map --steps 20 --warmup 5 --map-engine tc5 --no-cpu-baseline > $OUT/bench_map_tc5.json 2> $OUT/bench_map_tc5.err; say "bench emcdr_map tc5 rc=$? $(el)"
timeout 300 python bench.py --workload emcdr_map --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_map.json 2> $OUT/bench_map.err; say "bench emcdr_map composed rc=$? $(el)"
timeout 300 python bench.py --workload conet_5m --steps 10 --warmup 3 --repeats 5 --dense-engine 1 --no-cpu-baseline > $OUT/bench_conet_tc5.json 2> $OUT/bench_conet_tc5.err; say "bench conet_5m tcgen05 dense rc=$? $(el)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/map_tc5_launches.csv \
  python bench.py --workload emcdr_map --steps 3 --warmup 3 --repeats 1 --map-engine tc5 --no-cpu-baseline --no-e2e > /dev/null 2>&1; say "ncu launch list (map tc5) rc=$? $(el)"
python - <<PY
import json
for f in ('bench_map_tc5','bench_map','bench_conet_tc5'):
    try:
        d = json.loads(open('$OUT/' + f + '.json').read().strip().splitlines()[-1])
        print(f, 'value %.3e us/step %.2f frac %.4f e2e %.3e launches %s' % (d['value'], d['ms_per_step'] * 1e3, d['roofline']['frac'], d['e2e']['value'], d['gpu_launches']))
    except Exception as e:
        print(f, 'ERR', e, open('$OUT/' + f + '.err').read()[-900:])
PY
tail -40 $OUT/map_tc5_launches.csv | cut -d, -f5,12,14- | tail -30
cat $OUT/summary.txt
