#!/bin/bash
set -u
OUT=gpurun_out/c17
mkdir -p $OUT
say() { echo "$1" | tee -a $OUT/summary.txt; }
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 ))s; }
timeout 300 python bench.py --workload emcdr_map --steps 20 --warmup 5 > $OUT/bench_map.json 2> $OUT/bench_map.err; say "bench emcdr_map rc=$? $(el)"
timeout 300 python bench.py --workload emcdr_map --steps 20 --warmup 5 --map-engine tc5 --no-cpu-baseline > $OUT/bench_map_tc5.json 2> $OUT/bench_map_tc5.err; say "bench emcdr_map tc5 rc=$? $(el)"
timeout 400 python bench.py --workload conet_5m --steps 10 --warmup 3 --repeats 5 > $OUT/bench_conet.json 2> $OUT/bench_conet.err; say "bench conet_5m rc=$? $(el)"
timeout 300 python bench.py --workload conet_5m --steps 10 --warmup 3 --repeats 5 --dense-engine 1 --no-cpu-baseline > $OUT/bench_conet_tc5.json 2> $OUT/bench_conet_tc5.err; say "bench conet_5m tcgen05 dense rc=$? $(el)"
python - <<PY
import json
for f in ('bench_map','bench_map_tc5','bench_conet','bench_conet_tc5'):
    try:
        d = json.loads(open('$OUT/' + f + '.json').read().strip().splitlines()[-1])
        print(f, 'value %.3e us/step %.2f frac %.4f e2e %.3e launches %s cpu %s' % (d['value'], d['ms_per_step'] * 1e3, d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'], d.get('cpu_baseline', {}).get('value')))
    except Exception as e:
        print(f, 'ERR', e, open('$OUT/' + f + '.err').read()[-900:])
PY
cat $OUT/summary.txt
