#!/bin/bash
set -u
OUT=gpurun_out/c15
mkdir -p $OUT
say() { echo "$1" | tee -a $OUT/summary.txt; }
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 ))s; }
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > $OUT/gpu_suite.log 2>&1; say "gpu suite rc=$? $(el)"
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; say "smoke rc=$? $(el)"
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_k20.json 2> $OUT/bench_k20.err; say "bench K=20 rc=$? $(el)"
timeout 300 python bench.py --workload conet_5m --steps 10 --warmup 3 --repeats 5 > $OUT/bench_conet.json 2> $OUT/bench_conet.err; say "bench conet_5m rc=$? $(el)"
timeout 300 python bench.py --workload emcdr_map --steps 20 --warmup 5 > $OUT/bench_map.json 2> $OUT/bench_map.err; say "bench emcdr_map rc=$? $(el)"
timeout 300 python bench.py --workload emcdr_map --steps 20 --warmup 5 --map-engine tc5 --no-cpu-baseline > $OUT/bench_map_tc5.json 2> $OUT/bench_map_tc5.err; say "bench emcdr_map tc5 rc=$? $(el)"
timeout 600 python scripts/bench_bitgcf.py --scale 1.0 --steps 3 --warmup 1 > $OUT/bitgcf_n1.log 2>&1; say "bitgcf scale 1.0 N=1 rc=$? $(el)"
tail -4 $OUT/gpu_suite.log; tail -2 $OUT/bitgcf_n1.log | cut -c1-600
python - <<PY
import json
for f in ('bench_k20','bench_conet','bench_map','bench_map_tc5'):
    try:
        d = json.loads(open('$OUT/' + f + '.json').read().strip().splitlines()[-1])
        print(f, 'value %.3e us/step %.2f frac %.4f e2e %.3e' % (d['value'], d['ms_per_step'] * 1e3, d['roofline']['frac'], d['e2e']['value']), json.dumps(d.get('variants', ''))[:700])
    except Exception as e:
        print(f, 'ERR', e, open('$OUT/' + f + '.err').read()[-600:])
PY
cat $OUT/summary.txt
