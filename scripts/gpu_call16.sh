#!/bin/bash
set -u
OUT=gpurun_out/c16
mkdir -p $OUT
say() { echo "$1" | tee -a $OUT/summary.txt; }
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 ))s; }
timeout 300 python -m pytest tests/test_gpu_hot_rows.py tests/test_gpu_steps.py tests/test_gpu_lazy_tables.py -q --timeout 300 -p no:cacheprovider > $OUT/steps.log 2>&1; say "steps/hot/lazy tests rc=$? $(el)"
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_k20.json 2> $OUT/bench_k20.err; say "bench K=20 rc=$? $(el)"
timeout 300 python bench.py --steps 200 --warmup 5 --repeats 5 --no-extras --no-cpu-baseline --no-e2e > $OUT/bench_k200.json 2> $OUT/bench_k200.err; say "bench K=200 rc=$? $(el)"
timeout 300 python bench.py --workload conet_5m --steps 10 --warmup 3 --repeats 5 > $OUT/bench_conet.json 2> $OUT/bench_conet.err; say "bench conet_5m rc=$? $(el)"
timeout 300 python bench.py --workload emcdr_map --steps 20 --warmup 5 > $OUT/bench_map.json 2> $OUT/bench_map.err; say "bench emcdr_map rc=$? $(el)"
timeout 300 python bench.py --workload emcdr_map --steps 20 --warmup 5 --map-engine tc5 --no-cpu-baseline > $OUT/bench_map_tc5.json 2> $OUT/bench_map_tc5.err; say "bench emcdr_map tc5 rc=$? $(el)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 20 --warmup 5 --repeats 3 --no-extras --no-cpu-baseline --no-e2e > /dev/null 2>&1; say "ncu launch list rc=$? $(el)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:train_steps_staged -s 1 -c 1 -o $OUT/staged_k20 \
  python bench.py --steps 20 --warmup 5 --repeats 2 --no-extras --no-cpu-baseline --no-e2e > /dev/null 2>&1; say "ncu full rc=$? $(el)"
tail -3 $OUT/steps.log
python - <<PY
import json
for f in ('bench_k20','bench_k200','bench_conet','bench_map','bench_map_tc5'):
    try:
        d = json.loads(open('$OUT/' + f + '.json').read().strip().splitlines()[-1])
        print(f, 'value %.3e us/step %.2f frac %.4f e2e %s' % (d['value'], d['ms_per_step'] * 1e3, d['roofline']['frac'], d['e2e'] and '%.3e' % d['e2e']['value']), json.dumps(d.get('variants', ''))[:900])
    except Exception as e:
        print(f, 'ERR', e, open('$OUT/' + f + '.err').read()[-600:])
PY
cat $OUT/summary.txt
