#!/bin/bash
# call 22: lazy tables, rolling loader pipeline with guarded waits + publisher warp ("claim and fill at gather time")
set -u
OUT=gpurun_out/c22
mkdir -p $OUT
say() { echo "$1" | tee -a $OUT/summary.txt; }
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 ))s; }
timeout 400 python -m pytest tests/test_gpu_hot_rows.py tests/test_gpu_steps.py tests/test_gpu_lazy_tables.py tests/test_gpu_trainer.py -q --timeout 300 -p no:cacheprovider > $OUT/steps.log 2>&1; say "steps/hot/lazy/trainer tests rc=$? $(el)"
tail -3 $OUT/steps.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-e2e > $OUT/bench_k20.json 2> $OUT/bench_k20.err; say "bench K=20 rc=$? $(el)"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-e2e --grad-mode fresh > $OUT/bench_k20_fresh.json 2> $OUT/bench_k20_fresh.err; say "bench K=20 fresh rc=$? $(el)"
timeout 300 python bench.py --steps 200 --warmup 5 --repeats 5 --no-extras --no-cpu-baseline --no-e2e > $OUT/bench_k200.json 2> $OUT/bench_k200.err; say "bench K=200 rc=$? $(el)"
timeout 300 python bench.py --steps 200 --warmup 5 --repeats 5 --no-extras --no-cpu-baseline --no-e2e --grad-mode fresh > $OUT/bench_k200_fresh.json 2> $OUT/bench_k200_fresh.err; say "bench K=200 fresh rc=$? $(el)"
timeout 200 python scripts/trace_steps.py > $OUT/trace_plain.txt 2>&1; say "trace plain rc=$? $(el)"
timeout 200 python scripts/trace_steps.py lazy > $OUT/trace_lazy.txt 2>&1; say "trace lazy rc=$? $(el)"
python - <<PY
import json
for f in ('bench_k20','bench_k20_fresh','bench_k200','bench_k200_fresh'):
    try:
        d = json.loads(open('$OUT/' + f + '.json').read().strip().splitlines()[-1])
        print(f, 'value %.3e us/step %.3f frac %.4f' % (d['value'], d['ms_per_step'] * 1e3, d['roofline']['frac']), d['timing'])
    except Exception as e:
        print(f, 'ERR', e, open('$OUT/' + f + '.err').read()[-600:])
PY
tail -12 $OUT/trace_plain.txt
tail -8 $OUT/trace_lazy.txt
cat $OUT/summary.txt
