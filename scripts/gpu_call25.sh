#!/bin/bash
# call 25: EMCDR map step on tcgen05 by default; the whole GPU suite with the tcgen05 dense engine switched on by the environment
set -u
OUT=gpurun_out/c25
mkdir -p $OUT
say() { echo "$1" | tee -a $OUT/summary.txt; }
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 ))s; }
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $OUT/gpu_suite.log 2>&1; say "gpu suite rc=$? $(el)"
tail -3 $OUT/gpu_suite.log
XDR_DENSE_ENGINE=1 timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $OUT/gpu_suite_dense1.log 2>&1; say "gpu suite, tcgen05 dense engine on, rc=$? $(el)"
tail -15 $OUT/gpu_suite_dense1.log
timeout 300 python scripts/bench_dense_engines.py > $OUT/dense_engines.jsonl 2> $OUT/dense_engines.err; say "dense engines per shape rc=$? $(el)"
cat $OUT/dense_engines.jsonl | cut -c1-400
timeout 300 python bench.py --workload emcdr_map --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_map.json 2> $OUT/bench_map.err; say "bench emcdr_map (default engine) rc=$? $(el)"
python - <<PY
import json
for f in ('bench_map',):
    try:
        d = json.loads(open('$OUT/' + f + '.json').read().strip().splitlines()[-1])
        print(f, 'value %.3e us/step %.2f frac %.4f e2e %.3e launches %s' % (d['value'], d['ms_per_step'] * 1e3, d['roofline']['frac'], d['e2e']['value'], d['gpu_launches']))
    except Exception as e:
        print(f, 'ERR', e, open('$OUT/' + f + '.err').read()[-900:])
PY
cat $OUT/summary.txt
