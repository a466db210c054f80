#!/bin/bash
# call 32: eager map step with the torch glue trimmed (no reduction of a 0-d loss, gradients adopted instead of copied)
set -u
OUT=gpurun_out/c32
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_engines.py tests/test_gpu_models.py tests/test_gpu_trainer.py -q --timeout 300 -p no:cacheprovider > $OUT/tests.log 2>&1; echo "engines/models/trainer tests rc=$?"
tail -3 $OUT/tests.log
timeout 300 python bench.py --workload emcdr_map --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_map.json 2> $OUT/bench_map.err; echo "bench emcdr_map rc=$?"
XDR_EAGER_TC5=0 timeout 300 python bench.py --workload emcdr_map --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_map_noeager.json 2> $OUT/bench_map_noeager.err; echo "bench emcdr_map (eager off) rc=$?"
python - <<PY
import json
for f in ('bench_map','bench_map_noeager'):
    try:
        d = json.loads(open('$OUT/' + f + '.json').read().strip().splitlines()[-1])
        print(f, 'value %.3e us/step %.2f e2e %.3e launches %s' % (d['value'], d['ms_per_step'] * 1e3, d['e2e']['value'], d['gpu_launches']))
    except Exception as e:
        print(f, 'ERR', e, open('$OUT/' + f + '.err').read()[-900:])
PY
