#!/bin/bash
# Final verification of a round, as the driver will do it: GPU suite, smoke(), the default bench line, the reference arm, model-step lines
set -u
OUT=gpurun_out/final
mkdir -p $OUT
say() { echo "$1" | tee -a $OUT/summary.txt; }
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 ))s; }
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > $OUT/gpu_suite.log 2>&1; say "gpu suite rc=$? $(el)"
tail -3 $OUT/gpu_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1; say "smoke rc=$? $(el)"
tail -1 $OUT/smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_k20.json 2> $OUT/bench_k20.err; say "bench default rc=$? $(el)"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; say "bench reference arm rc=$? $(el)"
timeout 300 python bench.py --steps 200 --warmup 5 --repeats 5 --no-extras --no-cpu-baseline --no-e2e > $OUT/bench_k200.json 2> $OUT/bench_k200.err; say "bench K=200 rc=$? $(el)"
timeout 300 python bench.py --workload emcdr_map --steps 20 --warmup 5 > $OUT/bench_map.json 2> $OUT/bench_map.err; say "bench emcdr_map rc=$? $(el)"
timeout 400 python bench.py --workload conet_5m --steps 10 --warmup 3 --repeats 5 > $OUT/bench_conet.json 2> $OUT/bench_conet.err; say "bench conet_5m rc=$? $(el)"
python - <<PY
import json
for f in ('bench_k20','bench_ref','bench_k200','bench_map','bench_conet'):
    try:
        d = json.loads(open('$OUT/' + f + '.json').read().strip().splitlines()[-1])
        print(f, 'value %.4e %s us/step %.3f frac %s e2e %s launches %s cpu %s' % (d['value'], d['unit'], d['ms_per_step'] * 1e3, d.get('roofline', {}).get('frac'), d.get('e2e') and '%.3e' % d['e2e']['value'], d.get('gpu_launches'), d.get('cpu_baseline', {}).get('value')))
    except Exception as e:
        print(f, 'ERR', e, open('$OUT/' + f + '.err').read()[-600:])
PY
cat $OUT/summary.txt
