#!/bin/bash
# call 41: final verification of the round -- CoNet on the tcgen05 dense engine by default (per-model engine), faster frob_sum_fwd:
# whole GPU suite, smoke, the default bench line, the conet_5m line, one ncu launch list of a CoNet step
set -u
OUT=gpurun_out/c41
mkdir -p $OUT
say() { echo "$1" | tee -a $OUT/summary.txt; }
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 ))s; }
timeout 400 python -m pytest tests/ -q -m gpu --timeout 300 -p no:cacheprovider > $OUT/gpu_suite.log 2>&1; say "gpu suite rc=$? $(el)"
tail -3 $OUT/gpu_suite.log; grep -E "^(FAILED|ERROR)" $OUT/gpu_suite.log | head -20
timeout 100 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1; say "smoke rc=$? $(el)"
tail -1 $OUT/smoke.log
timeout 200 python bench.py --workload conet_5m --steps 10 --warmup 3 --repeats 5 > $OUT/conet.json 2> $OUT/conet.err; say "conet_5m (default) rc=$? $(el)"
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_k20.json 2> $OUT/bench_k20.err; say "bench default rc=$? $(el)"
python - <<PY
import json
for f in ('bench_k20', 'conet'):
    try:
        d = json.loads(open('$OUT/' + f + '.json').read().strip().splitlines()[-1])
        print(f, 'value %.4e us/step %.2f [%s .. %s] frac %.4f e2e %s launches %s loss %s engine %s' % (
            d['value'], d['ms_per_step'] * 1e3, d['timing'].get('min_ms'), d['timing'].get('max_ms'), d['roofline']['frac'],
            d.get('e2e') and '%.3e' % d['e2e']['value'], d.get('gpu_launches'), d.get('loss_mean'), d['config'].get('engine')))
    except Exception as e:
        print(f, 'ERR', e, open('$OUT/' + f + '.err').read()[-600:])
PY
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -s 690 -c 100 --csv --log-file $OUT/conet_launches_ncu.csv \
  python bench.py --workload conet_5m --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline > $OUT/conet_ncu.log 2>&1; say "ncu launch list rc=$? $(el)"
cat $OUT/summary.txt
