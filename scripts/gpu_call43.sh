#!/bin/bash
# call 43 (the round's last GPU seconds): the default state -- two lanes inside a graph capture -- tests of the touched paths + the conet_5m line
set -u
OUT=gpurun_out/c43
mkdir -p $OUT
timeout 45 python -m pytest tests/test_gpu_tc5_dense.py tests/test_gpu_trainer.py tests/test_gpu_models.py tests/test_gpu_variants.py tests/test_gpu_kernels.py -q -m gpu --timeout 40 \
  -k "cross_pair or graphed or conet or CoNet or frob" -p no:cacheprovider > $OUT/tests.log 2>&1; echo "tests rc=$?" | tee $OUT/summary.txt
tail -3 $OUT/tests.log; grep -E "^(FAILED|ERROR)" $OUT/tests.log | head
timeout 40 python bench.py --workload conet_5m --steps 10 --warmup 3 --repeats 5 --no-cpu-baseline > $OUT/conet.json 2> $OUT/conet.err; echo "conet rc=$?" | tee -a $OUT/summary.txt
python - <<PY
import json
try:
    d = json.loads(open('$OUT/conet.json').read().strip().splitlines()[-1])
    print('default: us/step %.2f [%s .. %s] e2e %.3e loss %s lanes %s' % (d['ms_per_step'] * 1e3, d['timing'].get('min_ms'), d['timing'].get('max_ms'), d['e2e']['value'], d.get('loss_mean'), d['config'].get('graph_lanes')))
except Exception as e:
    print('ERR', e, open('$OUT/conet.err').read()[-500:])
PY
