#!/bin/bash
# call 42 (last GPU seconds of the round): the independent launches of a cross-stitch layer on parallel streams (ops.set_cross_streams)
set -u
OUT=gpurun_out/c42
mkdir -p $OUT
say() { echo "$1" | tee -a $OUT/summary.txt; }
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 ))s; }
timeout 60 python -m pytest tests/test_gpu_tc5_dense.py tests/test_gpu_trainer.py tests/test_gpu_models.py tests/test_gpu_kernels.py -q -m gpu --timeout 50 \
  -k "cross_pair or graphed_conet or conet or frob" -p no:cacheprovider > $OUT/tests.log 2>&1; say "new tests rc=$? $(el)"
tail -3 $OUT/tests.log; grep -E "^(FAILED|ERROR)" $OUT/tests.log | head
XDR_CROSS_STREAMS=4 timeout 40 python -m pytest tests/test_gpu_models.py tests/test_gpu_variants.py -q -m gpu --timeout 30 -k "conet or CoNet" -p no:cacheprovider > $OUT/tests_streams.log 2>&1; say "conet tests, 4 streams rc=$? $(el)"
tail -2 $OUT/tests_streams.log
for n in 4 2; do
XDR_CROSS_STREAMS=$n timeout 40 python bench.py --workload conet_5m --steps 10 --warmup 3 --repeats 3 --no-cpu-baseline > $OUT/conet_streams$n.json 2> $OUT/conet_streams$n.err; say "conet $n streams rc=$? $(el)"
python - <<PY
import json
try:
    d = json.loads(open('$OUT/conet_streams$n.json').read().strip().splitlines()[-1])
    print('streams $n: us/step %.2f [%s .. %s] e2e %.3e loss %s' % (d['ms_per_step'] * 1e3, d['timing'].get('min_ms'), d['timing'].get('max_ms'), d['e2e']['value'], d.get('loss_mean')))
except Exception as e:
    print('ERR', e, open('$OUT/conet_streams$n.err').read()[-500:])
PY
done
