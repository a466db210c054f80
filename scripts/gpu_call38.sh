#!/bin/bash
# call 38: FusedStepRunner through xdr_train_steps_host (one library call per chunk)
set -u
OUT=gpurun_out/c38
mkdir -p $OUT
timeout 500 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_steps.py -q --timeout 300 -p no:cacheprovider > $OUT/tests.log 2>&1; echo "trainer/steps tests rc=$?"
tail -3 $OUT/tests.log
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/b$i.json 2> $OUT/b$i.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('$OUT/b$i.json').read().strip().splitlines()[-1])
e = d['e2e']
print('run $i: frac %.4f e2e %.3e' % (d['roofline']['frac'], e['value']), e['samples_ms'], 'loss', e['loss_mean'])
PY
done
