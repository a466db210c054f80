#!/bin/bash
# call 29: register-blocked, double-buffered dense forward tiles (dense_fwd2_kernel)
set -u
OUT=gpurun_out/c29
mkdir -p $OUT
say() { echo "$1" | tee -a $OUT/summary.txt; }
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 ))s; }
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py tests/test_gpu_variants.py tests/test_gpu_engines.py -q --timeout 300 -p no:cacheprovider > $OUT/tests.log 2>&1; say "kernels/models/variants/engines tests rc=$? $(el)"
tail -4 $OUT/tests.log
timeout 300 python scripts/bench_dense_engines.py > $OUT/dense_engines.jsonl 2> $OUT/dense_engines.err; say "dense per shape rc=$? $(el)"
cut -c1-260 $OUT/dense_engines.jsonl
XDR_DENSE_FWD2=0 timeout 300 python scripts/bench_dense_engines.py > $OUT/dense_engines_old.jsonl 2> /dev/null; say "dense per shape, old forward rc=$? $(el)"
cut -c1-120 $OUT/dense_engines_old.jsonl
timeout 300 python bench.py --workload conet_5m --steps 10 --warmup 3 --repeats 5 --no-cpu-baseline > $OUT/bench_conet.json 2> $OUT/bench_conet.err; say "bench conet rc=$? $(el)"
python - <<PY
import json
d = json.loads(open('$OUT/bench_conet.json').read().strip().splitlines()[-1])
print('conet us/step %.1f e2e %.3e' % (d['ms_per_step'] * 1e3, d['e2e']['value']))
PY
cat $OUT/summary.txt
