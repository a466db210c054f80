#!/bin/bash
set -u
OUT=gpurun_out/c8
mkdir -p $OUT
say() { echo "$1" | tee -a $OUT/summary.txt; }
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 ))s; }
LIBDIR=$PWD/recbole-cdr_b200/recbole_cdr_b200/lib
timeout 300 python -m pytest tests/test_gpu_lazy_tables.py tests/test_gpu_steps.py -q --timeout 300 -p no:cacheprovider > $OUT/lazy.log 2>&1; say "lazy + steps rc=$? $(el)"
timeout 300 python -m pytest tests/test_gpu_engines.py -q --timeout 300 -p no:cacheprovider -k "conet_fused or adagrad" > $OUT/engines_fix.log 2>&1; say "conet/adagrad tests rc=$? $(el)"
for v in "" _vB _vE _vF; do
  for mode in fresh accumulate; do
    for k in 20 200; do
      XDR_LIB=$LIBDIR/libxdr$v.so timeout 200 python bench.py --steps $k --warmup 5 --repeats 7 --no-cpu-baseline --no-e2e --no-extras --grad-mode $mode \
        > $OUT/bench${v}_${mode}_k$k.json 2> $OUT/bench${v}_${mode}_k$k.err
      python - <<PY | tee -a $OUT/summary.txt
import json
try:
    d = json.loads(open('$OUT/bench${v}_${mode}_k$k.json').read().strip().splitlines()[-1])
    print('lib$v $mode K=$k: %.3f us/step frac %.3f' % (d['ms_per_step'] * 1e3, d['roofline']['frac']))
except Exception as e:
    print('lib$v $mode K=$k: FAILED', e)
PY
    done
  done
done
say "variants done $(el)"
timeout 200 python scripts/trace_steps.py lazy > $OUT/trace_lazy.txt 2>&1; say "trace lazy rc=$? $(el)"
tail -9 $OUT/trace_lazy.txt
tail -4 $OUT/lazy.log $OUT/engines_fix.log
cat $OUT/summary.txt
