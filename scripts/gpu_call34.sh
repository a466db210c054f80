#!/bin/bash
# call 34: compute-sanitizer memcheck over the kernels that changed this round (lazy tables, dense_fwd2 tiles, tc5_mlp eager / correction)
set -u
OUT=gpurun_out/c34
mkdir -p $OUT
say() { echo "$1" | tee -a $OUT/summary.txt; }
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 ))s; }
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 77 --log-file $OUT/memcheck_lazy.log python -m pytest tests/test_gpu_lazy_tables.py tests/test_gpu_hot_rows.py -q -x --timeout 400 -p no:cacheprovider -k "not full" > $OUT/memcheck_lazy.out 2>&1; say "memcheck lazy/hot rc=$? $(el)"
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 77 --log-file $OUT/memcheck_dense.log python -m pytest tests/test_gpu_kernels.py -q -x --timeout 400 -p no:cacheprovider -k "dense_fwd_bwd or cross_stitch" > $OUT/memcheck_dense.out 2>&1; say "memcheck dense rc=$? $(el)"
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 77 --log-file $OUT/memcheck_tc5.log python -m pytest tests/test_gpu_engines.py -q -x --timeout 400 -p no:cacheprovider -k "tc5_mlp and (129 or 1000 or 127)" > $OUT/memcheck_tc5.out 2>&1; say "memcheck tc5_mlp rc=$? $(el)"
for f in lazy dense tc5; do tail -2 $OUT/memcheck_$f.out; grep -c "Invalid\|ERROR SUMMARY" $OUT/memcheck_$f.log; tail -2 $OUT/memcheck_$f.log; done
cat $OUT/summary.txt
