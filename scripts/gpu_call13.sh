#!/bin/bash
set -u
OUT=gpurun_out/c13
mkdir -p $OUT
say() { echo "$1" | tee -a $OUT/summary.txt; }
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 ))s; }
timeout 300 python -m pytest tests/test_gpu_tc5_dense.py -q --timeout 200 -p no:cacheprovider > $OUT/tc5_dense.log 2>&1; say "tc5 dense tests rc=$? $(el)"
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > $OUT/gpu_suite.log 2>&1; say "gpu suite rc=$? $(el)"
XDR_SECTIONS=emcdr_map_step,dtcdr_both_step,conet_both_step timeout 600 python scripts/bench_new_kernels.py > $OUT/dense_models.log 2>&1; say "model steps (engine on) rc=$? $(el)"
timeout 300 python bench.py --workload emcdr_map --steps 20 --warmup 5 --map-engine "" --no-cpu-baseline > $OUT/bench_map_composed.json 2> $OUT/bench_map_composed.err; say "bench map composed+tc5 dense rc=$? $(el)"
XDR_SMALLRUN=1 XDR_SECTIONS=conet_both_step timeout 400 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file $OUT/conet_launches.csv \
  python scripts/bench_new_kernels.py > $OUT/conet_ncu.log 2>&1; say "conet ncu launch list rc=$? $(el)"
tail -5 $OUT/tc5_dense.log; tail -4 $OUT/gpu_suite.log
grep -h "composed\|tcgen05" $OUT/dense_models.log | cut -c1-230
cat $OUT/summary.txt
