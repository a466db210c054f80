#!/bin/bash
set -u
OUT=gpurun_out/c7
mkdir -p $OUT
LIBDIR=$PWD/recbole-cdr_b200/recbole_cdr_b200/lib
XDR_LIB=$LIBDIR/libxdr_vT.so timeout 200 python scripts/trace_steps.py fillerdbg > $OUT/trace_fillerdbg.txt 2>&1; echo "trace rc=$?"
tail -3 $OUT/trace_fillerdbg.txt
XDR_SECTIONS=emcdr_map_step XDR_BENCH_TC5=1 timeout 300 python scripts/bench_new_kernels.py > $OUT/map_step.log 2>&1; echo "map rc=$?"
grep tcgen05 $OUT/map_step.log | cut -c1-260
timeout 300 python -m pytest tests/test_gpu_engines.py -q --timeout 300 -p no:cacheprovider -k "tc5" > $OUT/tc5_tests.log 2>&1; echo "tc5 tests rc=$?"
tail -3 $OUT/tc5_tests.log
