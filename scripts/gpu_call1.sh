#!/bin/bash
# Round-2 GPU call 1: validate on hardware everything written without a GPU, then the regular suite and the bench.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_call1.sh'
set -u
mkdir -p gpurun_out
OUT=gpurun_out/c1
mkdir -p $OUT
say() { echo "$1" | tee -a $OUT/summary.txt; }
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 ))s; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/smi.txt 2>&1
python __graft_entry__.py > $OUT/build.log 2>&1; say "build rc=$? $(el)"

# 1. tcgen05 descriptor experiment (standalone, one CTA per variant, each variant in a child process)
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -lineinfo -o /tmp/ubench_tcgen05 scripts/ubench_tcgen05.cu > $OUT/tcgen05.log 2>&1 \
  && timeout 300 /tmp/ubench_tcgen05 >> $OUT/tcgen05.log 2>&1
say "ubench_tcgen05 rc=$? $(el)"

# 2. library self-test GEMMs on tcgen05, one process per case (a faulting descriptor is a sticky error)
: > $OUT/tc5_selftest.log
fails=0
for id in "test_tc5_selftest_gemm_all_majors[0-0]" "test_tc5_selftest_gemm_all_majors[0-1]" "test_tc5_selftest_gemm_all_majors[1-0]" \
          "test_tc5_selftest_gemm_all_majors[1-1]" "test_tc5_selftest_gemm_bf16x3_all_majors[0-0]" \
          "test_tc5_selftest_gemm_bf16x3_all_majors[0-1]" "test_tc5_selftest_gemm_bf16x3_all_majors[1-0]" \
          "test_tc5_selftest_gemm_bf16x3_all_majors[1-1]" "test_tc5_selftest_gemm_bf16x3_all_majors[2-0]" \
          "test_tc5_selftest_gemm_bf16x3_all_majors[2-1]"; do
  XDR_RUN_UNVALIDATED=1 timeout 120 python -m pytest "tests/test_gpu_engines.py::$id" -q --timeout 60 >> $OUT/tc5_selftest.log 2>&1
  rc=$?; echo "   $id rc=$rc" | tee -a $OUT/summary.txt; [ $rc -ne 0 ] && fails=$((fails+1))
done
say "tc5 self-test failing cases: $fails $(el)"

# 3. every other unvalidated test (no -x: the full list of failures is the result), tc5 kernels in their own process
XDR_RUN_UNVALIDATED=1 timeout 900 python -m pytest tests/test_gpu_engines.py -q --timeout 120 -k "not tc5" -p no:cacheprovider \
  > $OUT/unvalidated.log 2>&1
say "unvalidated (not tc5) rc=$? $(el)"
XDR_RUN_UNVALIDATED=1 timeout 400 python -m pytest tests/test_gpu_engines.py -q --timeout 120 -k "tc5 and not selftest" -p no:cacheprovider \
  > $OUT/unvalidated_tc5.log 2>&1
say "unvalidated (tc5) rc=$? $(el)"

# 4. the regular gpu suite
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > $OUT/gpu_suite.log 2>&1
say "gpu suite rc=$? $(el)"
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; say "smoke rc=$? $(el)"

# 5. bench: driver protocol (K=20), K=200, plain launches
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_k20.json 2> $OUT/bench_k20.err; say "bench K=20 rc=$? $(el)"
timeout 400 python bench.py --steps 200 --warmup 5 --repeats 5 --no-extras --no-cpu-baseline > $OUT/bench_k200.json 2> $OUT/bench_k200.err; say "bench K=200 rc=$? $(el)"
timeout 300 python bench.py --steps 20 --warmup 5 --coop 0 --no-extras --no-cpu-baseline --no-e2e > $OUT/bench_k20_plain.json 2> $OUT/bench_k20_plain.err; say "bench K=20 plain launch rc=$? $(el)"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; say "bench reference rc=$? $(el)"

# 6. new-kernel timings next to the paths they replace
timeout 600 python scripts/bench_new_kernels.py > $OUT/new_kernels.log 2>&1; say "bench_new_kernels rc=$? $(el)"

# 7. ncu: launch list of the bench command, one full capture of the persistent kernel
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 20 --warmup 5 --repeats 3 --no-extras --no-cpu-baseline --no-e2e > /dev/null 2>&1; say "ncu launch list rc=$? $(el)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:train_steps_staged -s 1 -c 1 -o $OUT/staged_k20 \
  python bench.py --steps 20 --warmup 5 --repeats 2 --no-extras --no-cpu-baseline --no-e2e > /dev/null 2>&1; say "ncu full rc=$? $(el)"
tail -n 5 $OUT/*.log | tail -n 120
cat $OUT/summary.txt
