#!/bin/bash
# First GPU call of the next session: validates on hardware what was written while no GPU was reachable.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/r2_gpu_session.sh'
# Everything is bounded by `timeout`; results land in gpurun_out/ (logs, jsonl rows, ncu launch list).
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
# 1. memory checker on the smallest case of each new kernel (a wild pointer must not take the box down later)
XDR_RUN_UNVALIDATED=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 \
    python -m pytest tests/test_gpu_unvalidated.py -x -q -k "map_loss_matches_oracle and 33 or conet_fused_matches_oracle and 63 or sparse_optim and sgd or full_sort_topk and 300" \
    > gpurun_out/sanitizer.log 2>&1
echo "sanitizer rc=$?" | tee -a gpurun_out/summary.txt
# 2. the hardware parity tests of the new kernels, then the regular gpu suite
XDR_RUN_UNVALIDATED=1 timeout 900 python -m pytest tests/test_gpu_unvalidated.py -q --timeout 300 > gpurun_out/unvalidated.log 2>&1
echo "unvalidated rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/gpu_suite.log 2>&1
echo "gpu suite rc=$?" | tee -a gpurun_out/summary.txt
# 3. timings next to the paths they replace
timeout 900 python scripts/bench_new_kernels.py > gpurun_out/new_kernels.log 2>&1
echo "bench_new_kernels rc=$?" | tee -a gpurun_out/summary.txt
# 4. launch list of one fused CoNet step + one full capture of the fused kernel
XDR_SMALL=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/new_kernels_launches.csv python scripts/bench_new_kernels.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conet_kernel -s 2 -c 1 \
    -o gpurun_out/tc_conet python scripts/bench_new_kernels.py > /dev/null 2>&1
# 4b. the alternative tile engines (tc_tile.cuh XDR_TC_MODE): bf16x3 (parity-grade) and one TF32 pass (diagnostic upper bound)
for m in 1 2; do
  XDR_EXTRA_NVCC_FLAGS="-DXDR_TC_MODE=$m" XDR_BUILD_DIR=build_m$m XDR_LIB_NAME=libxdr_m$m.so python recbole-cdr_b200/build.py > gpurun_out/build_m$m.log 2>&1
done
XDR_LIB=$PWD/recbole-cdr_b200/recbole_cdr_b200/lib/libxdr_m1.so XDR_RUN_UNVALIDATED=1 timeout 600 python -m pytest tests/test_gpu_unvalidated.py -q \
    -k "tc_mlp or conet_fused or tc_engine or full_sort_topk" --timeout 300 > gpurun_out/unvalidated_bf16x3.log 2>&1
echo "unvalidated (bf16x3 engine) rc=$?" | tee -a gpurun_out/summary.txt
for m in 1 2; do
  XDR_LIB=$PWD/recbole-cdr_b200/recbole_cdr_b200/lib/libxdr_m$m.so XDR_SECTIONS=emcdr_map_step,conet_both_step,full_sort_topk XDR_SKIP_CHECK=$((m-1)) \
      timeout 600 python scripts/bench_new_kernels.py > gpurun_out/new_kernels_m$m.log 2>&1
  echo "bench_new_kernels (XDR_TC_MODE=$m) rc=$?" | tee -a gpurun_out/summary.txt
done
# 5. the tcgen05 descriptor experiment (which operand layouts / descriptor readings the hardware accepts)
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -lineinfo -o /tmp/ubench_tcgen05 scripts/ubench_tcgen05.cu > gpurun_out/tcgen05.log 2>&1 \
    && timeout 120 /tmp/ubench_tcgen05 >> gpurun_out/tcgen05.log 2>&1
echo "ubench_tcgen05 rc=$?" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/*.log
cat gpurun_out/summary.txt
