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
# 5. the tcgen05 descriptor experiment (which operand layouts / descriptor readings the hardware accepts)
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -lineinfo -o /tmp/ubench_tcgen05 scripts/ubench_tcgen05.cu > gpurun_out/tcgen05.log 2>&1 \
    && timeout 120 /tmp/ubench_tcgen05 >> gpurun_out/tcgen05.log 2>&1
echo "ubench_tcgen05 rc=$?" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/*.log
cat gpurun_out/summary.txt
