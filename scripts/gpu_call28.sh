#!/bin/bash
# call 28: the remaining `ncu --set full` captures (graph kernels, sampler, row-sparse optimizer, fused top-k)
set -u
OUT=gpurun_out/c28
mkdir -p $OUT
XDR_NCU_JOBS='graph,transfer,neg_sample,row-sparse,topk' timeout 420 ncu --set full --clock-control none --profile-from-start off \
  -k 'regex:neg_sample|prop_elementwise|spmm_work|sparse_optim|topk_|transfer_norm' -c 14 -f -o $OUT/rows2 python scripts/ncu_rows.py > $OUT/ncu_rows2.log 2>&1; echo "ncu rc=$?"
tail -3 $OUT/ncu_rows2.log
ncu -i $OUT/rows2.ncu-rep --page raw --csv > $OUT/rows2_raw.csv 2>/dev/null; echo "csv rc=$?"
ls -la $OUT
if [ $(stat -c %s $OUT/rows2.ncu-rep) -gt 45000000 ]; then rm -f $OUT/rows2.ncu-rep; fi
