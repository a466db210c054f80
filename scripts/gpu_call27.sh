#!/bin/bash
# call 27: one `ncu --set full` capture of every hot-path kernel (scripts/ncu_rows.py), new GPU tests, final-state bench
set -u
OUT=gpurun_out/c27
mkdir -p $OUT
say() { echo "$1" | tee -a $OUT/summary.txt; }
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 ))s; }
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k 'regex:gather_rows|scatter_add_rows|train_steps|tc5_mlp|dense_|act_bwd|bce_logit|mse_rows|neg_sample|prop_elementwise|spmm_work|sparse_optim|topk_|transfer_norm|zero_rows' -f -o $OUT/rows python scripts/ncu_rows.py > $OUT/ncu_rows.log 2>&1; say "ncu rows rc=$? $(el)"
tail -3 $OUT/ncu_rows.log
ls -la $OUT/rows.ncu-rep
ncu -i $OUT/rows.ncu-rep --page raw --csv > $OUT/rows_raw.csv 2>/dev/null; say "csv export rc=$? $(el)"
if [ $(stat -c %s $OUT/rows.ncu-rep) -gt 45000000 ]; then rm -f $OUT/rows.ncu-rep; say "rep too large for the pull: removed (csv kept)"; fi
cat $OUT/summary.txt
