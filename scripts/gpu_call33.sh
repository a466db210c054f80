#!/bin/bash
# call 33: ncu --set full (with source) of the current tc5_mlp_kernel inside the eager map step
set -u
OUT=gpurun_out/c33
mkdir -p $OUT
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc5_mlp -s 6 -c 1 -f -o $OUT/tc5_mlp_eager \
  python bench.py --workload emcdr_map --steps 3 --warmup 3 --repeats 1 --no-cpu-baseline --no-e2e > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
ls -la $OUT
