#!/bin/bash
# Round-2 GPU call 3: lazy tables with batched claims + bulk zero stores: parity, timing, accumulation-noise diagnostic.
set -u
OUT=gpurun_out/c3
mkdir -p $OUT
say() { echo "$1" | tee -a $OUT/summary.txt; }
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 ))s; }
python __graft_entry__.py > $OUT/build.log 2>&1; say "build rc=$? $(el)"
timeout 600 python -m pytest tests/test_gpu_lazy_tables.py tests/test_gpu_steps.py tests/test_gpu_z_fullsize.py -q --timeout 300 -p no:cacheprovider > $OUT/lazy.log 2>&1; say "lazy + steps + fullsize rc=$? $(el)"
timeout 300 python scripts/diag_accum_noise.py > $OUT/accum_noise.json 2> $OUT/accum_noise.err; say "accum noise diag rc=$? $(el)"
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/bench_k20.json 2> $OUT/bench_k20.err; say "bench K=20 rc=$? $(el)"
timeout 400 python bench.py --steps 200 --warmup 5 --repeats 5 --no-cpu-baseline --no-e2e --no-extras > $OUT/bench_k200.json 2> $OUT/bench_k200.err; say "bench K=200 rc=$? $(el)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:train_steps_staged -s 1 -c 1 -o $OUT/staged_fresh_k20 \
  python bench.py --steps 20 --warmup 5 --repeats 2 --no-extras --no-cpu-baseline --no-e2e > /dev/null 2>&1; say "ncu full (fresh) rc=$? $(el)"
tail -n 6 $OUT/*.log | tail -n 40
cat $OUT/summary.txt
