#!/bin/bash
# Round-2 GPU call 2: lazily zeroed gradient tables (parity + bench + ncu), tcgen05 map-step timing, CoNet hot-row diagnostic.
set -u
OUT=gpurun_out/c2
mkdir -p $OUT
say() { echo "$1" | tee -a $OUT/summary.txt; }
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 ))s; }
python __graft_entry__.py > $OUT/build.log 2>&1; say "build rc=$? $(el)"
timeout 600 python -m pytest tests/test_gpu_lazy_tables.py -q --timeout 300 -p no:cacheprovider > $OUT/lazy.log 2>&1; say "lazy tables rc=$? $(el)"
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > $OUT/gpu_suite.log 2>&1; say "gpu suite rc=$? $(el)"
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_k20.json 2> $OUT/bench_k20.err; say "bench K=20 rc=$? $(el)"
timeout 400 python bench.py --steps 200 --warmup 5 --repeats 5 --no-cpu-baseline > $OUT/bench_k200.json 2> $OUT/bench_k200.err; say "bench K=200 rc=$? $(el)"
timeout 300 python bench.py --steps 20 --warmup 5 --coop 0 --no-extras --no-cpu-baseline --no-e2e > $OUT/bench_k20_plain.json 2> $OUT/bench_k20_plain.err; say "bench K=20 plain rc=$? $(el)"
B=16384 timeout 600 python scripts/diag_conet_hot.py > $OUT/diag_conet.json 2> $OUT/diag_conet.err; say "conet diag rc=$? $(el)"
XDR_SECTIONS=emcdr_map_step XDR_BENCH_TC5=1 timeout 300 python scripts/bench_new_kernels.py > $OUT/map_step.log 2>&1; say "map step engines rc=$? $(el)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 20 --warmup 5 --repeats 3 --no-extras --no-cpu-baseline --no-e2e > /dev/null 2>&1; say "ncu launch list rc=$? $(el)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:train_steps_staged -s 1 -c 1 -o $OUT/staged_fresh_k20 \
  python bench.py --steps 20 --warmup 5 --repeats 2 --no-extras --no-cpu-baseline --no-e2e > /dev/null 2>&1; say "ncu full (fresh) rc=$? $(el)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:train_steps_staged -s 1 -c 1 -o $OUT/staged_fresh_k200 \
  python bench.py --steps 200 --warmup 5 --repeats 2 --no-extras --no-cpu-baseline --no-e2e > /dev/null 2>&1; say "ncu full (fresh, K=200) rc=$? $(el)"
tail -n 6 $OUT/*.log | tail -n 80
cat $OUT/summary.txt
