#!/bin/bash
set -u
OUT=gpurun_out/c4
mkdir -p $OUT
say() { echo "$1" | tee -a $OUT/summary.txt; }
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 ))s; }
python __graft_entry__.py > $OUT/build.log 2>&1; say "build rc=$? $(el)"
timeout 200 python scripts/trace_steps.py lazy > $OUT/trace_lazy.txt 2>&1; say "trace lazy rc=$? $(el)"
timeout 200 python scripts/trace_steps.py > $OUT/trace_plain.txt 2>&1; say "trace plain rc=$? $(el)"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extras > $OUT/bench_k20.json 2> $OUT/bench_k20.err; say "bench K=20 rc=$? $(el)"
XDR_SECTIONS=emcdr_map_step XDR_BENCH_TC5=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/map_launches.csv \
  python scripts/bench_new_kernels.py > $OUT/map_ncu.log 2>&1; say "map step ncu launch list rc=$? $(el)"
cat $OUT/trace_lazy.txt | tail -30
cat $OUT/summary.txt
