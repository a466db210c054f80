#!/bin/bash
# Multi-GPU call:  gpurun --gpus N -- 'bash scripts/gpu_multi.sh'   (N = 2, 4 or 8)
set -u
N=$(python -c "import torch; print(torch.cuda.device_count())")
OUT=gpurun_out/multi$N
mkdir -p $OUT
say() { echo "$1" | tee -a $OUT/summary.txt; }
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 ))s; }
nvidia-smi topo -m > $OUT/topo.txt 2>&1
if [ "${SKIP_TESTS:-0}" != 1 ]; then
  XDR_RUN_UNVALIDATED=1 timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 300 -p no:cacheprovider > $OUT/multi_tests.log 2>&1; say "multi-GPU tests on $N GPUs rc=$? $(el)"
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_k20.json 2> $OUT/bench_k20.err; say "bench N=$N K=20 rc=$? $(el)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 \
  bench.py --gpus $N --steps 200 --warmup 5 --repeats 5 --no-e2e > $OUT/bench_k200.json 2> $OUT/bench_k200.err; say "bench N=$N K=200 rc=$? $(el)"
if [ "${SKIP_REF:-0}" != 1 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --impl reference --gpus $N --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; say "reference arm N=$N rc=$? $(el)"
fi
if [ "${SKIP_A2A:-0}" != 1 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
    scripts/bench_a2a.py > $OUT/a2a_bench.log 2>&1; say "bench_a2a rc=$? $(el)"
fi
if [ "${BITGCF_SCALE:-0}" != 0 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 \
    scripts/bench_bitgcf.py --scale $BITGCF_SCALE > $OUT/bitgcf.log 2>&1; say "bitgcf scale $BITGCF_SCALE on $N GPUs rc=$? $(el)"
fi
tail -n 3 $OUT/*.log $OUT/*.err | tail -n 60
python - <<PY
import json
for f in ('bench_k20', 'bench_k200'):
    try:
        d = json.loads(open('$OUT/' + f + '.json').read().strip().splitlines()[-1])
        print(f, 'N=%d value %.3e us/step %.2f frac/GPU %.3f e2e %s parity %s' % (d['n_gpus'], d['value'], d['ms_per_step'] * 1e3, d['roofline']['frac'],
              d['e2e'] and '%.3e' % d['e2e']['value'], d.get('sharded_parity', {}).get('all_ranks_ok')))
    except Exception as e:
        print(f, 'ERR', e)
PY
cat $OUT/summary.txt
