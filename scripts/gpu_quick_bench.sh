mkdir -p gpurun_out/final2
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final2/bench_k20.json 2> gpurun_out/final2/bench_k20.err; echo rc=$?
python - <<PY
import json
d = json.loads(open('gpurun_out/final2/bench_k20.json').read().strip().splitlines()[-1])
print('value %.4e us/step %.3f frac %.4f' % (d['value'], d['ms_per_step']*1e3, d['roofline']['frac']))
print(d['e2e'])
PY
