#!/bin/bash
# round-2 session Y: cosine-only form in the neighbour forward kernel, hoisted gate / accumulate loads in the linear epilogue
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/y_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/y_pytest.log; tail -3 gpurun_out/y_pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err
tail -c 300 gpurun_out/y_bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/y_bench.json').read().strip().split('\n')[-1])
print(b['value'], b['ms_per_step'], b['e2e']['value'], b['eval']['value'], b['roofline']['frac'], b['gpu_launches'])
for k,v in b['kernels'].items(): print(k, round(v['ms_per_step']*1e3,1),'us', (b['rooflines'].get(k) or {}).get('frac'))
print('large', b['large_batch']['value'], b['large_batch']['ms_per_step'])
PY
