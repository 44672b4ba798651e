#!/bin/bash
# round-2 session S: parity suite + single-GPU bench after the e2e / K5 / neighbour-kernel changes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s_pytest.log; tail -12 gpurun_out/s_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err
tail -c 400 gpurun_out/s_bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/s_bench.json').read().strip().split('\n')[-1])
print(b['value'], b['ms_per_step'], b['e2e'], b['eval']['value'], b['roofline']['frac'], b['gpu_launches'])
for k,v in b['kernels'].items(): print(k, round(v['ms_per_step']*1e3,1),'us', v['calls_per_step'], (b['rooflines'].get(k) or {}).get('frac'))
print('large', b['large_batch']['value'], b['large_batch']['ms_per_step'])
print('eval', json.dumps(b['eval'])[:600])
PY
