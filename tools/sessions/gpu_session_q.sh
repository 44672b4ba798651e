#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/q_pytest.log; tail -3 gpurun_out/q_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
tail -c 300 gpurun_out/q_bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/q_bench.json').read().strip().split('\n')[-1])
print(b['value'], b['ms_per_step'], b['e2e'], b['cpu_baseline']['value'], b['eval']['value'], b['eval']['cpu_baseline'], b['roofline']['frac'], b['roofline']['traffic'])
for k,v in b['kernels'].items(): print(k, round(v['ms_per_step']*1e3,1),'us', v['calls_per_step'], (b['rooflines'].get(k) or {}).get('frac'))
PY
