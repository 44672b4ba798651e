#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/m_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/m_pytest.log; tail -6 gpurun_out/m_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/m_smoke.log 2>&1; tail -2 gpurun_out/m_smoke.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err
tail -c 300 gpurun_out/m_bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/m_bench.json').read().strip().split('\n')[-1])
print(b['value'], b['ms_per_step'], b['e2e'], b['eval']['value'], b['roofline']['frac'], b['gpu_launches'])
for k,v in b['kernels'].items(): print(k, round(v['ms_per_step']*1e3,1),'us', v['calls_per_step'], (b['rooflines'].get(k) or {}).get('frac'))
print('large', b['large_batch']['value'], b['large_batch']['ms_per_step'])
PY
timeout 300 python bench.py --sharded-1gpu --steps 20 --warmup 5 --no-cpu-baseline --no-profile --large-bs 0 --eval-steps 0 > gpurun_out/m_bench_sharded_1rank.json 2> gpurun_out/m_bench_sharded_1rank.err
python - <<'PY'
import json
try:
    b=json.loads(open('gpurun_out/m_bench_sharded_1rank.json').read().strip().split('\n')[-1])
    print('sharded 1 rank', b['value'], b['ms_per_step'])
except Exception as e: print('no line', e)
PY
