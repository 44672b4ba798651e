#!/bin/bash
# round-2 session T: fused wgrad slab reduction (grid barrier), state update on the side stream, shared query rows in eval
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/t_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t_pytest.log; tail -8 gpurun_out/t_pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err
tail -c 400 gpurun_out/t_bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/t_bench.json').read().strip().split('\n')[-1])
print(b['value'], b['ms_per_step'], b['e2e']['value'], b['eval']['value'], b['roofline']['frac'], b['gpu_launches'])
for k,v in b['kernels'].items(): print(k, round(v['ms_per_step']*1e3,1),'us', v['calls_per_step'], (b['rooflines'].get(k) or {}).get('frac'))
print('large', b['large_batch']['value'], b['large_batch']['ms_per_step'])
print('eval', json.dumps(b['eval'])[:500])
PY
for v in "PFO_WGRAD_FUSED=0" "PFO_STORE_OVERLAP=0" "PFO_WGRAD_FUSED=0 PFO_STORE_OVERLAP=0"; do
  env $v timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-profile --large-bs 0 --eval-steps 0 > gpurun_out/t_var.json 2> gpurun_out/t_var.err
  python - "$v" <<'PY'
import json, sys
try:
    b=json.loads(open('gpurun_out/t_var.json').read().strip().split('\n')[-1])
    print(sys.argv[1], b['value'], b['ms_per_step'], b['e2e']['value'], b['gpu_launches'])
except Exception as e: print(sys.argv[1], 'no line', e)
PY
done
