#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r_pytest.log; tail -3 gpurun_out/r_pytest.log
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-profile --eval-steps 0 > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
tail -c 200 gpurun_out/r_bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r_bench.json').read().strip().split('\n')[-1])
print(b['value'], b['ms_per_step'], b['e2e'], b['gpu_launches'])
PY
timeout 300 python bench.py --sharded-1gpu --steps 20 --warmup 5 --no-cpu-baseline --no-profile --large-bs 0 --eval-steps 0 > gpurun_out/r_bench_sharded_1rank.json 2> gpurun_out/r_bench_sharded_1rank.err
python - <<'PY'
import json
try:
    b=json.loads(open('gpurun_out/r_bench_sharded_1rank.json').read().strip().split('\n')[-1])
    print('sharded 1 rank', b['value'], b['ms_per_step'])
except Exception as e: print('no line', e)
PY
