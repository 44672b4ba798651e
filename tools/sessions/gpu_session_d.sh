#!/bin/bash
# 1-GPU session: programmatic dependent launch (parity + speed), 1-rank sharded machinery, procedural stream, K1 sweep.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_trainer.py -x -q > gpurun_out/d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/d_pytest.log; tail -4 gpurun_out/d_pytest.log
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q -s > gpurun_out/d_pytest_sharded.log 2>&1
echo "pytest rc=$?" >> gpurun_out/d_pytest_sharded.log; tail -12 gpurun_out/d_pytest_sharded.log
for pdl in 1 0; do
  PFO_PDL=$pdl timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-profile --eval-steps 0 > gpurun_out/d_bench_pdl$pdl.json 2> gpurun_out/d_bench_pdl$pdl.err
  python - <<PY
import json
b=json.loads(open('gpurun_out/d_bench_pdl$pdl.json').read().strip().split('\n')[-1])
print('PDL=$pdl', b['value'], b['ms_per_step'], b['e2e']['value'])
PY
done
timeout 200 python tools/k1_sweep.py > gpurun_out/d_k1_sweep.txt 2>&1; cat gpurun_out/d_k1_sweep.txt | tail -12
timeout 300 python bench.py --procedural --users 1000000 --items 5000 --events 50000000 --steps 10 --warmup 3 --eval-steps 2 > gpurun_out/d_bench_procedural_1gpu.json 2> gpurun_out/d_bench_procedural_1gpu.err
tail -c 300 gpurun_out/d_bench_procedural_1gpu.err
python - <<'PY'
import json
try:
    b=json.loads(open('gpurun_out/d_bench_procedural_1gpu.json').read().strip().split('\n')[-1])
    print('procedural 1gpu', b['value'], b['ms_per_step'], b['e2e'], b['eval_users_per_sec'], b['config']['parallelism'])
except Exception as e: print('no line', e)
PY
