#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_trainer.py -x -q > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -5 gpurun_out/b_pytest.log
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -x -q > gpurun_out/b_pytest2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_pytest2.log
tail -3 gpurun_out/b_pytest2.log
nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/cos_probe.cu -o /tmp/cos_probe && /tmp/cos_probe > gpurun_out/b_cos_probe.txt 2>&1
cat gpurun_out/b_cos_probe.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/b_bench.json').read().strip().split('\n')[-1])
print(b['value'], b['ms_per_step'], b['e2e'])
for k,v in b['kernels'].items(): print(k, round(v['ms_per_step']*1e3,1),'us', (b['rooflines'].get(k) or {}).get('frac'))
print('large', b['large_batch']['value'], {k:(round(v['ms_per_step']*1e3,1), round(v['frac'],3)) for k,v in b['large_batch']['rooflines'].items()})
print('eval', b['eval']['value'], b['eval']['roofline'])
PY
timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --workload tgat --layers 2 --neighbors 20 --bs 2048 --large-bs 0 > gpurun_out/b_bench_tgat.json 2> gpurun_out/b_bench_tgat.err
tail -c 400 gpurun_out/b_bench_tgat.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/b_bench_tgat.json').read().strip().split('\n')[-1])
print('tgat', b['value'], b['ms_per_step'], b['config']['cuda_graph'])
PY
