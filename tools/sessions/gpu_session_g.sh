#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --procedural --users 1000000 --items 5000 --events 50000000 --steps 10 --warmup 3 --eval-steps 2 > gpurun_out/g_bench_procedural_1gpu.json 2> gpurun_out/g_bench_procedural_1gpu.err
tail -c 400 gpurun_out/g_bench_procedural_1gpu.err
python - <<'PY'
import json
try:
    b=json.loads(open('gpurun_out/g_bench_procedural_1gpu.json').read().strip().split('\n')[-1])
    print('procedural 1gpu', b['value'], b['ms_per_step'], b['e2e'], b['eval_users_per_sec'], b['config']['parallelism'])
except Exception as e: print('no line', e)
PY
timeout 300 python bench.py --sharded-1gpu --steps 20 --warmup 5 --no-cpu-baseline --large-bs 0 --eval-steps 0 > gpurun_out/g_bench_sharded_1rank.json 2> gpurun_out/g_bench_sharded_1rank.err
tail -c 400 gpurun_out/g_bench_sharded_1rank.err
python - <<'PY'
import json
try:
    b=json.loads(open('gpurun_out/g_bench_sharded_1rank.json').read().strip().split('\n')[-1])
    print('sharded 1 rank', b['value'], b['ms_per_step'], b['gpu_launches'])
    for k,v in (b['kernels'] or {}).items(): print(k, round(v['ms_per_step']*1e3,1),'us', v['calls_per_step'])
except Exception as e: print('no line', e)
PY
