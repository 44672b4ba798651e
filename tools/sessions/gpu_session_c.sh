#!/bin/bash
# 2-GPU session: node-sharded parity (eager, then graph), then short bench lines sharded vs replicated.
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29611 tools/check_sharded.py ours --no-graph > gpurun_out/c_check_ours_eager.log 2>&1; echo "rc=$?" >> gpurun_out/c_check_ours_eager.log
tail -4 gpurun_out/c_check_ours_eager.log
run 29612 tools/check_sharded.py ours > gpurun_out/c_check_ours.log 2>&1; echo "rc=$?" >> gpurun_out/c_check_ours.log
tail -4 gpurun_out/c_check_ours.log
for m in tgn jodie tgat; do
  run 29613 tools/check_sharded.py $m > gpurun_out/c_check_$m.log 2>&1; echo "rc=$?" >> gpurun_out/c_check_$m.log
  tail -3 gpurun_out/c_check_$m.log
done
run 29614 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --large-bs 0 > gpurun_out/c_bench_2gpu_sharded.json 2> gpurun_out/c_bench_2gpu_sharded.err
tail -c 300 gpurun_out/c_bench_2gpu_sharded.err
python - <<'PY'
import json
try:
    b=json.loads(open('gpurun_out/c_bench_2gpu_sharded.json').read().strip().split('\n')[-1])
    print('sharded x2', b['value'], b['ms_per_step'], b['e2e'], b['eval_users_per_sec'], b['config']['parallelism'], b['config']['cuda_graph'])
except Exception as e: print('no sharded line', e)
PY
run 29615 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --large-bs 0 --parallelism replicated > gpurun_out/c_bench_2gpu_replicated.json 2> gpurun_out/c_bench_2gpu_replicated.err
python - <<'PY'
import json
try:
    b=json.loads(open('gpurun_out/c_bench_2gpu_replicated.json').read().strip().split('\n')[-1])
    print('replicated x2', b['value'], b['ms_per_step'], b['e2e'], b['eval_users_per_sec'])
except Exception as e: print('no replicated line', e)
PY
