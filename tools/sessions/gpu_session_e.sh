#!/bin/bash
# 2-GPU session with strict timeouts: node-sharded parity (hang dump after 100 s), then short bench lines.
mkdir -p gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
for m in tgat ours tgn jodie; do
  run 170 $((29610 + RANDOM % 80)) tools/check_sharded.py $m > gpurun_out/e_check_$m.log 2>&1; echo "rc=$?" >> gpurun_out/e_check_$m.log
  grep -a "single GPU\|Error\|error\|rc=\|step " gpurun_out/e_check_$m.log | tail -12
done
run 170 $((29700 + RANDOM % 80)) tools/check_sharded.py ours --no-graph > gpurun_out/e_check_ours_eager.log 2>&1; echo "rc=$?" >> gpurun_out/e_check_ours_eager.log
grep -a "single GPU\|Error\|rc=" gpurun_out/e_check_ours_eager.log | tail -5
run 200 29631 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --large-bs 0 > gpurun_out/e_bench_2gpu_sharded.json 2> gpurun_out/e_bench_2gpu_sharded.err
tail -c 300 gpurun_out/e_bench_2gpu_sharded.err
python - <<'PY'
import json
try:
    b=json.loads(open('gpurun_out/e_bench_2gpu_sharded.json').read().strip().split('\n')[-1])
    print('sharded x2', b['value'], b['ms_per_step'], b['e2e'], b['eval_users_per_sec'], b['config']['parallelism'], b['config']['cuda_graph'])
except Exception as e: print('no sharded line', e)
PY
