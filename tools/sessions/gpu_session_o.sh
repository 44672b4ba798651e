#!/bin/bash
mkdir -p gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
for m in ours jodie; do
  run 170 $((29610 + RANDOM % 80)) tools/check_sharded.py $m > gpurun_out/o_check_$m.log 2>&1; echo "rc=$?" >> gpurun_out/o_check_$m.log
  grep -a "single GPU\|Error\|rc=\|File \"/tmp/code" gpurun_out/o_check_$m.log | tail -6
done
export PFO_HANG_DUMP_S=150
run 180 $((29700 + RANDOM % 80)) bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --large-bs 0 --eval-steps 4 > gpurun_out/o_bench_2gpu.json 2> gpurun_out/o_bench_2gpu.err
tail -c 200 gpurun_out/o_bench_2gpu.err
python - <<'PY'
import json
try:
    b=json.loads(open('gpurun_out/o_bench_2gpu.json').read().strip().split('\n')[-1])
    print('x2', b['value'], b['ms_per_step'], b['e2e']['value'], b['eval_users_per_sec'], b['config'].get('exchange_transport'))
except Exception as e: print('no line', e)
PY
