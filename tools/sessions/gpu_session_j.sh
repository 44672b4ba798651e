#!/bin/bash
mkdir -p gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
for m in ours tgat; do
  run 170 $((29610 + RANDOM % 80)) tools/check_sharded.py $m > gpurun_out/j_check_$m.log 2>&1; echo "rc=$?" >> gpurun_out/j_check_$m.log
  grep -a "single GPU\|Error\|rc=\|File \"/tmp/code" gpurun_out/j_check_$m.log | tail -8
done
export PFO_HANG_DUMP_S=150
for tp in peer nccl; do
  PFO_TRANSPORT=$tp run 180 $((29700 + RANDOM % 80)) bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --large-bs 0 --eval-steps 4 > gpurun_out/j_bench_2gpu_$tp.json 2> gpurun_out/j_bench_2gpu_$tp.err
  tail -c 200 gpurun_out/j_bench_2gpu_$tp.err
  python - $tp <<'PY'
import json, sys
try:
    b=json.loads(open(f'gpurun_out/j_bench_2gpu_{sys.argv[1]}.json').read().strip().split('\n')[-1])
    print(sys.argv[1], b['value'], b['ms_per_step'], b['e2e']['value'], b['eval_users_per_sec'], b['config'].get('exchange_transport'))
except Exception as e: print(sys.argv[1], 'no line', e)
PY
done
