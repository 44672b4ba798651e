#!/bin/bash
# 8-GPU session: node-sharded bench at the default workload, BASELINE config 4 (10M x 5k x 1B, procedural), replicated for comparison.
mkdir -p gpurun_out
export PFO_HANG_DUMP_S=230
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
summ() { python - "$1" "$2" <<'PY'
import json, sys
try:
    b=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
    print(sys.argv[2], 'events/s', round(b['value']), 'ms/step', round(b['ms_per_step'],3), 'e2e', round(b['e2e']['value']) if b.get('e2e') else None, 'eval users/s', b.get('eval_users_per_sec'), b['config']['parallelism'], 'graph', b['config']['cuda_graph'])
except Exception as e: print(sys.argv[2], 'no line', e)
PY
}
run 250 29641 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --large-bs 0 > gpurun_out/h_bench_8gpu_sharded.json 2> gpurun_out/h_bench_8gpu_sharded.err
tail -c 300 gpurun_out/h_bench_8gpu_sharded.err; summ gpurun_out/h_bench_8gpu_sharded.json sharded_x8
export PFO_HANG_DUMP_S=400
run 430 29642 bench.py --gpus 8 --config4 --steps 20 --warmup 5 --eval-steps 4 --eval-bs 256 --no-cpu-baseline > gpurun_out/h_bench_8gpu_config4.json 2> gpurun_out/h_bench_8gpu_config4.err
tail -c 600 gpurun_out/h_bench_8gpu_config4.err; summ gpurun_out/h_bench_8gpu_config4.json config4_x8
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader > gpurun_out/h_smi_after.txt 2>&1
export PFO_HANG_DUMP_S=200
run 220 29643 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --large-bs 0 --parallelism replicated > gpurun_out/h_bench_8gpu_replicated.json 2> gpurun_out/h_bench_8gpu_replicated.err
summ gpurun_out/h_bench_8gpu_replicated.json replicated_x8
