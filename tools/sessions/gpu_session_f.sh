#!/bin/bash
mkdir -p gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
for m in ours tgat; do
  run 170 $((29610 + RANDOM % 80)) tools/check_sharded.py $m > gpurun_out/f_check_$m.log 2>&1; echo "rc=$?" >> gpurun_out/f_check_$m.log
  grep -a "single GPU\|Error\|rc=\|File \"/tmp/code" gpurun_out/f_check_$m.log | tail -12
done
