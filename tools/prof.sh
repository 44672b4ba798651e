#!/bin/bash
# ncu passes of one training step on the bench workload (run on the GPU box through gpurun):
#   tools/prof.sh <tag> [extra bench args]
# 1. launch list (gpu__time_duration.sum of every kernel) of warm-up + one timed step -> gpurun_out/<tag>_launches.csv
# 2. --set full capture of this repo's hot kernels in one warm step                    -> gpurun_out/<tag>_full.ncu-rep
TAG=${1:-prof}; shift
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --eval-steps 0 --no-graph $*"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu_launch.log 2>&1
K='regex:linear_tma_kernel|wgrad_tma_kernel|attn_nbr|neighbor_|mv_select|bpr_kernel|store_messages|persist_rank|mark_nodes|gather_state|cell_|fold_'
ncu --set full --clock-control none --import-source on -k "$K" --launch-skip ${SKIP:-90} --launch-count ${COUNT:-32} -f -o gpurun_out/${TAG}_full $B > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out/ | grep ${TAG}
