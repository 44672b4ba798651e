set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile --eval-steps 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r1b_launches.csv $B > gpurun_out/r1b_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'linear_kernel|wgrad_kernel|attn_nbr|neighbor_recent|mv_select|store_messages|persist_rank|gather_state' --launch-skip 120 --launch-count 45 -o gpurun_out/r1b_full $B > gpurun_out/r1b_ncu_full.log 2>&1
ls -la gpurun_out/
