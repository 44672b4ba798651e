#!/bin/bash
# round-2 session Z: ncu launch list of one step + full capture of the two kernels changed after session X
mkdir -p gpurun_out
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches_one_step.csv python bench.py --ncu-step --warmup 3 > gpurun_out/z_ncu_launch.log 2>&1
wc -l gpurun_out/r2g_launches_one_step.csv
timeout 200 ncu --profile-from-start off --set full --clock-control none --kernel-name 'regex:linear_tma_kernel|attn_nbr_fwd' -f -o /tmp/r2g_full python bench.py --ncu-step --warmup 3 > gpurun_out/z_ncu_full.log 2>&1
tail -1 gpurun_out/z_ncu_full.log
ncu -i /tmp/r2g_full.ncu-rep --page raw --csv > gpurun_out/r2g_full_raw.csv 2> gpurun_out/z_ncu_export.err; ls -la gpurun_out/r2g_full_raw.csv
