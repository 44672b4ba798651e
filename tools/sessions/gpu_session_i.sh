#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_kernels.py tests/test_gpu_model.py -x -q > gpurun_out/i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/i_pytest.log; tail -5 gpurun_out/i_pytest.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_step.py > gpurun_out/i_sanitizer_memcheck.log 2>&1
echo "rc=$?" >> gpurun_out/i_sanitizer_memcheck.log; tail -6 gpurun_out/i_sanitizer_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_step.py > gpurun_out/i_sanitizer_racecheck.log 2>&1
echo "rc=$?" >> gpurun_out/i_sanitizer_racecheck.log; tail -6 gpurun_out/i_sanitizer_racecheck.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_one_step.csv python bench.py --ncu-step --warmup 3 > gpurun_out/i_ncu_launch.log 2>&1
tail -2 gpurun_out/i_ncu_launch.log; wc -l gpurun_out/r2_launches_one_step.csv
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/r2_full python bench.py --ncu-step --warmup 3 > gpurun_out/i_ncu_full.log 2>&1
tail -2 gpurun_out/i_ncu_full.log; ls -la gpurun_out/r2_full.ncu-rep
timeout 300 python bench.py --workload tgat --layers 2 --neighbors 20 --bs 8192 --steps 10 --warmup 5 --no-cpu-baseline --large-bs 0 --eval-steps 0 > gpurun_out/i_bench_tgat_config3.json 2> gpurun_out/i_bench_tgat_config3.err
tail -c 300 gpurun_out/i_bench_tgat_config3.err
python - <<'PY'
import json
try:
    b=json.loads(open('gpurun_out/i_bench_tgat_config3.json').read().strip().split('\n')[-1])
    print('tgat config3 bs8192', b['value'], b['ms_per_step'], b['config']['cuda_graph'])
except Exception as e: print('no line', e)
PY
