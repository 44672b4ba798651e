#!/bin/bash
# First GPU session of round 2: parity suite, unchanged main.py on the overlay (all model names), both bench arms.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
python -c "import os;print('cpus',os.cpu_count())" >> gpurun_out/a_smi.txt
rm -f gpurun_out/side_by_side.log
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
# unchanged main.py: every model name with --test_run (overlay and the reference's own torch-CUDA path), then one full
# config-1 epoch of `ours` on the overlay
timeout 900 python tools/run_reference_main.py --models ours,tgn,jodie,dyrep,tgat --stock --log gpurun_out/main_logs \
    > gpurun_out/a_main_test_run.log 2>&1
tail -12 gpurun_out/a_main_test_run.log
timeout 900 python tools/run_reference_main.py --models ours --log gpurun_out/main_logs_epoch -- --bs 128 --epoch 1 --drop_out 0.1 \
    > gpurun_out/a_main_full_epoch.log 2>&1
tail -3 gpurun_out/a_main_full_epoch.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/a_bench_reference.json 2> gpurun_out/a_bench_reference.err
tail -c 600 gpurun_out/a_bench_reference.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
tail -c 1500 gpurun_out/a_bench.json
