#!/bin/bash
# Final 1-GPU validation: whole GPU suite, smoke, ncu launch list + full capture (exported to CSV on the box), both bench arms.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/l_pytest.log; tail -4 gpurun_out/l_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/l_smoke.log 2>&1; tail -2 gpurun_out/l_smoke.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_one_step.csv python bench.py --ncu-step --warmup 3 > gpurun_out/l_ncu_launch.log 2>&1
wc -l gpurun_out/r2_launches_one_step.csv
timeout 700 ncu --profile-from-start off --set full --clock-control none -f -o /tmp/r2_full python bench.py --ncu-step --warmup 3 > gpurun_out/l_ncu_full.log 2>&1
tail -2 gpurun_out/l_ncu_full.log
ncu -i /tmp/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv 2> gpurun_out/l_ncu_export.err; ls -la /tmp/r2_full.ncu-rep gpurun_out/r2_full_raw.csv
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/l_bench_reference.json 2> gpurun_out/l_bench_reference.err
tail -c 300 gpurun_out/l_bench_reference.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/l_bench.json').read().strip().split('\n')[-1])
print(b['value'], b['ms_per_step'], b['e2e'], b['cpu_baseline'], b['eval']['value'], b['eval']['cpu_baseline'], b['roofline']['frac'])
for k,v in b['kernels'].items(): print(k, round(v['ms_per_step']*1e3,1),'us', (b['rooflines'].get(k) or {}).get('frac'))
PY
