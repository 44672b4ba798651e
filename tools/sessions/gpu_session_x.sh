#!/bin/bash
# round-2 session X: final 1-GPU validation -- grid variants, GPU suite, smoke, ncu launch list + full capture of the main kernels, full bench line
mkdir -p gpurun_out
for v in "PFO_NONE=0" "PFO_ATTN_FWD_CTAS=18" "PFO_ATTN_BWD_CTAS=8"; do
  env $v timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-profile --large-bs 0 --eval-steps 0 > gpurun_out/x_var.json 2> gpurun_out/x_var.err
  python - "$v" <<'PY'
import json, sys
try:
    b=json.loads(open('gpurun_out/x_var.json').read().strip().split('\n')[-1])
    print(sys.argv[1], round(b['value']), round(b['ms_per_step'],4), round(b['e2e']['value']))
except Exception as e: print(sys.argv[1], 'no line', e)
PY
done
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/x_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/x_pytest.log; tail -3 gpurun_out/x_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/x_smoke.log 2>&1; tail -2 gpurun_out/x_smoke.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_one_step.csv python bench.py --ncu-step --warmup 3 > gpurun_out/x_ncu_launch.log 2>&1
wc -l gpurun_out/r2f_launches_one_step.csv
timeout 420 ncu --profile-from-start off --set full --clock-control none --kernel-name 'regex:linear_tma_kernel|wgrad_tma|attn_nbr|mv_select|neighbor_recent|store_messages|bpr_kernel|cell_|gather_state' -f -o /tmp/r2f_full python bench.py --ncu-step --warmup 3 > gpurun_out/x_ncu_full.log 2>&1
tail -2 gpurun_out/x_ncu_full.log
ncu -i /tmp/r2f_full.ncu-rep --page raw --csv > gpurun_out/r2f_full_raw.csv 2> gpurun_out/x_ncu_export.err; ls -la /tmp/r2f_full.ncu-rep gpurun_out/r2f_full_raw.csv
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/x_bench.json').read().strip().split('\n')[-1])
print(b['value'], b['ms_per_step'], b['e2e'], b['cpu_baseline'], b['eval']['value'], b['eval']['cpu_baseline'], b['roofline']['frac'])
for k,v in b['kernels'].items(): print(k, round(v['ms_per_step']*1e3,1),'us', (b['rooflines'].get(k) or {}).get('frac'))
PY
