#!/bin/bash
# round-2 session W: grid of the neighbour forward kernel (one resident wave vs the guessed 10 CTAs per SM), suite with the new defaults
mkdir -p gpurun_out
for w in 0 5 10 12; do
  PFO_ATTN_FWD_CTAS=$w timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --large-bs 0 --eval-steps 0 > gpurun_out/w_c$w.json 2> gpurun_out/w_c$w.err
  python - $w <<'PY'
import json, sys
try:
    b=json.loads(open('gpurun_out/w_c%s.json' % sys.argv[1]).read().strip().split('\n')[-1])
    k=b['kernels']
    print('fwd ctas/SM', sys.argv[1], round(b['value']), round(b['ms_per_step'],4), 'fwd', round(k['pfo_attn_nbr_fwd']['ms_per_step']*1e3,1), 'bwd', round(k['pfo_attn_nbr_bwd']['ms_per_step']*1e3,1), 'bpr', round(k['pfo_bpr']['ms_per_step']*1e3,1), 'linear', round(k['pfo_linear_tf32']['ms_per_step']*1e3,1))
except Exception as e: print(sys.argv[1], 'no line', e)
PY
done
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/w_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/w_pytest.log; tail -3 gpurun_out/w_pytest.log
