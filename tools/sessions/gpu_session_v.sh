#!/bin/bash
# round-2 session V: ring depth wanted by the linear kernel (PFO_LINEAR_MIN_STAGES) against the step time
mkdir -p gpurun_out
for w in 2 3 4 5; do
  PFO_LINEAR_MIN_STAGES=$w timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --large-bs 0 --eval-steps 0 > gpurun_out/v_w$w.json 2> gpurun_out/v_w$w.err
  python - $w <<'PY'
import json, sys
try:
    b=json.loads(open('gpurun_out/v_w%s.json' % sys.argv[1]).read().strip().split('\n')[-1])
    k=b['kernels']
    print('want', sys.argv[1], round(b['value']), round(b['ms_per_step'],4), 'linear', round(k['pfo_linear_tf32']['ms_per_step']*1e3,1), 'wgrad', round(k['pfo_wgrad_tf32']['ms_per_step']*1e3,1))
except Exception as e: print(sys.argv[1], 'no line', e)
PY
done
PFO_LINEAR_MIN_STAGES=3 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "linear_tf32 or attn_nbr" 2>&1 | tail -2
