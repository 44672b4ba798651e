#!/bin/bash
# round-2 last session: two more knob values against the final code
mkdir -p gpurun_out
for v in "PFO_NONE=0" "PFO_LINEAR_MIN_STAGES=7" "PFO_ATTN_FWD_CTAS=24" "PFO_ATTN_FWD_CTAS=36"; do
  env $v timeout 60 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-profile --large-bs 0 --eval-steps 0 > gpurun_out/zz_var.json 2> gpurun_out/zz_var.err
  python - "$v" <<'PY'
import json, sys
try:
    b=json.loads(open('gpurun_out/zz_var.json').read().strip().split('\n')[-1])
    print(sys.argv[1], round(b['value']), round(b['ms_per_step'],4), round(b['e2e']['value']))
except Exception as e: print(sys.argv[1], 'no line', e)
PY
done
