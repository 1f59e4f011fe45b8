#!/bin/bash
# Timing probes of the fused forward's epilogue (run on the GPU box): IQL_FUSED_DBG switches parts of the epilogue
# off (the training results are wrong in these runs; only the fused_fwd time is read).
export IQL_B200_DEBUG=1  # the IQL_* switches below are debug facilities behind this master flag
for d in ${@:-0 1 2 3 4 8 15}; do
  IQL_FUSED_DBG=$d timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --fast-init 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k={r['kernel']:r['us'] for r in d['kernels']}
print('dbg=$d', 'fused_fwd_us', k.get('fused_fwd'), 'steps/s', round(d['value']))" || echo "dbg=$d failed"
done
