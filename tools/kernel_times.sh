#!/bin/bash
# Per-kernel CUDA-event times of the default bench workload (run on the GPU box).  Extra arguments go to bench.py.
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --fast-init "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('steps/s', round(d['value']), 'e2e', round(d['e2e']['value']), d['config']['workload'])
for r in d['kernels']: print(f\"  {r['kernel']:<20} {r['us']:8.2f} us  {r['gbs']:8.1f} GB/s  {r['tflops']:7.2f} TF/s  frac {r['frac']}\")"
