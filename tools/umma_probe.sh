#!/bin/bash
# Profiling experiment (results of the training step are WRONG under IQL_UMMA_DBG != 0): per-phase times of the
# default workload with parts of the tcgen05 kernel switched off, to see which resource bounds each phase.
#   1 = no global stores in the epilogue, 2 = operands of 4 problems only (L2 hits), 4 = no epilogue at all,
#   8 = no MMAs (TMA only), 16 = no TMA (MMA only), 32 = no tensor-map prefetch
export IQL_B200_DEBUG=1  # the IQL_* switches below are debug facilities behind this master flag
mkdir -p gpurun_out
for pair in 0 1; do
  for dbg in 0 32 4 36 12 20 6 14; do
    if [ $pair = 0 ]; then export IQL_B200_NO_CTA2=1; else unset IQL_B200_NO_CTA2; fi
    IQL_UMMA_DBG=$dbg timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --fast-init \
      > gpurun_out/probe_p${pair}_d${dbg}.json 2> gpurun_out/probe_p${pair}_d${dbg}.err || echo "pair=$pair dbg=$dbg failed"
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/probe_p*_d*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as ex:
        print(f, "unreadable", ex); continue
    ks = {k["kernel"]: k["us"] for k in d["kernels"]}
    print(f.split("/")[-1], {k: ks[k] for k in ("first_fwd_fwd", "hidden_fwd", "hidden_wgrad", "hidden_dgrad")})
PY
