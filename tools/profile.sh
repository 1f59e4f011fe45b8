#!/bin/bash
# Run on the GPU box (under gpurun): launch list + full ncu capture of the engine's kernels.
#   tools/profile.sh <tag>     -> gpurun_out/launches_<tag>.csv, gpurun_out/prof_<tag>.ncu-rep
export IQL_B200_DEBUG=1  # the IQL_* switches below are debug facilities behind this master flag
set -u
TAG=${1:-r01}
export IQL_B200_GRAPHS=0
CMD="python bench.py --steps 2 --warmup 3 --inner 2 --no-cpu-baseline --no-eager --no-dropin --fast-init"
KRE='regex:^(bwd_chain|adam_polyak|advance|colsum|first_fwd|first_wgrad|gather|last_bwd|last_bwd_v4|loss|refresh_shadow|umma_gemm|fused_fwd|out_fwd|out_fwd_rows|simt_gemm)_kernel'
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -s 60 -c 120 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k "$KRE" -s 60 -c 12 \
    -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu2_${TAG}.log 2>&1
ls -la gpurun_out/ | grep ${TAG}
