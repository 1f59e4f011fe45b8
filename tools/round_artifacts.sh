#!/bin/bash
# Run on the GPU box (under gpurun): every measured artifact tools/make_profile_summary.py turns into profiles/<tag>_*.
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2>> gpurun_out/bench_${TAG}.err
for w in hopper_ens64 pen_sweep256 hopper_single antmaze_jsrl; do
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --workload $w > gpurun_out/bench_${TAG}_$w.json 2>> gpurun_out/bench_${TAG}.err
done
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --workload stress_4x1024 > gpurun_out/bench_${TAG}_stress.json 2>> gpurun_out/bench_${TAG}.err
python tools/dropin_rate.py > gpurun_out/dropin_rate_${TAG}.txt 2>&1
IQL_FUSED_TRACE=1 python tools/fused_trace.py halfcheetah_ens64 > gpurun_out/fused_trace_${TAG}.txt 2>&1
python tools/umma_rate.py > gpurun_out/umma_rate_${TAG}.txt 2>&1
bash tools/profile.sh ${TAG}
tail -c 600 gpurun_out/bench_${TAG}.json
