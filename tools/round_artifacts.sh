#!/bin/bash
# Run on the GPU box (under gpurun): every measured artifact tools/make_profile_summary.py turns into profiles/<tag>_*.
TAG=${1:-r02}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_${TAG}_ref.json 2>> gpurun_out/bench_${TAG}.err
[ -n "$ALL_CONFIGS" ] && python bench.py --all-configs > gpurun_out/all_configs_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
python tools/act_latency.py > gpurun_out/act_latency_${TAG}.json 2>&1
IQL_FUSED_TRACE=1 python tools/fused_trace.py halfcheetah_ens64 > gpurun_out/fused_trace_${TAG}.txt 2>&1
python tools/chain_trace.py 64 > gpurun_out/chain_trace_${TAG}.txt 2>&1
python tools/umma_rate.py > gpurun_out/umma_rate_${TAG}.txt 2>&1
bash tools/profile.sh ${TAG}
tail -c 600 gpurun_out/bench_${TAG}.json
