#!/bin/bash
# Short round-end session on one GPU (fits in ~8 minutes): the full GPU test suite, smoke(), the bench line, then the ncu launch list of the bench
# command (eager launches of ONE step; shares per kernel only -- numbers printed under ncu are never bench values).
#   bash tools/gpu_short_session.sh <tag>
set -x
mkdir -p gpurun_out
TAG=${1:-r02h}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
timeout 480 python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -24 | tee gpurun_out/pytest_${TAG}.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/bench_${TAG}_nvfp4.json 2> gpurun_out/bench_${TAG}_nvfp4.err
tail -c 3000 gpurun_out/bench_${TAG}_nvfp4.json
tail -3 gpurun_out/bench_${TAG}_nvfp4.err
K='regex:gemm_|quantize_|norm_|add_kernel|mul_add_kernel|activation_kernel|cast_kernel|split_mod|gemv_awq|litela|rope_|attention_fp16|dwconv|lora_partials'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 4800 --csv --log-file gpurun_out/launches_${TAG}_nvfp4.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --no-secondary --no-legs --no-full > gpurun_out/ncu_b_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_b_${TAG}.log | cut -c1-300
ls -la gpurun_out | tail -8
