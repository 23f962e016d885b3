#!/bin/bash
# Profiling session (one GPU): ncu launch list of one bench step (eager launches, same workload as the default command), ncu --set full of the
# top GEMM and of the quantizer, then the bench line itself (never under ncu).
#   bash tools/gpu_profile_session.sh <tag>
set -x
mkdir -p gpurun_out
TAG=${1:-r02e}
# our kernels only (torch's weight-initialisation kernels come first and are not part of a step); an eager step launches ~570 of them:
# skip the 3 warm-up steps, list one timed step
K='regex:gemm_|quantize_|norm_|add_kernel|mul_add_kernel|activation_kernel|cast_kernel|split_mod|gemv_awq|litela'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 1710 -c 570 --csv --log-file gpurun_out/launches_${TAG}_nvfp4.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --no-secondary --no-legs > gpurun_out/ncu_b_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_b_${TAG}.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_nvfp4_cluster -c 2 -o gpurun_out/prof_${TAG}_gemm_nvfp4 -f \
    python tools/ncu_gemm_one.py --precision nvfp4 --bn 1024 > gpurun_out/ncu_full_gemm_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_gemm_${TAG}.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quantize_v2 -c 2 -o gpurun_out/prof_${TAG}_quant_nvfp4 -f \
    python tools/ncu_quant_one.py --precision nvfp4 > gpurun_out/ncu_full_quant_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_quant_${TAG}.log
python bench.py > gpurun_out/bench_${TAG}_nvfp4.json 2> gpurun_out/bench_${TAG}_nvfp4.err
tail -c 1500 gpurun_out/bench_${TAG}_nvfp4.json
tail -3 gpurun_out/bench_${TAG}_nvfp4.err
ls -la gpurun_out | tail -12
