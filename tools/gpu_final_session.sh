#!/bin/bash
# Round-end session on one GPU: the full GPU test suite, smoke(), the ncu launch list of the bench command (eager launches), ncu --set full of the
# attention kernel / the cluster GEMM / the quantizer, the reference arm and the bench line itself (never under ncu).
#   bash tools/gpu_final_session.sh <tag>
set -x
mkdir -p gpurun_out
TAG=${1:-r02h}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_${TAG}.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
K='regex:gemm_|quantize_|norm_|add_kernel|mul_add_kernel|activation_kernel|cast_kernel|split_mod|gemv_awq|litela|rope_|attention_fp16|dwconv'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 3600 --csv --log-file gpurun_out/launches_${TAG}_nvfp4.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --no-secondary --no-legs --no-full > gpurun_out/ncu_b_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_b_${TAG}.log | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_fp16_v2 -s 2 -c 1 -o gpurun_out/prof_${TAG}_attention -f \
    python tools/attn_bench.py --iters 1 > gpurun_out/ncu_full_attn_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_attn_${TAG}.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_nvfp4_cluster -c 2 -o gpurun_out/prof_${TAG}_gemm_nvfp4 -f \
    python tools/ncu_gemm_one.py --precision nvfp4 --bn 1024 > gpurun_out/ncu_full_gemm_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_gemm_${TAG}.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quantize_v2 -c 2 -o gpurun_out/prof_${TAG}_quant_nvfp4 -f \
    python tools/ncu_quant_one.py --precision nvfp4 > gpurun_out/ncu_full_quant_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_quant_${TAG}.log
timeout 300 python tools/attn_bench.py 2>&1 | tail -6
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
tail -c 600 gpurun_out/bench_${TAG}_reference.json
timeout 900 python bench.py > gpurun_out/bench_${TAG}_nvfp4.json 2> gpurun_out/bench_${TAG}_nvfp4.err
tail -c 2500 gpurun_out/bench_${TAG}_nvfp4.json
tail -3 gpurun_out/bench_${TAG}_nvfp4.err
ls -la gpurun_out | tail -14
