#!/bin/bash
# Short round-end session on one GPU (fits in ~6 minutes of box time): the full GPU test suite, smoke(), the bench line, then evidence for profiles/:
# the ncu launch list of the kernels that changed last (rope / reduction), ncu --set full of the attention kernel and of the LiteLA epilogue kernel.
# Numbers printed under ncu are never bench values.
#   bash tools/gpu_short_session.sh <tag> [full-list]
set -x
mkdir -p gpurun_out
TAG=${1:-r02j}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
timeout 400 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -18 | tee gpurun_out/pytest_${TAG}.txt
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 240 python bench.py > gpurun_out/bench_${TAG}_nvfp4.json 2> gpurun_out/bench_${TAG}_nvfp4.err
tail -c 1500 gpurun_out/bench_${TAG}_nvfp4.json
tail -3 gpurun_out/bench_${TAG}_nvfp4.err
if [ "$2" = "full-list" ]; then
  K='regex:gemm_|quantize_|norm_|add_kernel|mul_add_kernel|activation_kernel|cast_kernel|split_mod|gemv_awq|litela|rope_|attention_fp16|dwconv|lora_partials'
  C=4800
else
  K='regex:rope_|lora_partials|norm_warp|add_kernel'
  C=1400
fi
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c $C --csv --log-file gpurun_out/launches_${TAG}_nvfp4.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --no-secondary --no-legs --no-full > gpurun_out/ncu_b_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_b_${TAG}.log | cut -c1-200
timeout 120 ncu --set full --clock-control none --import-source on -k regex:attention_fp16_v2 -s 2 -c 1 -o gpurun_out/prof_${TAG}_attention -f \
    python tools/attn_bench.py --iters 1 > gpurun_out/ncu_full_attn_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_attn_${TAG}.log | cut -c1-200
timeout 120 ncu --set full --clock-control none --import-source on -k regex:gemm_w4a4_kernel -s 2 -c 1 -o gpurun_out/prof_${TAG}_litela -f \
    python tools/litela_bench.py --precision nvfp4 --iters 1 > gpurun_out/ncu_full_litela_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_litela_${TAG}.log | cut -c1-200
ls -la gpurun_out | tail -12
