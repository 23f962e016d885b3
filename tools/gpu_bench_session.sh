#!/bin/bash
# One GPU session: bench (both precisions), per-shape sweep, ncu launch list, ncu full capture of the top kernels.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
python tools/op_sweep.py --out gpurun_out/sweep_${TAG}.json > gpurun_out/sweep_${TAG}.log 2>&1
tail -20 gpurun_out/sweep_${TAG}.log
K='regex:gemm_w4a4_kernel|quantize_kernel'
# launch list of ONE timed step (cold-cache, serialised): skip the 3 warm-up steps (3 x 608 launches of our kernels)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 1824 -c 608 --csv --log-file gpurun_out/launches_${TAG}_int4.csv \
    python bench.py --steps 1 --warmup 3 --precision int4 --no-graph --skip-cpu > gpurun_out/ncu_b_int4.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 1824 -c 608 --csv --log-file gpurun_out/launches_${TAG}_nvfp4.csv \
    python bench.py --steps 1 --warmup 3 --precision nvfp4 --no-graph --skip-cpu > gpurun_out/ncu_b_nvfp4.log 2>&1
# full capture of the dominant kernel on the primary shape (3 launches each)
for P in int4 nvfp4; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_w4a4_kernel -s 6 -c 2 -o gpurun_out/prof_${TAG}_gemm_${P} -f \
      python tools/op_sweep.py --precision $P --shapes primary --iters 3 --bn 0 --out gpurun_out/tmp.json > gpurun_out/ncu_full_${P}.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:quantize_kernel -s 4 -c 2 -o gpurun_out/prof_${TAG}_quant_${P} -f \
      python tools/op_sweep.py --precision $P --shapes primary --iters 3 --bn 0 --out gpurun_out/tmp.json > gpurun_out/ncu_fullq_${P}.log 2>&1
done
ls -la gpurun_out
