#!/bin/bash
# Profiling session: ALU micro-benchmark, ncu launch lists of one bench step, ncu --set full of the top kernels,
# then the two bench lines (never under ncu).
set -x
mkdir -p gpurun_out
TAG=${1:-r01i}
LAUNCHES=${2:-589}   # kernels per bench step: 304 GEMM + 228 quantize + 57 more quantizes of the split large-M MLPs
nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench/alu_rates.cu -o gpurun_out/alu_rates && ./gpurun_out/alu_rates | tee gpurun_out/alu_rates_${TAG}.txt
rm -f gpurun_out/alu_rates
K='regex:gemm_w4a4|quantize'
for P in int4 nvfp4; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s $((3 * LAUNCHES)) -c $LAUNCHES --csv --log-file gpurun_out/launches_${TAG}_$P.csv \
      python bench.py --steps 1 --warmup 3 --precision $P --no-graph --skip-cpu > gpurun_out/ncu_b_$P.log 2>&1
  tail -2 gpurun_out/ncu_b_$P.log
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_w4a4 -s 3 -c 1 -o gpurun_out/prof_${TAG}_gemm_${P} -f \
      python tools/op_sweep.py --precision $P --shapes primary --iters 3 --bn 0 --out gpurun_out/tmp.json > gpurun_out/ncu_full_${P}.log 2>&1
  tail -2 gpurun_out/ncu_full_${P}.log
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:quantize_v2 -s 3 -c 1 -o gpurun_out/prof_${TAG}_quant_int4 -f \
    python tools/op_sweep.py --precision int4 --shapes primary --iters 3 --bn 0 --out gpurun_out/tmp.json > gpurun_out/ncu_fullq.log 2>&1
tail -2 gpurun_out/ncu_fullq.log
for P in int4 nvfp4; do
  python bench.py --steps 8 --warmup 3 --precision $P > gpurun_out/bench_${TAG}_$P.json 2> gpurun_out/bench_${TAG}_$P.err
  cat gpurun_out/bench_${TAG}_$P.json | cut -c1-600
  tail -3 gpurun_out/bench_${TAG}_$P.err
done
ls -la gpurun_out | tail -20
