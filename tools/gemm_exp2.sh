#!/bin/bash
# Ablation of the fused fc1 epilogue (EPI_QUANT): NB200_GEMM_DEBUG bits 4 (no MMAs), 8 (no epilogue), 16 (no atomics), 32 (no quantise math)
mkdir -p gpurun_out
OUT=gpurun_out/${1:-exp2}.txt
: > $OUT
for P in nvfp4 int4; do
  for D in 0 16 32 48 8 4; do
    echo "#### fused precision=$P debug=$D" >> $OUT
    NB200_GEMM_DEBUG=$D timeout 120 python tools/gemm_prof.py --fused --precision $P --bn 0 --M 4352 --K 3072 --N 12288 2>&1 | tail -16 >> $OUT
  done
done
grep -E "####|==|KERNEL|epi|mma wait" $OUT
