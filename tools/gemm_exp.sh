#!/bin/bash
# Ablation timings of the fused GEMM (NB200_GEMM_DEBUG bits; results are invalid, timings are the point)
mkdir -p gpurun_out
OUT=gpurun_out/${1:-exp}.txt
: > $OUT
for P in int4 nvfp4; do
  for D in 0 1 2 4 8 5 6 12 14; do
    if [ $P = nvfp4 ] && [ $D != 0 ] && [ $D != 4 ] && [ $D != 8 ] && [ $D != 12 ]; then continue; fi
    echo "#### precision=$P debug=$D" >> $OUT
    NB200_GEMM_DEBUG=$D timeout 120 python tools/gemm_prof.py --precision $P --bn 0 --M 4352 --K 3072 --N 12288 2>&1 | tail -16 >> $OUT
  done
done
grep -E "####|==|KERNEL" $OUT
