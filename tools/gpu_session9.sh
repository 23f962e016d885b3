#!/bin/bash
set -x
mkdir -p gpurun_out
TAG=${1:-r01l}
python tools/gpu_run_tests.py --tag ${TAG}_tests --timeout 300 --files tests/test_gpu_fused.py tests/test_gpu_litela.py 2>&1 | tail -30 | grep -v "^pass"
echo "#### fused int4 (256-wide)"
python tools/gemm_prof.py --fused --precision int4 --bn 0 --M 4352 --K 3072 --N 12288 2>&1 | grep -E "==|KERNEL|mma wait op|epi wait|epi pre"
for D in 0 4 8 12; do
  echo "#### 2cta nvfp4 debug=$D"
  NB200_GEMM_DEBUG=$D python tools/gemm_prof.py --precision nvfp4 --bn 512 --M 4352 --K 3072 --N 12288 2>&1 | grep -E "==|KERNEL|mma wait|epi wait|epi pre|producer"
done
python bench.py --steps 8 --warmup 3 --skip-cpu > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print(d['dtype'], 'ms/step', round(d['ms_per_step'],2), 'img/s', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'frac', round(d['roofline']['frac'],3), 'secondary', d['secondary'])"
tail -3 gpurun_out/bench_${TAG}.err
