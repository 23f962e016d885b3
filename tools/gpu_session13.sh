#!/bin/bash
set -x
mkdir -p gpurun_out
TAG=${1:-r01p}
timeout 1500 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_fused.py tests/test_gpu_cpp_twin.py -m gpu -q --timeout 180 -x 2>&1 | tail -6
python tools/gemm_prof.py --precision int4 --bn 0 --M 4352 --K 3072 --N 12288 2>&1 | grep -E "==|KERNEL|mma wait|epi wait|producer"
python tools/op_sweep.py --precision int4 --out gpurun_out/sweep_${TAG}.json --bn 0 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(f\"{r['precision']:6s} M={r['M']:5d} K={r['K']:5d} N={r['N']:5d} | quant {r['quant_us']:6.1f}us ({r['quant_frac_hbm']*100:4.1f}%) | gemm {r['gemm_bn0_us']:6.1f}us {r['gemm_bn0_tflops']:6.0f}TF ({r['gemm_bn0_frac']*100:4.1f}%)\")
"
echo "#### fused int4"
python tools/gemm_prof.py --fused --precision int4 --bn 0 --M 4352 --K 3072 --N 12288 2>&1 | grep -E "=="
python bench.py --steps 8 --warmup 3 --skip-cpu > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print(d['dtype'], 'ms/step', round(d['ms_per_step'],2), 'img/s', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'frac', round(d['roofline']['frac'],3), 'secondary', d['secondary'])"
tail -3 gpurun_out/bench_${TAG}.err
