#!/bin/bash
set -x
mkdir -p gpurun_out
TAG=${1:-r01f}
python tools/gpu_run_tests.py --tag ${TAG}_tests --timeout 200 --files tests/test_gpu_gemm.py tests/test_gpu_fused.py 2>&1 | tail -70 | grep -v "^pass" 
python tools/gemm_prof.py --precision int4 --bn 0 2>&1 | tail -16
python tools/gemm_prof.py --precision nvfp4 --bn 0,512 --M 4352 --K 3072 --N 12288 2>&1 | tail -30
python tools/op_sweep.py --out gpurun_out/sweep_${TAG}.json --bn 0,512 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(f\"{r['precision']:6s} M={r['M']:5d} K={r['K']:5d} N={r['N']:5d} | quant {r['quant_us']:6.1f}us ({r['quant_frac_hbm']*100:4.1f}%) | auto {r['gemm_bn0_us']:6.1f}us {r['gemm_bn0_tflops']:6.0f}TF ({r['gemm_bn0_frac']*100:4.1f}%) | 2cta {r['gemm_bn512_us']:6.1f}us {r['gemm_bn512_tflops']:6.0f}TF ({r['gemm_bn512_frac']*100:4.1f}%)\")
"
for P in int4 nvfp4; do
  python bench.py --steps 8 --warmup 3 --precision $P --skip-cpu > gpurun_out/bench_${TAG}_$P.json 2> gpurun_out/bench_${TAG}_$P.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}_$P.json')); print('$P', 'ms/step', round(d['ms_per_step'],2), 'img/s', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'gemm frac', round(d['roofline']['frac'],3), 'avg gemm us', round(d['roofline']['avg_launch_us'],1))"
  tail -3 gpurun_out/bench_${TAG}_$P.err
done
