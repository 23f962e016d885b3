#!/bin/bash
set -x
mkdir -p gpurun_out
TAG=${1:-r01b}
python tools/gpu_run_tests.py --tag ${TAG}_tests --timeout 200 --files tests/test_gpu_quantize.py tests/test_gpu_repack.py 2>&1 | tail -40
python tools/gpu_run_tests.py --tag ${TAG}_tests2 --timeout 200 --files tests/test_gpu_gemm.py tests/test_gpu_fused.py -k "linear_module or mlp_module" 2>&1 | tail -10
python tools/op_sweep.py --out gpurun_out/sweep_${TAG}.json --bn 0 > gpurun_out/sweep_${TAG}.log 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/sweep_'+__import__('sys').argv[1] if False else 'gpurun_out/sweep_TAGX.json'.replace('TAGX','%s'))) if False else None
PY
python -c "
import json
d=json.load(open('gpurun_out/sweep_${TAG}.json'))
for r in d['rows']:
    print(f\"{r['precision']:6s} M={r['M']:5d} K={r['K']:5d} N={r['N']:5d} | quant {r['quant_us']:6.1f}us {r['quant_GBs']:6.0f}GB/s ({r['quant_frac_hbm']*100:4.1f}%) | gemm {r['gemm_bn0_us']:6.1f}us {r['gemm_bn0_tflops']:6.0f}TF ({r['gemm_bn0_frac']*100:4.1f}%)\")
"
for P in int4 nvfp4; do
  python bench.py --steps 8 --warmup 3 --precision $P > gpurun_out/bench_${TAG}_$P.json 2> gpurun_out/bench_${TAG}_$P.err
  tail -c 2500 gpurun_out/bench_${TAG}_$P.json; tail -3 gpurun_out/bench_${TAG}_$P.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quantize_v2_kernel -s 3 -c 1 -o gpurun_out/prof_${TAG}_quant_int4 -f python tools/op_sweep.py --precision int4 --shapes primary --iters 3 --bn 0 --out gpurun_out/tmp.json > gpurun_out/ncu_fullq_int4.log 2>&1
ls gpurun_out
