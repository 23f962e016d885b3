#!/bin/bash
set -x
mkdir -p gpurun_out
TAG=${1:-r01h}
python tools/gpu_run_tests.py --tag ${TAG}_tests --timeout 300 --files tests/test_gpu_cpp_twin.py tests/test_gpu_fused.py 2>&1 | tail -40 | grep -v "^pass"
python -m pytest tests/test_gpu_glue.py -x -q -m gpu 2>&1 | tail -5
python tools/glue_bench.py 2>&1 | tee gpurun_out/glue_bench_${TAG}.jsonl | grep -E "norm"
