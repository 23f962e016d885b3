#!/bin/bash
# One GPU session: bench (both precisions), ncu launch list, ncu full capture of the two top kernels.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
python bench.py --steps 8 --warmup 3 --precision int4 > gpurun_out/bench_${TAG}_int4.json 2> gpurun_out/bench_${TAG}_int4.err
tail -c 3000 gpurun_out/bench_${TAG}_int4.json; tail -5 gpurun_out/bench_${TAG}_int4.err
python bench.py --steps 8 --warmup 3 --precision nvfp4 --skip-cpu > gpurun_out/bench_${TAG}_nvfp4.json 2> gpurun_out/bench_${TAG}_nvfp4.err
tail -c 3000 gpurun_out/bench_${TAG}_nvfp4.json; tail -5 gpurun_out/bench_${TAG}_nvfp4.err
# launch list (cold-cache, serialised): same command, eager launches so kernels are visible by name
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 700 --csv --log-file gpurun_out/launches_${TAG}_int4.csv \
    python bench.py --steps 1 --warmup 3 --precision int4 --no-graph --skip-cpu > gpurun_out/ncu_b_int4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 700 --csv --log-file gpurun_out/launches_${TAG}_nvfp4.csv \
    python bench.py --steps 1 --warmup 3 --precision nvfp4 --no-graph --skip-cpu > gpurun_out/ncu_b_nvfp4.log 2>&1
ls -la gpurun_out
