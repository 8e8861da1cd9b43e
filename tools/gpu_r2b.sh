#!/bin/bash
TAG=${1:-r2b}
mkdir -p gpurun_out
step() { local t0=$(date +%s); local lim=$1; shift; timeout $lim "$@"; local rc=$?; echo "[step rc=$rc $(( $(date +%s) - t0 ))s] $*" | cut -c1-170; }
step 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -12 gpurun_out/pytest_gpu_$TAG.log
step 300 python tools/bench_raster.py c3 2>&1 | tee gpurun_out/bench_raster_c3_$TAG.log | cut -c1-400
step 300 python tools/bench_raster.py c4 2>&1 | tee gpurun_out/bench_raster_c4_$TAG.log | cut -c1-400
step 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_20_$TAG.log 2>&1
GSICP_HOST_LM=1 step 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-variants > gpurun_out/bench_20_hostlm_$TAG.log 2>&1
python tools/parse_bench.py gpurun_out/bench_20_$TAG.log gpurun_out/bench_20_hostlm_$TAG.log
step 300 ncu --set full --clock-control none --import-source on -k regex:"align_lm|knn_kernel|covariance_kernel" -s 6 -c 6 -o gpurun_out/prof_lm_$TAG python tools/prof_frame.py 3 > gpurun_out/prof_lm_$TAG.log 2>&1
