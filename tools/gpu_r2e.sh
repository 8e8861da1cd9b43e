#!/bin/bash
TAG=${1:-r2e}
mkdir -p gpurun_out
step() { local t0=$(date +%s); local lim=$1; shift; timeout $lim "$@"; local rc=$?; echo "[step rc=$rc $(( $(date +%s) - t0 ))s] $*" | cut -c1-170; }
step 200 python -m pytest tests/test_exchange_gpu.py tests/test_gicp_gpu.py -m gpu -q --timeout 120 --maxfail 3 > gpurun_out/pytest_exchange_$TAG.log 2>&1; tail -6 gpurun_out/pytest_exchange_$TAG.log | cut -c1-300
step 600 python -m pytest tests -m gpu -q --timeout 300 --maxfail 6 --deselect tests/test_exchange_gpu.py --deselect tests/test_gicp_gpu.py --deselect tests/test_x1_slam_gpu.py > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -12 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
step 120 python tools/bench_raster.py c3 2>&1 | tee gpurun_out/bench_raster_c3_$TAG.log | cut -c1-300
step 120 python tools/bench_raster.py c4 2>&1 | tee gpurun_out/bench_raster_c4_$TAG.log | cut -c1-300
step 300 python tools/run_slam.py --impl ours --frames 100 --timeout 240 2>&1 | tee gpurun_out/run_slam_ours_$TAG.log | tail -3 | cut -c1-2500
step 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_20_$TAG.log 2>&1
python tools/parse_bench.py gpurun_out/bench_20_$TAG.log
