#!/bin/bash
TAG=${1:-r2j}
mkdir -p gpurun_out
step() { local t0=$(date +%s); local lim=$1; shift; timeout $lim "$@"; local rc=$?; echo "[step rc=$rc $(( $(date +%s) - t0 ))s] $*" | cut -c1-170; }
step 600 python -m pytest tests/test_gicp_gpu.py tests/test_gicp_reference.py tests/test_full_size_gpu.py tests/test_raster_gpu.py tests/test_exchange_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_a_$TAG.log 2>&1; tail -12 gpurun_out/pytest_a_$TAG.log | cut -c1-300
step 120 python tools/prof_align.py 4 > gpurun_out/prof_align_$TAG.log 2>&1; tail -8 gpurun_out/prof_align_$TAG.log | cut -c1-400
step 120 python tools/bench_raster.py c3 2>&1 | tee gpurun_out/bench_raster_c3_$TAG.log | cut -c1-300
step 120 python tools/bench_raster.py c4 2>&1 | tee gpurun_out/bench_raster_c4_$TAG.log | cut -c1-300
step 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_20_$TAG.log 2>&1
python tools/parse_bench.py gpurun_out/bench_20_$TAG.log
