#!/bin/bash
TAG=${1:-r2h}
mkdir -p gpurun_out
step() { local t0=$(date +%s); local lim=$1; shift; timeout $lim "$@"; local rc=$?; echo "[step rc=$rc $(( $(date +%s) - t0 ))s] $*" | cut -c1-170; }
step 600 python -m pytest tests/test_gicp_gpu.py tests/test_gicp_reference.py tests/test_full_size_gpu.py tests/test_map_table_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_gicp_$TAG.log 2>&1; tail -15 gpurun_out/pytest_gicp_$TAG.log | cut -c1-300
step 120 python tools/prof_align.py 6 > gpurun_out/prof_align_$TAG.log 2>&1; tail -12 gpurun_out/prof_align_$TAG.log | cut -c1-400
step 200 python bench.py --config c1 --steps 50 --warmup 5 > gpurun_out/large_c1_n1_$TAG.log 2>&1; python tools/parse_large.py gpurun_out/large_c1_n1_$TAG.log
step 400 python -m pytest tests/test_exchange_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_exchange_$TAG.log 2>&1; tail -5 gpurun_out/pytest_exchange_$TAG.log | cut -c1-300
step 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_20_$TAG.log 2>&1
python tools/parse_bench.py gpurun_out/bench_20_$TAG.log
