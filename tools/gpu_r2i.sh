#!/bin/bash
TAG=${1:-r2i}
mkdir -p gpurun_out
step() { local t0=$(date +%s); local lim=$1; shift; timeout $lim "$@"; local rc=$?; echo "[step rc=$rc $(( $(date +%s) - t0 ))s] $*" | cut -c1-170; }
step 600 python -m pytest tests/test_gicp_gpu.py tests/test_gicp_reference.py tests/test_full_size_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_gicp_$TAG.log 2>&1; tail -8 gpurun_out/pytest_gicp_$TAG.log | cut -c1-300
step 120 python tools/prof_align.py 4 > gpurun_out/prof_align_$TAG.log 2>&1; tail -8 gpurun_out/prof_align_$TAG.log | cut -c1-400
export PYTHONPATH=$PWD
step 150 python tests/test_exchange_gpu.py > gpurun_out/exchange_lazy_$TAG.log 2>&1; tail -4 gpurun_out/exchange_lazy_$TAG.log | cut -c1-300
CUDA_MODULE_LOADING=EAGER step 280 python tests/test_exchange_gpu.py > gpurun_out/exchange_eager_$TAG.log 2>&1; tail -4 gpurun_out/exchange_eager_$TAG.log | cut -c1-300
step 300 python -m pytest tests/test_x1_slam_gpu.py -m gpu -q --timeout 280 > gpurun_out/pytest_x1_$TAG.log 2>&1; tail -4 gpurun_out/pytest_x1_$TAG.log | cut -c1-300
