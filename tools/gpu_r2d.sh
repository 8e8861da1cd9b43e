#!/bin/bash
TAG=${1:-r2d}
mkdir -p gpurun_out
step() { local t0=$(date +%s); local lim=$1; shift; timeout $lim "$@"; local rc=$?; echo "[step rc=$rc $(( $(date +%s) - t0 ))s] $*" | cut -c1-170; }
step 300 python -m pytest tests/test_exchange_gpu.py -m gpu -q -x > gpurun_out/pytest_exchange_$TAG.log 2>&1; tail -5 gpurun_out/pytest_exchange_$TAG.log
if ! grep -q " passed" gpurun_out/pytest_exchange_$TAG.log || grep -q "failed" gpurun_out/pytest_exchange_$TAG.log; then
  GSICP_DEBUG_SYNC=1 step 300 python -m pytest tests/test_exchange_gpu.py -m gpu -q -x -k "False" > gpurun_out/pytest_exchange_dbg_$TAG.log 2>&1
  grep -n "gsicp debug\|Error\|error" gpurun_out/pytest_exchange_dbg_$TAG.log | head -20
fi
step 900 python -m pytest tests -m gpu -q --deselect tests/test_exchange_gpu.py > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -15 gpurun_out/pytest_gpu_$TAG.log
step 600 python tools/run_slam.py --impl ours --frames 100 --keep 2>&1 | tee gpurun_out/run_slam_ours_$TAG.log | tail -3 | cut -c1-1500
step 900 python tools/run_slam.py --impl reference --frames 100 2>&1 | tee gpurun_out/run_slam_ref_$TAG.log | tail -3 | cut -c1-1500
