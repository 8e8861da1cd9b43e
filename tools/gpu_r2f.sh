#!/bin/bash
TAG=${1:-r2f}
mkdir -p gpurun_out
step() { local t0=$(date +%s); local lim=$1; shift; timeout $lim "$@"; local rc=$?; echo "[step rc=$rc $(( $(date +%s) - t0 ))s] $*" | cut -c1-170; }
step 420 python -m pytest tests/test_exchange_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_exchange_$TAG.log 2>&1; tail -25 gpurun_out/pytest_exchange_$TAG.log | cut -c1-300
step 900 python -m pytest tests -m gpu -q --timeout 300 --maxfail 8 --deselect tests/test_exchange_gpu.py > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -14 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
step 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
step 120 python tools/bench_raster.py c3 2>&1 | tee gpurun_out/bench_raster_c3_$TAG.log | cut -c1-300
step 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_20_$TAG.log 2>&1
python tools/parse_bench.py gpurun_out/bench_20_$TAG.log
