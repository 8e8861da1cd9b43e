#!/bin/bash
# Short GPU call: parity tests, C4 kernel breakdown, one bench line.
mkdir -p gpurun_out
step() { local t0=$(date +%s); local lim=$1; shift; timeout $lim "$@"; local rc=$?; echo "[step rc=$rc $(( $(date +%s) - t0 ))s] $*" | cut -c1-160; }
step 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
step 100 python tools/bench_large.py c4 10 > gpurun_out/large_c4_n1.log 2>&1; tail -2 gpurun_out/large_c4_n1.log
step 400 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1
python tools/parse_bench.py gpurun_out/bench_quick.log
