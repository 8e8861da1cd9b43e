#!/bin/bash
# Short GPU call: parity tests, smoke, the default bench line of both arms.
mkdir -p gpurun_out
step() { local t0=$(date +%s); local lim=$1; shift; timeout $lim "$@"; local rc=$?; echo "[step rc=$rc $(( $(date +%s) - t0 ))s] $*" | cut -c1-160; }
step 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
step 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
step 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.log 2>&1
step 600 python bench.py > gpurun_out/bench_default.log 2>&1
python tools/parse_bench.py gpurun_out/bench_default.log gpurun_out/bench_ref.log
