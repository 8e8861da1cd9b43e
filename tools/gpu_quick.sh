#!/bin/bash
# Short GPU call: parity tests + the default bench line.
mkdir -p gpurun_out
step() { local t0=$(date +%s); local lim=$1; shift; timeout $lim "$@"; local rc=$?; echo "[step rc=$rc $(( $(date +%s) - t0 ))s] $*" | cut -c1-160; }
step 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
step 600 python bench.py --no-cpu-baseline > gpurun_out/bench_default.log 2>&1
python tools/parse_bench.py gpurun_out/bench_default.log
python - <<PY
import json
d=json.loads([x for x in open("gpurun_out/bench_default.log") if x.startswith("{")][-1])
print("loss_variants", d.get("loss_variants")); print("schedules", d.get("schedules"))
PY
