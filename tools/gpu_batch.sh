#!/bin/bash
# One GPU call = tests + goldens + benches + profiles (the pod queue is long: batch everything).  Every step is
# wrapped in `timeout` and timed so a slow or hanging step is visible and bounded.
mkdir -p gpurun_out
step() { local t0=$(date +%s); local lim=$1; shift; timeout $lim "$@"; local rc=$?; echo "[step rc=$rc $(( $(date +%s) - t0 ))s] $*" | cut -c1-160; }
step 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
step 120 python tests/golden/make_raster_golden.py > gpurun_out/golden.log 2>&1 && cp tests/golden/raster_ref_small.npz gpurun_out/; tail -2 gpurun_out/golden.log
step 200 python -X faulthandler tools/host_breakdown2.py > gpurun_out/host_breakdown2.log 2>&1; tail -9 gpurun_out/host_breakdown2.log
step 600 python bench.py > gpurun_out/bench_ovl.log 2>&1
step 100 python tools/bench_large.py c4 10 > gpurun_out/large_c4_n1.log 2>&1; tail -2 gpurun_out/large_c4_n1.log
step 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.log 2>&1
step 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
step 200 python tools/gpu_diag_raster.py > gpurun_out/diag_raster.log 2>&1; grep -E "mismatch|time ms" gpurun_out/diag_raster.log | tail -12
grep -h "step rc" gpurun_out/*.log 2>/dev/null
python - <<PY
import json
for f in ("bench_ovl","bench_ref"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3))
        if "schedules" in d: print("   ", d["schedules"])
        if "kernels" in d: print("   ", {k: round(v["ms_per_step"]*1e3,1) for k,v in d["kernels"].items()})
    except Exception as e:
        print(f, "FAILED", e, open(f"gpurun_out/{f}.log").read()[-1500:])
PY
step 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 1 --overlap 0 --no-cpu-baseline --no-roofline --no-variants > gpurun_out/b_ncu.log 2>&1
step 400 ncu --set full --clock-control none --import-source on -k regex:"render_backward|render_forward|tile_sort|mapping_loss|knn_kernel|linearize_kernel|error_kernel" -s 10 -c 11 -o gpurun_out/prof_r1f python tools/prof_frame.py 3 > gpurun_out/prof_r1f.log 2>&1
ls -la gpurun_out | tail -15
