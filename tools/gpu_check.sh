#!/bin/bash
# Short single-GPU check after a rasterizer kernel change: parity tests, per-kernel timings, ncu of the render kernels.
TAG=${1:-chk}
mkdir -p gpurun_out
step() { local t0=$(date +%s); local lim=$1; shift; timeout $lim "$@"; local rc=$?; echo "[step rc=$rc $(( $(date +%s) - t0 ))s] $*" | cut -c1-170; }
step 600 python -m pytest tests/test_raster_gpu.py tests/test_full_size_gpu.py tests/test_x1_slam_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_chk_$TAG.log 2>&1; tail -6 gpurun_out/pytest_chk_$TAG.log | cut -c1-300
step 120 python tools/bench_raster.py c3 2>&1 | tee gpurun_out/bench_raster_c3_$TAG.log | cut -c1-300
step 120 python tools/bench_raster.py c4 2>&1 | tee gpurun_out/bench_raster_c4_$TAG.log | cut -c1-300
step 60 python tools/prof_align.py 3 2>&1 | tail -4 | cut -c1-300
step 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_20_$TAG.log 2>&1
python tools/parse_bench.py gpurun_out/bench_20_$TAG.log
[ "${NCU:-1}" = "1" ] && step 300 ncu --set full --clock-control none --import-source on -k regex:"render_backward|render_forward|gaussian_backward" -s 6 -c 6 -o gpurun_out/prof_$TAG python tools/prof_frame.py 3 > gpurun_out/prof_$TAG.log 2>&1
