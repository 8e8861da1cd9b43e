#!/bin/bash
# Full single-GPU batch (gpurun): parity tests, smoke, both bench arms, launch list, ncu captures of our top kernels and of the
# reference's kernels ("kernel to beat").  Every step is wrapped in `timeout` and timed.
TAG=${1:-r2a}
mkdir -p gpurun_out
step() { local t0=$(date +%s); local lim=$1; shift; timeout $lim "$@"; local rc=$?; echo "[step rc=$rc $(( $(date +%s) - t0 ))s] $*" | cut -c1-170; }
step 1200 python -m pytest tests -m gpu -q --timeout 300 --durations=8 > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -22 gpurun_out/pytest_gpu_$TAG.log | cut -c1-200
step 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
step 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$TAG.log 2>&1
step 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_20_$TAG.log 2>&1
step 600 python bench.py > gpurun_out/bench_default_$TAG.log 2>&1
python tools/parse_bench.py gpurun_out/bench_20_$TAG.log gpurun_out/bench_default_$TAG.log gpurun_out/bench_ref_$TAG.log
tail -3 gpurun_out/bench_20_$TAG.log | cut -c1-600
if [ "${NCU:-1}" = "1" ]; then
step 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-variants > gpurun_out/b_ncu_$TAG.log 2>&1
step 400 ncu --set full --clock-control none --import-source on -k regex:"render_backward|render_forward|tile_sort|preprocess_kernel|gaussian_backward|emit_binned|knn_kernel|covariance_kernel|align_lm" -s 12 -c 14 -o gpurun_out/prof_$TAG python tools/prof_frame.py 3 > gpurun_out/prof_$TAG.log 2>&1
[ "${NCU_REF:-1}" = "1" ] && step 400 ncu --set full --clock-control none -k regex:"renderCUDA|preprocessCUDA|duplicateWithKeys|identifyTileRanges|DeviceRadixSort|computeCov2DCUDA" -s 16 -c 16 -o gpurun_out/prof_ref_$TAG python tools/prof_frame.py 3 reference > gpurun_out/prof_ref_$TAG.log 2>&1
fi
ls -la gpurun_out | tail -12
