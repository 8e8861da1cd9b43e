#!/bin/bash
# Short GPU call: parity tests, loss micro-benchmark, bench lines for the three mapper-loss variants.
mkdir -p gpurun_out
step() { local t0=$(date +%s); local lim=$1; shift; timeout $lim "$@"; local rc=$?; echo "[step rc=$rc $(( $(date +%s) - t0 ))s] $*" | cut -c1-160; }
step 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
step 200 python tools/bench_loss.py > gpurun_out/bench_loss.log 2>&1; tail -3 gpurun_out/bench_loss.log
for l in l1 ssim_torch ssim_fused; do
  step 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --loss $l > gpurun_out/bench_$l.log 2>&1
  python tools/parse_bench.py gpurun_out/bench_$l.log
done
step 200 python bench.py --impl reference --steps 8 --warmup 2 --loss ssim_torch > gpurun_out/bench_ref_ssim.log 2>&1; tail -c 400 gpurun_out/bench_ref_ssim.log
