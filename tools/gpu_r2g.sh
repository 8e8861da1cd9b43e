#!/bin/bash
# 2-GPU batch: sharding parity (both transports, persistent LM kernel with in-kernel exchange), new render_backward parity +
# timing, LM phase marks, large configs at 1 and N GPUs.
N=${1:-2}; TAG=${2:-r2g}
mkdir -p gpurun_out
step() { local t0=$(date +%s); local lim=$1; shift; timeout $lim "$@"; local rc=$?; echo "[step rc=$rc $(( $(date +%s) - t0 ))s] $*" | cut -c1-200; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
step 600 python -m pytest tests/test_multigpu_gpu.py tests/test_map_table_gpu.py tests/test_raster_gpu.py tests/test_full_size_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_multigpu_$TAG.log 2>&1; tail -25 gpurun_out/pytest_multigpu_$TAG.log | cut -c1-250
step 120 python tools/bench_raster.py c3 2>&1 | tee gpurun_out/bench_raster_c3_$TAG.log | cut -c1-300
step 120 python tools/bench_raster.py c4 2>&1 | tee gpurun_out/bench_raster_c4_$TAG.log | cut -c1-300
step 120 python tools/prof_align.py 6 > gpurun_out/prof_align_$TAG.log 2>&1; tail -14 gpurun_out/prof_align_$TAG.log | cut -c1-400
step 200 python bench.py --config c1 --steps 50 --warmup 5 > gpurun_out/large_c1_n1_$TAG.log 2>&1; tail -1 gpurun_out/large_c1_n1_$TAG.log | cut -c1-500
for cfg in c4 c5; do
  step 300 python bench.py --config $cfg --steps 10 --warmup 3 > gpurun_out/large_${cfg}_n1_$TAG.log 2>&1; tail -1 gpurun_out/large_${cfg}_n1_$TAG.log | cut -c1-700
  step 300 $TR --nproc-per-node $N --master-port 29702 bench.py --gpus $N --config $cfg --steps 10 --warmup 3 > gpurun_out/large_${cfg}_n${N}_$TAG.log 2>&1; tail -1 gpurun_out/large_${cfg}_n${N}_$TAG.log | cut -c1-700
done
