#!/bin/bash
# Multi-GPU batch (gpurun --gpus N): sharding parity test, then the large configs at 1 and N GPUs.
N=${1:-2}; TAG=${2:-r2c}
mkdir -p gpurun_out
step() { local t0=$(date +%s); local lim=$1; shift; timeout $lim "$@"; local rc=$?; echo "[step rc=$rc $(( $(date +%s) - t0 ))s] $*" | cut -c1-200; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m 2>/dev/null | head -12
step 900 python -m pytest tests/test_multigpu_gpu.py tests/test_gicp_gpu.py tests/test_map_table_gpu.py -m gpu -q > gpurun_out/pytest_multigpu_$TAG.log 2>&1; tail -25 gpurun_out/pytest_multigpu_$TAG.log
for cfg in c4 c5; do
  step 300 python bench.py --config $cfg --steps 10 --warmup 3 > gpurun_out/large_${cfg}_n1_$TAG.log 2>&1; tail -1 gpurun_out/large_${cfg}_n1_$TAG.log | cut -c1-700
  step 300 $TR --nproc-per-node $N --master-port 29702 bench.py --gpus $N --config $cfg --steps 10 --warmup 3 > gpurun_out/large_${cfg}_n${N}_$TAG.log 2>&1; tail -1 gpurun_out/large_${cfg}_n${N}_$TAG.log | cut -c1-700
done
