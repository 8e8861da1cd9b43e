#!/bin/bash
# Multi-GPU batch (run with gpurun --gpus N): sharding parity test, bench.py under torchrun, large-config scaling.
N=${1:-2}
mkdir -p gpurun_out
step() { local t0=$(date +%s); local lim=$1; shift; timeout $lim "$@"; local rc=$?; echo "[step rc=$rc $(( $(date +%s) - t0 ))s] $*" | cut -c1-200; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
step 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_multigpu.log 2>&1; tail -5 gpurun_out/pytest_multigpu.log
step 400 $TR --nproc-per-node $N --master-port 29701 bench.py --gpus $N --steps 15 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; tail -c 1500 gpurun_out/bench_n$N.log
for cfg in c4 c5; do
  step 400 python tools/bench_large.py $cfg 10 > gpurun_out/large_${cfg}_n1.log 2>&1; tail -1 gpurun_out/large_${cfg}_n1.log
  step 400 $TR --nproc-per-node $N --master-port 29702 tools/bench_large.py $cfg 10 > gpurun_out/large_${cfg}_n$N.log 2>&1; tail -2 gpurun_out/large_${cfg}_n$N.log
done
