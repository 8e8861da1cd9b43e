#!/bin/bash
# 8-GPU batch (gpurun --gpus 8): sharding parity (2 ranks), strong scaling of the large configs at 1/2/4/8, replicas at 8.
TAG=${1:-multi}
mkdir -p gpurun_out
step() { local t0=$(date +%s); local lim=$1; shift; timeout $lim "$@"; local rc=$?; echo "[step rc=$rc $(( $(date +%s) - t0 ))s] $*" | cut -c1-200; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m 2>/dev/null | head -11 | cut -c1-120
step 400 python -m pytest tests/test_multigpu_gpu.py tests/test_gicp_gpu.py tests/test_gicp_reference.py -m gpu -q --timeout 250 > gpurun_out/pytest_multigpu_$TAG.log 2>&1; tail -6 gpurun_out/pytest_multigpu_$TAG.log | cut -c1-250
for cfg in c4 c5; do
  step 200 python bench.py --config $cfg --steps 10 --warmup 3 > gpurun_out/large_${cfg}_n1_$TAG.log 2>&1
  for n in 2 4 8; do
    step 200 $TR --nproc-per-node $n --master-port $((29700 + n)) bench.py --gpus $n --config $cfg --steps 10 --warmup 3 > gpurun_out/large_${cfg}_n${n}_$TAG.log 2>&1
  done
  python tools/parse_large.py gpurun_out/large_${cfg}_n*_$TAG.log | cut -c1-600
done
step 300 $TR --nproc-per-node 8 --master-port 29720 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/bench_replicas_n8_$TAG.log 2>&1
python tools/parse_bench.py gpurun_out/bench_replicas_n8_$TAG.log
