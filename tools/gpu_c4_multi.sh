#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python tools/bench_large.py c4 10 > gpurun_out/large_c4_n1.log 2>&1; grep -v "^\[" gpurun_out/large_c4_n1.log | tail -2
timeout 300 $TR --nproc-per-node $N --master-port 29702 tools/bench_large.py c4 10 > gpurun_out/large_c4_n$N.log 2>&1; grep "^{" gpurun_out/large_c4_n$N.log | tail -2
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
