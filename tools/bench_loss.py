"""Mapper-loss micro-benchmark: fused CUDA op vs the reference's PyTorch formulation, 640x480, fwd+bwd.
Prints one JSON line: device time per call (CUDA events, L2 flushed between calls), kernel launches, achieved GB/s
against the algorithmic bytes (forward: read 8 planes + write 9 map planes; backward: read 8 + 9 planes, write 4)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from gs_icp_slam_b200 import _lib  # noqa: E402
from gs_icp_slam_b200 import loss as L  # noqa: E402

dev = torch.device("cuda:0")
H, W = 480, 640
g = torch.Generator().manual_seed(0)
gt = torch.rand((3, H, W), generator=g).to(dev)
gtd = (torch.rand((1, H, W), generator=g) * 4).to(dev)
img = (gt + 0.1 * torch.randn((3, H, W), generator=g).to(dev)).clamp(0, 1)
dep = gtd + 0.05 * torch.randn((1, H, W), generator=g).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(fn, n=30):
    ts = []
    for i in range(n + 5):
        a, b = img.clone().requires_grad_(True), dep.clone().requires_grad_(True)
        flush.fill_(i & 0xff)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(a, b).backward()
        e1.record()
        torch.cuda.synchronize()
        if i >= 5:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e3


t_fused = run(lambda a, b: L.mapping_loss(a, b, gt, gtd))
t_torch = run(lambda a, b: bench.torch_mapper_loss(a, b, gt, gtd))
_lib.prof_reset()
_lib.prof_enable(True)
for _ in range(20):
    a, b = img.clone().requires_grad_(True), dep.clone().requires_grad_(True)
    flush.fill_(1)
    L.mapping_loss(a, b, gt, gtd).backward()
torch.cuda.synchronize()
_lib.prof_enable(False)
pr = {k: v[0] / max(v[1], 1) * 1e3 for k, v in _lib.prof_read().items() if v[1]}
plane = H * W * 4
alg = {"loss_forward": (8 + 9) * plane, "loss_backward": (8 + 9 + 4) * plane}
print(json.dumps({"config": "mapper loss fwd+bwd, 640x480", "fused_us_per_call_incl_autograd": t_fused, "torch_us_per_call": t_torch,
                  "speedup": t_torch / t_fused, "kernels_us": pr,
                  "achieved_GBps": {k: alg[k] / (pr[k] * 1e-6) / 1e9 for k in alg if k in pr}}))
os._exit(0)
