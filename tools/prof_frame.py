"""Runs a few SLAM frames of the bench workload (resident mode) — the command ncu wraps for profiles/."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cam, gmap, frames = bench.make_sequence(n + 1, bench.MAP_P)
eng = bench.Ours(cam, gmap, frames, torch.device("cuda:0"), 1, 0, loss="ssim_fused")
for i in range(n):
    eng.step(i, resident=True)
torch.cuda.synchronize()
print("done", n)
