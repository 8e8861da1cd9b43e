"""Finer wall-clock breakdown of the mapper iteration's backward.  Development tool."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from gs_icp_slam_b200 import rasterizer as R  # noqa: E402

n = 16
cam, gmap, frames = bench.make_sequence(2, bench.MAP_P)
eng = bench.Ours(cam, gmap, frames, torch.device("cuda:0"), 1, 0)
f = frames[1]
c, m = f["d_cam"], eng.map
acc = {}


def tick(name, t0):
    torch.cuda.synchronize()
    t = time.perf_counter()
    acc.setdefault(name, []).append(t - t0)
    return t


e = torch.Tensor([])
for it in range(n):
    rs = eng.Settings(cam["H"], cam["W"], c["tanfovx"], c["tanfovy"], eng.bg, 1.0, c["viewmatrix"], c["projmatrix"], 0, c["campos"], False, False)
    torch.cuda.synchronize(); t = time.perf_counter()
    depth, color, radii, is_used = eng.Rasterizer(rs)(means3D=m["means3D"], means2D=eng.means2D, opacities=m["opacities"], shs=m["shs"], scales=m["scales"], rotations=m["rotations"])
    t = tick("fwd(autograd)", t)
    loss = (color - f["d_rgb"]).abs().mean() + 0.1 * (depth - f["d_depth"]).abs().mean(); t = tick("loss", t)
    gcol, gdep = torch.autograd.grad(loss, [color, depth], retain_graph=True); t = tick("grad(loss->image)", t)
    ctx = color.grad_fn
    sv = ctx.saved_tensors
    t = tick("saved_tensors", t)
    out = R.rasterize_gaussians_backward(eng.bg, m["means3D"].detach(), sv[5], e, m["scales"].detach(), m["rotations"].detach(), 1.0, e,
                                         c["viewmatrix"], c["projmatrix"], c["tanfovx"], c["tanfovy"], gdep, gcol, m["shs"].detach(), 0,
                                         c["campos"], sv[7], ctx.num_rendered, sv[8], sv[9], False)
    t = tick("raster_bwd(direct)", t)
    del out
    loss.backward(); t = tick("loss.backward(full)", t)
    for k in m:
        m[k].grad = None
    eng.means2D.grad = None
    t = tick("clear grads", t)
for k, v in acc.items():
    v = np.array(v[4:]) * 1e3
    print(f"{k:24s} median {np.median(v):8.3f} ms  min {v.min():8.3f}  max {v.max():8.3f}")
print("cpu_count", os.cpu_count(), "load", os.getloadavg(), flush=True)
os._exit(0)
