"""Wall-clock breakdown of one SLAM frame (host view): where the non-kernel time goes.  Development tool."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402

n = 24
cam, gmap, frames = bench.make_sequence(n + 1, bench.MAP_P)
eng = bench.Ours(cam, gmap, frames, torch.device("cuda:0"), 1, 0)
acc = {}


def tick(name, t0):
    torch.cuda.synchronize()
    t = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + (t - t0)
    return t


for it in range(n):
    f, prev = frames[it + 1], frames[it]
    torch.cuda.synchronize()
    t = time.perf_counter()
    eng.reg.set_input_source(f["d_pts"]); t = tick("set_input_source", t)
    eng.reg.set_source_filter(f["n_trk"], f["filt"]); t = tick("set_source_filter", t)
    pose = eng.reg.align(prev["c2w"].astype(np.float32)); t = tick("align", t)
    corr, sqd = eng.reg.get_source_correspondence(); t = tick("get_corr", t)
    c, m = f["d_cam"], eng.map
    rs = eng.Settings(cam["H"], cam["W"], c["tanfovx"], c["tanfovy"], eng.bg, 1.0, c["viewmatrix"], c["projmatrix"], 0, c["campos"], False, False)
    depth, color, radii, is_used = eng.Rasterizer(rs)(means3D=m["means3D"], means2D=eng.means2D, opacities=m["opacities"], shs=m["shs"], scales=m["scales"], rotations=m["rotations"])
    t = tick("raster_fwd", t)
    loss = (color - f["d_rgb"]).abs().mean() + 0.1 * (depth - f["d_depth"]).abs().mean(); t = tick("loss_fwd", t)
    loss.backward(); t = tick("backward(loss+raster)", t)
    lv = loss.item(); t = tick("item", t)
    for k in m:
        m[k].grad = None
    eng.means2D.grad = None
    if it == 3:
        acc.clear()
tot = sum(acc.values())
for k, v in acc.items():
    print(f"{k:24s} {v / (n - 4) * 1e3:8.3f} ms")
print(f"{'total':24s} {tot / (n - 4) * 1e3:8.3f} ms   cpu_count={os.cpu_count()} load={os.getloadavg()}")
