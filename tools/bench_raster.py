"""Device time of the rasterizer kernels for each render-backward variant on the C3 frame (and C4 with `c4`):
    python tools/bench_raster.py [c3|c4]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer  # noqa: E402
from gs_icp_slam_b200 import _lib  # noqa: E402
from gs_icp_slam_b200 import synthetic as S  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
dev = torch.device("cuda:0")
if which == "c4":
    W, H, P, sc, seed = 1280, 960, 1000000, 2.0, 4
else:
    W, H, P, sc, seed = 640, 480, 300000, 1.0, 3
cam = dict(S.TUM)
cam.update(W=W, H=H, fx=cam["fx"] * sc, fy=cam["fy"] * sc, cx=cam["cx"] * sc, cy=cam["cy"] * sc)
g = S.gaussian_map(P, seed, scale=sc)
cm = S.camera_matrices(S.trajectory_pose(3, 200 if which == "c3" else 20, scale=sc), cam)
t = {k: torch.from_numpy(v).to(dev).requires_grad_(True) for k, v in g.items()}
c = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in cm.items()}
m2 = torch.zeros_like(t["means3D"], requires_grad=True)
gen = torch.Generator(device="cpu").manual_seed(5)
gcol, gdep = torch.randn((3, H, W), generator=gen).to(dev), torch.randn((1, H, W), generator=gen).to(dev)
rs = GaussianRasterizationSettings(H, W, c["tanfovx"], c["tanfovy"], torch.zeros(3, device=dev), 1.0, c["viewmatrix"],
                                   c["projmatrix"], 0, c["campos"], False, False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def it():
    depth, color, radii, used = GaussianRasterizer(rs)(means3D=t["means3D"], means2D=m2, opacities=t["opacities"], shs=t["shs"],
                                                       scales=t["scales"], rotations=t["rotations"])
    ((color * gcol).sum() + (depth * gdep).sum()).backward()
    for k in t:
        t[k].grad = None
    m2.grad = None


for v in (0, 1):
    _lib.lib.gsicp_test_set_bwd_variant(v)
    for _ in range(3):
        it()
    torch.cuda.synchronize()
    _lib.prof_reset()
    _lib.prof_enable(True)
    for i in range(20):
        flush.fill_(i)
        it()
    torch.cuda.synchronize()
    _lib.prof_enable(False)
    k = {n: round(ms / cnt * 1e3, 1) for n, (ms, cnt) in _lib.prof_read().items() if cnt}
    print(f"{which} variant {v}: render_backward {k.get('render_backward')} us   all: {k}", flush=True)
_lib.lib.gsicp_test_set_bwd_variant(0)
