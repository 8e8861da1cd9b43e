"""GPU diagnostic: parity numbers and per-call timings of our rasterizer vs the reference CUDA build.
Not a test and not the bench — a quick look for development (run under gpurun)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gs_icp_slam_b200 import rasterizer as R  # noqa: E402
from oracle import ref_cuda  # noqa: E402
from tests.util import psnr, rel_err, scene_tensors  # noqa: E402

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0))


def ev_time(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


for P, size, deg in [(20000, (320, 240), 0), (100000, (640, 480), 0), (300000, (640, 480), 0), (1000000, (1280, 960), 0)]:
    scale = 2.0 if P >= 1000000 else 1.0
    g, cm, t, c, cam = scene_tensors(P, 3, dev, sh_degree=deg, size=size, scale=scale)
    W, H = size
    bg = torch.zeros(3, device=dev)
    e = torch.Tensor([])
    fw = lambda: R.rasterize_gaussians(bg, t["means3D"], e, t["opacities"], t["scales"], t["rotations"], 1.0, e,
                                       c["viewmatrix"], c["projmatrix"], c["tanfovx"], c["tanfovy"], H, W, t["shs"], deg,
                                       c["campos"], False, False)
    n, depth, color, radii, is_used, geom, binning, img = fw()
    ref = ref_cuda.RefRaster(bg, t["means3D"], t["shs"], None, t["opacities"].reshape(-1), t["scales"], t["rotations"], None,
                             c["viewmatrix"], c["projmatrix"], c["campos"], c["tanfovx"], c["tanfovy"], H, W, deg)
    print(f"--- P={P} {W}x{H} deg={deg}: R ours={n} ref={ref.num_rendered}  V={int((radii>0).sum())}")
    print("  radii mismatches:", int((radii != ref.radii).sum()), " is_used mismatches:", int((is_used != ref.is_used).sum()))
    if n == ref.num_rendered:
        pl, rg = R.export_binning(n, H, W, binning, img)
        rpl, rrg = ref.export()
        print("  point_list mismatches:", int((pl != rpl).sum()), " ranges mismatches:", int((rg != rrg).sum()))
    col, rcol = color.cpu().numpy(), ref.color.cpu().numpy()
    dep, rdep = depth.cpu().numpy(), ref.depth.cpu().numpy()
    print(f"  color L1 {np.abs(col-rcol).mean():.3e} max {np.abs(col-rcol).max():.3e} psnr {psnr(col, rcol):.1f}  depth L1 {np.abs(dep-rdep).mean():.3e} max {np.abs(dep-rdep).max():.3e}")
    gen = torch.Generator(device="cpu").manual_seed(5)
    gcol = torch.randn((3, H, W), generator=gen).to(dev)
    gdep = torch.randn((1, H, W), generator=gen).to(dev)
    bw = lambda: R.rasterize_gaussians_backward(bg, t["means3D"], radii, e, t["scales"], t["rotations"], 1.0, e,
                                                c["viewmatrix"], c["projmatrix"], c["tanfovx"], c["tanfovy"], gdep, gcol,
                                                t["shs"], deg, c["campos"], geom, n, binning, img, False)
    ours = bw()
    rgrad = ref.backward(gcol, gdep)
    for name, o in zip(["means2D", "colors", "opacity", "means3D", "cov3D", "sh", "scales", "rotations"], ours):
        print(f"  grad {name:9s} rel err {rel_err(o.cpu().numpy(), rgrad[name].cpu().numpy()):.3e}  max|ref| {float(rgrad[name].abs().max()):.3e}")
    t_fw, t_bw = ev_time(fw), ev_time(bw)

    def ref_fw():
        r = ref_cuda.RefRaster(bg, t["means3D"], t["shs"], None, t["opacities"].reshape(-1), t["scales"], t["rotations"],
                               None, c["viewmatrix"], c["projmatrix"], c["campos"], c["tanfovx"], c["tanfovy"], H, W, deg)
        r.free()

    t0 = time.time()
    for _ in range(5):
        ref_fw()
    t_ref_fw = (time.time() - t0) / 5 * 1e3
    t0 = time.time()
    for _ in range(5):
        ref.backward(gcol, gdep)
    t_ref_bw = (time.time() - t0) / 5 * 1e3
    print(f"  time ms: ours fwd {t_fw:.3f} bwd {t_bw:.3f} | reference (wall, incl. malloc+sync) fwd {t_ref_fw:.3f} bwd {t_ref_bw:.3f}")
    ref.free()
