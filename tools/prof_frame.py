"""Runs a few SLAM frames of the bench workload (resident mode, back to back) — the command ncu wraps for profiles/.
    python tools/prof_frame.py [frames] [ours|reference]
`reference` drives the reference's own torch extension (oracle/_ref/site) on the same frames: the "kernel to beat"
captures (renderCUDA, preprocessCUDA, duplicateWithKeys, DeviceRadixSort)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
which = sys.argv[2] if len(sys.argv) > 2 else "ours"
cam, max_corr, label, gmap, frames = bench.make_sequence(n + 1, "c3", 300000)
dev = torch.device("cuda:0")
if which == "ours":
    eng = bench.Ours(cam, max_corr, gmap, frames, dev, 1, 0, loss="fused")
    for i in range(n):
        eng.step(i, resident=True)
    torch.cuda.synchronize()
    eng.close()
else:
    from oracle import ref_ext

    dgr = ref_ext.diff_gaussian_rasterization()
    m = {k: torch.from_numpy(v).to(dev).requires_grad_(True) for k, v in gmap.items()}
    m2 = torch.zeros_like(m["means3D"], requires_grad=True)
    bg = torch.zeros(3, device=dev)
    for i in range(n):
        f = frames[bench.frame_index(i + 1, len(frames))]
        c = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in f["cam"].items()}
        rs = dgr.GaussianRasterizationSettings(cam["H"], cam["W"], c["tanfovx"], c["tanfovy"], bg, 1.0, c["viewmatrix"],
                                               c["projmatrix"], 0, c["campos"], False, False)
        depth, color, radii, used = dgr.GaussianRasterizer(rs)(means3D=m["means3D"], means2D=m2, opacities=m["opacities"],
                                                               shs=m["shs"], scales=m["scales"], rotations=m["rotations"])
        loss = bench.torch_mapper_loss(color, depth, torch.from_numpy(f["rgb"]).to(dev), torch.from_numpy(f["depth"]).to(dev))
        loss.backward()
    torch.cuda.synchronize()
print("done", n, which)
