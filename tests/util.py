"""Shared helpers for the tests: synthetic scenes as torch tensors, comparison metrics."""
import numpy as np
import torch

from gs_icp_slam_b200 import synthetic as S


def scene_tensors(P, seed, device, cam=S.TUM, frame=3, n_frames=20, sh_degree=0, scale=1.0, size=None):
    cam = dict(cam)
    if size is not None:
        W, H = size
        sx, sy = W / cam["W"], H / cam["H"]
        cam.update(W=W, H=H, fx=cam["fx"] * sx, fy=cam["fy"] * sy, cx=cam["cx"] * sx, cy=cam["cy"] * sy)
    g = S.gaussian_map(P, seed, scale=scale, sh_degree=sh_degree)
    cm = S.camera_matrices(S.trajectory_pose(frame, n_frames, scale=scale), cam)
    t = {k: torch.from_numpy(v).to(device) for k, v in g.items()}
    c = {k: (torch.from_numpy(v).to(device) if isinstance(v, np.ndarray) else v) for k, v in cm.items()}
    return g, cm, t, c, cam


def psnr(a, b):
    mse = float(((a - b) ** 2).mean())
    return 100.0 if mse == 0 else 10 * np.log10(1.0 / mse)


def rel_err(a, b):
    """max |a-b| / (max|b| + tiny): scale-aware error for gradient tensors summed by float atomics."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
