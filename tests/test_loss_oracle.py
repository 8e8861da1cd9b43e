"""CPU: the loss oracle (oracle/loss_oracle.py, float64) against the golden vectors produced by the reference's own
utils/loss_utils.py (tests/golden/make_loss_golden.py).  Tolerances are float32 round-off of the reference run."""
import os

import numpy as np
import torch

from oracle import loss_oracle as LO
from tests.golden.make_loss_golden import inputs

HERE = os.path.dirname(os.path.abspath(__file__))


def test_loss_oracle_matches_reference_golden():
    gold = np.load(os.path.join(HERE, "golden", "loss_ref_small.npz"))
    for name in ("a", "b"):
        H, W, seed = (int(v) for v in gold[f"{name}_shape"])
        image, depth, gt, gt_depth = (t.double() for t in inputs(H, W, seed))
        image.requires_grad_(True)
        depth.requires_grad_(True)
        loss, l1, s, ld = LO.mapping_loss(image, depth, gt, gt_depth)
        loss.backward()
        assert abs(float(loss.detach()) - float(gold[f"{name}_loss"])) <= 2e-6
        assert abs(float(l1) - float(gold[f"{name}_l1"])) <= 1e-6
        assert abs(float(s) - float(gold[f"{name}_ssim"])) <= 2e-6
        assert abs(float(ld) - float(gold[f"{name}_l1d"])) <= 1e-6
        smap, _ = LO.ssim(image.detach(), gt * (gt_depth > 0))
        assert np.abs(smap.numpy() - gold[f"{name}_ssim_map"]).max() <= 2e-4  # cancellation in sigma = E[x^2] - mu^2 (float32 ref)
        gi, gd = image.grad.numpy(), depth.grad.numpy()
        assert np.abs(gi - gold[f"{name}_grad_image"]).max() <= 2e-4 * np.abs(gold[f"{name}_grad_image"]).max()
        assert np.array_equal(np.sign(gd), np.sign(gold[f"{name}_grad_depth"]))
        assert np.abs(gd - gold[f"{name}_grad_depth"]).max() <= 1e-6 * np.abs(gold[f"{name}_grad_depth"]).max() + 1e-12


def test_loss_oracle_gradient_is_consistent():
    """Finite differences of the float64 oracle (the thing the CUDA kernel is compared with)."""
    image, depth, gt, gt_depth = (t.double() for t in inputs(20, 24, 7))
    image.requires_grad_(True)
    loss = LO.mapping_loss(image, depth, gt, gt_depth)[0]
    (g,) = torch.autograd.grad(loss, image)
    rng = np.random.default_rng(0)
    for _ in range(8):
        c, y, x = rng.integers(0, 3), rng.integers(0, 20), rng.integers(0, 24)
        d = torch.zeros_like(image)
        d[c, y, x] = 1e-6
        lp = LO.mapping_loss(image.detach() + d, depth, gt, gt_depth)[0]
        lm = LO.mapping_loss(image.detach() - d, depth, gt, gt_depth)[0]
        fd = float(lp - lm) / 2e-6
        assert abs(fd - float(g[c, y, x])) <= 1e-5 * max(abs(fd), 1e-4) + 1e-9


def test_bench_reference_arm_loss_is_the_reference_formulation():
    """bench.py's `torch_mapper_loss` (what the reference arm and the ssim_torch variant evaluate) against the golden vectors
    of the reference's own loss code."""
    import sys

    sys.path.insert(0, os.path.dirname(HERE))
    import bench

    gold = np.load(os.path.join(HERE, "golden", "loss_ref_small.npz"))
    for name in ("a", "b"):
        H, W, seed = (int(v) for v in gold[f"{name}_shape"])
        image, depth, gt, gt_depth = inputs(H, W, seed)
        image.requires_grad_(True)
        depth.requires_grad_(True)
        loss = bench.torch_mapper_loss(image, depth, gt, gt_depth)
        loss.backward()
        assert abs(float(loss.detach()) - float(gold[f"{name}_loss"])) <= 1e-6
        assert np.abs(image.grad.numpy() - gold[f"{name}_grad_image"]).max() <= 1e-5 * np.abs(gold[f"{name}_grad_image"]).max()
        assert np.abs(depth.grad.numpy() - gold[f"{name}_grad_depth"]).max() <= 1e-6 * np.abs(gold[f"{name}_grad_depth"]).max() + 1e-12
