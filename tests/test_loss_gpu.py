"""GPU parity of the fused mapping loss (csrc/loss.cu through the C ABI) with the CPU oracle (oracle/loss_oracle.py,
float64) and with the golden vectors of the reference's own utils/loss_utils.py.  Float32 kernel: loss terms within 2e-6
absolute, gradients within 2e-5 of the largest gradient entry."""
import os

import numpy as np
import pytest
import torch

from tests.golden.make_loss_golden import inputs

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(cuda, image, depth, gt, gt_depth, **kw):
    from gs_icp_slam_b200 import loss as L

    im = image.to(cuda).requires_grad_(True)
    dp = depth.to(cuda).requires_grad_(True)
    loss, parts = L.mapping_loss(im, dp, gt.to(cuda), gt_depth.to(cuda), return_parts=True, **kw)
    loss.backward()
    return float(loss.detach()), parts.cpu().numpy(), im.grad.cpu().numpy(), dp.grad.cpu().numpy()


@pytest.mark.parametrize("H,W,seed", [(37, 53, 1), (64, 48, 2), (480, 640, 3), (16, 16, 4), (5, 7, 5)])
def test_mapping_loss_matches_oracle(cuda, H, W, seed):
    from oracle import loss_oracle as LO

    image, depth, gt, gt_depth = inputs(H, W, seed)
    loss, parts, gi, gd = _run(cuda, image, depth, gt, gt_depth)
    i64, d64 = image.double().requires_grad_(True), depth.double().requires_grad_(True)
    o = LO.mapping_loss(i64, d64, gt.double(), gt_depth.double())
    o[0].backward()
    assert abs(loss - float(o[0])) <= 2e-6
    for k in range(3):
        assert abs(float(parts[k]) - float(o[k + 1])) <= 2e-6, k
    ri, rd = i64.grad.numpy(), d64.grad.numpy()
    assert np.abs(gi - ri).max() <= 2e-5 * np.abs(ri).max()
    assert np.abs(gd - rd).max() <= 1e-6 * np.abs(rd).max() + 1e-12
    assert np.array_equal(gi == 0, ri == 0)  # masked pixels get exactly zero gradient


def test_mapping_loss_matches_reference_golden(cuda):
    gold = np.load(os.path.join(HERE, "golden", "loss_ref_small.npz"))
    for name in ("a", "b"):
        H, W, seed = (int(v) for v in gold[f"{name}_shape"])
        loss, parts, gi, gd = _run(cuda, *inputs(H, W, seed))
        assert abs(loss - float(gold[f"{name}_loss"])) <= 3e-6
        assert abs(parts[1] - float(gold[f"{name}_ssim"])) <= 3e-6
        assert np.abs(gi - gold[f"{name}_grad_image"]).max() <= 2e-4 * np.abs(gold[f"{name}_grad_image"]).max()
        assert np.abs(gd - gold[f"{name}_grad_depth"]).max() <= 1e-6 * np.abs(gold[f"{name}_grad_depth"]).max() + 1e-12


def test_ssim_and_options(cuda):
    from gs_icp_slam_b200 import loss as L
    from oracle import loss_oracle as LO

    image, depth, gt, gt_depth = inputs(45, 70, 9)
    smap, mean = L.ssim(image.to(cuda), gt.to(cuda))
    omap, omean = LO.ssim(image.double(), gt.double())
    assert abs(float(mean) - float(omean)) <= 2e-6
    assert np.abs(smap.cpu().numpy() - omap.numpy()).max() <= 2e-4
    # other weights, no depth masking, upstream gradient != 1
    im = image.to(cuda).requires_grad_(True)
    dp = depth.to(cuda).requires_grad_(True)
    loss = L.mapping_loss(im, dp, gt.to(cuda), gt_depth.to(cuda), lambda_dssim=0.35, depth_weight=0.5, d_max=4.0, mask_by_depth=False)
    (3.0 * loss).backward()
    i64, d64 = image.double().requires_grad_(True), depth.double().requires_grad_(True)
    o = LO.mapping_loss(i64, d64, gt.double(), gt_depth.double(), 0.35, 0.5, 4.0, False)[0]
    (3.0 * o).backward()
    assert abs(float(loss) - float(o)) <= 2e-6
    assert np.abs(im.grad.cpu().numpy() - i64.grad.numpy()).max() <= 2e-5 * np.abs(i64.grad.numpy()).max()
    assert np.abs(dp.grad.cpu().numpy() - d64.grad.numpy()).max() <= 1e-6 * np.abs(d64.grad.numpy()).max()
    with pytest.raises(RuntimeError):
        L.mapping_loss(image, depth, gt, gt_depth)  # CPU tensors: no fallback
