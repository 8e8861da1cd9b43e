"""CPU oracle of the mapper's loss — TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product).

Plain PyTorch restatement, in float64 on the CPU, of
  utils/loss_utils.py:17-20   l1_loss
  utils/loss_utils.py:22-36   gaussian / create_window   (window taps rounded to float32 like the reference's torch.Tensor)
  utils/loss_utils.py:38-69   ssim / _ssim
  mp_Mapper.py:225-242        mask, loss_rgb, loss_d, total
Pinned by tests/test_loss_oracle.py against tests/golden/loss_ref_small.npz, which tests/golden/make_loss_golden.py generates
by importing the reference's own utils/loss_utils.py from /root/reference (CPU, float32)."""
from math import exp

import torch
import torch.nn.functional as F


def l1_loss(network_output, gt):
    loss = torch.abs(network_output - gt)
    loss = torch.where(gt != 0, loss, torch.zeros_like(loss))
    return loss, loss.mean()


def window(channel, dtype=torch.float64, window_size=11, sigma=1.5):
    g = torch.tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)], dtype=torch.float32)
    g = (g / g.sum()).unsqueeze(1)            # float32, as in the reference
    w2 = g.mm(g.t()).float().to(dtype)        # the 2-D window is formed in float32 (loss_utils.py:33)
    return w2.unsqueeze(0).unsqueeze(0).expand(channel, 1, window_size, window_size).contiguous()


def ssim(img, gt, window_size=11):
    img = torch.where(gt != 0, img, torch.zeros_like(img))
    ch = img.size(-3)
    w = window(ch, img.dtype, window_size)
    x, y = img.unsqueeze(0), gt.unsqueeze(0)
    p = window_size // 2
    mu1, mu2 = F.conv2d(x, w, padding=p, groups=ch), F.conv2d(y, w, padding=p, groups=ch)
    mu1_sq, mu2_sq, mu12 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(x * x, w, padding=p, groups=ch) - mu1_sq
    s2 = F.conv2d(y * y, w, padding=p, groups=ch) - mu2_sq
    s12 = F.conv2d(x * y, w, padding=p, groups=ch) - mu12
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu12 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return m.squeeze(0), m.mean()


def mapping_loss(image, depth, gt_image, gt_depth, lambda_dssim=0.2, depth_weight=0.1, d_max=10.0, mask_by_depth=True):
    """Returns (loss, Ll1, ssim, Ll1_depth) as tensors (differentiable w.r.t. image and depth)."""
    if mask_by_depth:
        gt_image = gt_image * (gt_depth > 0.)
    _, Ll1 = l1_loss(image, gt_image)
    _, s = ssim(image, gt_image)
    _, Ld = l1_loss(depth / d_max, gt_depth / d_max)
    loss = (1.0 - lambda_dssim) * Ll1 + lambda_dssim * (1.0 - s) + depth_weight * Ld
    return loss, Ll1, s, Ld
