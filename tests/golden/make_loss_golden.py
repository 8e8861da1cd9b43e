"""Generates tests/golden/loss_ref_small.npz by running the REFERENCE's own utils/loss_utils.py (imported from
/root/reference, CPU, float32) on seeded inputs, combined exactly as mp_Mapper.py:225-242 does.  Run in the build container
(the reference tree is not present on the GPU box); the fixture is committed.
    python tests/golden/make_loss_golden.py"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/utils/loss_utils.py"


def inputs(H, W, seed):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand((3, H, W), generator=g)
    gt_depth = torch.rand((1, H, W), generator=g) * 4.0
    gt_depth[:, : H // 5, : W // 3] = 0.0                      # invalid depth -> masked colour
    gt[:, H // 2, :] = 0.0                                     # exact zeros in the colour target as well
    image = (gt + 0.1 * torch.randn((3, H, W), generator=g)).clamp(0, 1)
    image[0, 3, 4] = gt[0, 3, 4]                               # a zero residual (sign(0) = 0)
    depth = gt_depth + 0.05 * torch.randn((1, H, W), generator=g)
    return image, depth, gt, gt_depth


def main():
    spec = importlib.util.spec_from_file_location("ref_loss_utils", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    for name, (H, W, seed) in {"a": (37, 53, 1), "b": (64, 48, 2)}.items():
        image, depth, gt, gt_depth = inputs(H, W, seed)
        image.requires_grad_(True)
        depth.requires_grad_(True)
        mask = (gt_depth > 0.).detach()
        gtm = gt * mask
        _, Ll1 = ref.l1_loss(image, gtm)
        smap, s = ref.ssim(image, gtm)
        _, Ld = ref.l1_loss(depth / 10., gt_depth / 10.)
        loss = (1.0 - 0.2) * Ll1 + 0.2 * (1.0 - s) + 0.1 * Ld
        loss.backward()
        out.update({f"{name}_shape": np.array([H, W, seed]), f"{name}_loss": loss.detach().numpy(), f"{name}_l1": Ll1.detach().numpy(),
                    f"{name}_ssim": s.detach().numpy(), f"{name}_l1d": Ld.detach().numpy(), f"{name}_ssim_map": smap.detach().numpy(),
                    f"{name}_grad_image": image.grad.numpy(), f"{name}_grad_depth": depth.grad.numpy()})
    np.savez_compressed(os.path.join(HERE, "loss_ref_small.npz"), **out)
    print("wrote loss_ref_small.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    main()
