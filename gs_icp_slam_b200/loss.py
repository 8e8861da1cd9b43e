"""Fused mapping loss: the caller side of the rasterizer in the reference's mapper.

Replaces, with two CUDA kernels (csrc/loss.cu) instead of ~50 PyTorch launches per iteration:
  utils/loss_utils.py:17-20   l1_loss(network_output, gt) -> (map, mean), zero where gt == 0
  utils/loss_utils.py:38-69   ssim(img, gt) -> (map, mean), img := where(gt != 0, img, 0), 11x11 Gaussian window
  mp_Mapper.py:225-242        loss = (1 - lambda_dssim) * Ll1 + lambda_dssim * (1 - ssim) + 0.1 * l1(depth / 10, gt_depth / 10)

`mapping_loss` is the fused form of the mapper's whole loss; `ssim` and `l1_loss` keep the reference functions'
signatures and return values so that `utils/loss_utils.py` can forward to them unchanged.  CUDA tensors only (no CPU path).
"""
import ctypes as C

import torch

from ._lib import check, lib


def _prep(t, shape_tail, name):
    if not t.is_cuda:
        raise RuntimeError(f"gs_icp_slam_b200.loss: {name} must be a CUDA tensor (no CPU fallback)")
    if t.dim() == 2:
        t = t.unsqueeze(0)
    if t.dim() != 3 or t.shape[0] != shape_tail:
        raise RuntimeError(f"{name} must have shape ({shape_tail}, H, W), got {tuple(t.shape)}")
    if t.dtype is not torch.float32:
        t = t.float()
    return t.contiguous()


class _MappingLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, depth, gt_image, gt_depth, lambda_dssim, depth_weight, d_max, mask_by_depth, want_map):
        img, dep = _prep(image, 3, "image"), _prep(depth, 1, "depth")
        gti, gtd = _prep(gt_image, 3, "gt_image"), _prep(gt_depth, 1, "gt_depth")
        H, W = img.shape[1], img.shape[2]
        if dep.shape[1:] != (H, W) or gti.shape != img.shape or gtd.shape[1:] != (H, W):
            raise RuntimeError("image / depth / ground-truth shapes do not match")
        dev = img.device
        work = torch.empty(int(lib.gsicp_mapping_loss_work_bytes(H, W)), dtype=torch.uint8, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        parts = torch.empty(3, dtype=torch.float32, device=dev)
        smap = torch.empty_like(img) if want_map else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        check(lib.gsicp_mapping_loss_forward(H, W, img.data_ptr(), dep.data_ptr(), gti.data_ptr(), gtd.data_ptr(),
                                             float(lambda_dssim), float(depth_weight), float(d_max), int(bool(mask_by_depth)),
                                             loss.data_ptr(), parts.data_ptr(), smap.data_ptr() if want_map else None,
                                             work.data_ptr(), stream),
              "gsicp_mapping_loss_forward")
        ctx.save_for_backward(img, dep, gti, gtd, work)
        ctx.cfg = (H, W, float(lambda_dssim), float(depth_weight), float(d_max), int(bool(mask_by_depth)))
        ctx.in_shapes = (image.shape, depth.shape)
        ctx.set_materialize_grads(False)  # no zero tensors for the non-differentiable outputs
        if want_map:
            ctx.mark_non_differentiable(parts, smap)
            return loss, parts, smap
        ctx.mark_non_differentiable(parts)
        return loss, parts

    @staticmethod
    def backward(ctx, grad_loss, *_unused):
        if grad_loss is None:
            return (None,) * 9
        img, dep, gti, gtd, work = ctx.saved_tensors
        H, W, lam, dw, dmax, mbd = ctx.cfg
        g_img, g_dep = torch.empty_like(img), torch.empty_like(dep)
        gl = grad_loss
        if gl.dtype is not torch.float32 or gl.device != img.device or not gl.is_contiguous():
            gl = gl.to(device=img.device, dtype=torch.float32).contiguous()
        stream = torch.cuda.current_stream(img.device).cuda_stream
        check(lib.gsicp_mapping_loss_backward(H, W, img.data_ptr(), dep.data_ptr(), gti.data_ptr(), gtd.data_ptr(), lam, dw, dmax,
                                              mbd, gl.data_ptr(), work.data_ptr(), g_img.data_ptr(), g_dep.data_ptr(), stream),
              "gsicp_mapping_loss_backward")
        return g_img.view(ctx.in_shapes[0]), g_dep.view(ctx.in_shapes[1]), None, None, None, None, None, None, None


def mapping_loss(image, depth, gt_image, gt_depth, lambda_dssim=0.2, depth_weight=0.1, d_max=10.0, mask_by_depth=True,
                 return_parts=False):
    """The mapper's loss (mp_Mapper.py:225-242) in one forward and one backward kernel.

    image [3,H,W], depth [1,H,W]: rasterizer outputs; gt_image, gt_depth: the keyframe's RGB-D.  mask_by_depth applies the
    reference's `gt_image = gt_image * (gt_depth > 0)` first.  Returns the scalar loss (differentiable w.r.t. image and depth);
    with return_parts also the non-differentiable tensor [Ll1, ssim, Ll1_depth]."""
    loss, parts = _MappingLoss.apply(image, depth, gt_image, gt_depth, lambda_dssim, depth_weight, d_max, mask_by_depth, False)
    return (loss, parts) if return_parts else loss


def ssim(img, gt, window_size=11, size_average=True):
    """utils/loss_utils.py:38-69 — returns (ssim_map[3,H,W], mean); differentiable w.r.t. img through the mean only (the
    reference's mapper never back-propagates through the map)."""
    if window_size != 11 or not size_average:
        raise NotImplementedError("gs_icp_slam_b200.loss.ssim: the fused kernel implements window_size=11, size_average=True")
    zero = torch.zeros((1,) + tuple(img.shape[-2:]), dtype=torch.float32, device=img.device)
    # lambda = 1, no depth term: loss = 1 - mean(ssim)  =>  mean = 1 - loss
    loss, _parts, smap = _MappingLoss.apply(img, zero, gt, zero, 1.0, 0.0, 1.0, False, True)
    return smap, 1.0 - loss


def l1_loss(network_output, gt):
    """utils/loss_utils.py:17-20 — returns (map, mean); three element-wise PyTorch ops, kept for signature parity."""
    loss = torch.abs(network_output - gt)
    loss = torch.where(gt != 0, loss, 0.)
    return loss, loss.mean()
