"""distCUDA2 on the C ABI: mean squared distance to the 3 nearest other points (simple-knn/spatial.cu:15-25)."""
import torch

from ._lib import check, lib


def distCUDA2(points):
    if not points.is_cuda:
        raise RuntimeError("distCUDA2: points must be a CUDA tensor (no CPU fallback)")
    p = points.detach().float().contiguous()
    if p.ndimension() != 2 or p.size(1) != 3:
        raise RuntimeError("distCUDA2: points must have shape (P, 3)")
    out = torch.zeros((p.size(0),), dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        check(lib.gsicp_dist2(p.size(0), p.data_ptr(), out.data_ptr(), torch.cuda.current_stream(p.device).cuda_stream),
              "gsicp_dist2")
    return out
