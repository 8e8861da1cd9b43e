"""Mapper bookkeeping on the device: the optimizer step, the Gaussian-table maintenance and the
keyframe hand-over to the tracker of scene/gaussian_model.py + mp_Mapper.py, on top of three C-ABI entry points
(csrc/map_table.cu): gsicp_adam_step, gsicp_table_compact, gsicp_trackable_target.  CUDA tensors only (no CPU path).

* `FusedAdam` — drop-in for `torch.optim.Adam(l, lr=0.0, eps=1e-15)` (gaussian_model.py:225): same constructor, same
  `param_groups` / `state[param] = {"step", "exp_avg", "exp_avg_sq"}` layout, so the reference's own `_prune_optimizer`,
  `cat_tensors_to_optimizer` and `replace_tensor_to_optimizer` keep working on it; `step()` is ONE kernel for all groups.
* `compact_rows(mask, tensors)` — `tensor[mask]` for many tensors sharing a row mask (prune_points, gaussian_model.py:428-446).
* `trackable_target(...)` / `GaussianTable.hand_over_to_tracker(reg, th)` — get_trackable_gaussians_tensor
  (gaussian_model.py:205-215) compacted on the device and installed as the GICP target without a D2H copy.
* `GaussianTable` — the subset of the reference's GaussianModel the mapper loop uses (add_from_pcd2_tensor, prune_points,
  training_setup, get_* activations), built on the pieces above.
"""
import ctypes as C

import torch

from ._lib import check, lib

from ._lib import _sig  # noqa: E402  (registers the bindings: tests/test_abi.py checks that every declared symbol is bound)

_vp = C.c_void_p
_sig("gsicp_adam_step", C.c_int, [C.c_int, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_size_t),
                                  C.POINTER(C.c_float), C.c_int, C.c_double, C.c_double, C.c_double, _vp])
_sig("gsicp_table_compact", C.c_longlong, [C.c_int, _vp, C.c_int, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_int), _vp])
_sig("gsicp_trackable_target", C.c_longlong, [C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_float, _vp, _vp, _vp, _vp])

MAX_TENSORS_PER_LAUNCH = 8


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (no weight decay, no amsgrad, maximize=False) with one launch per step."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        # group by (betas, eps, step, device): one launch per distinct combination (the mapper has exactly one)
        buckets = {}
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("gs_icp_slam_b200.map_table.FusedAdam: CUDA parameters only (no CPU fallback)")
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] = st["step"] + 1 if torch.is_tensor(st["step"]) else st["step"] + 1
                step = int(st["step"])
                key = (float(b1), float(b2), float(group["eps"]), step, p.device)
                buckets.setdefault(key, []).append((p, float(group["lr"]), st))
        for (b1, b2, eps, step, dev), items in buckets.items():
            for i in range(0, len(items), MAX_TENSORS_PER_LAUNCH):
                chunk = items[i:i + MAX_TENSORS_PER_LAUNCH]
                n = len(chunk)
                keep = []
                P, G, M, V = (_vp * n)(), (_vp * n)(), (_vp * n)(), (_vp * n)()
                cnt, lrs = (C.c_size_t * n)(), (C.c_float * n)()
                for k, (p, lr, st) in enumerate(chunk):
                    g = p.grad
                    if g.dtype is not torch.float32 or not g.is_contiguous():
                        g = g.float().contiguous()
                    if p.dtype is not torch.float32 or not p.is_contiguous():
                        raise RuntimeError("FusedAdam: float32 contiguous parameters only")
                    m, v = st["exp_avg"], st["exp_avg_sq"]
                    if not m.is_contiguous() or not v.is_contiguous():  # e.g. after boolean-mask pruning of a view
                        m, v = m.contiguous(), v.contiguous()
                        st["exp_avg"], st["exp_avg_sq"] = m, v
                    keep.append(g)
                    P[k], G[k], M[k], V[k] = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()
                    cnt[k], lrs[k] = p.numel(), lr
                with torch.cuda.device(dev):
                    check(lib.gsicp_adam_step(n, P, G, M, V, cnt, lrs, step, b1, b2, eps, _stream(dev)), "gsicp_adam_step")
                del keep
        return loss


def compact_rows(mask, tensors):
    """[t[mask] for t in tensors] for tensors sharing dim 0, with one scan and one scatter launch.  Returns new tensors."""
    if not tensors:
        return []
    dev = tensors[0].device
    rows = tensors[0].shape[0]
    if mask.shape[0] != rows:
        raise RuntimeError("mask length does not match the tensors")
    keep = mask.to(device=dev, dtype=torch.bool).contiguous()
    out, src = [], []
    for t in tensors:
        if t.shape[0] != rows or t.device != dev:
            raise RuntimeError("tensors must share dim 0 and the device")
        t = t.detach()
        src.append(t if t.is_contiguous() else t.contiguous())
        out.append(torch.empty_like(src[-1]))
    # zero-width arrays (f_rest at sh_degree 0 is [n, 0, 3]) carry no bytes: only their row count changes
    live = [k for k, t in enumerate(src) if rows and t.numel() // rows > 0]
    if not live or rows == 0:
        count = int(keep.sum().item())
        return [o[:count] for o in out]
    n = len(live)
    count = 0
    with torch.cuda.device(dev):
        for i in range(0, n, 40):
            m = min(40, n - i)
            S, D, RB = (_vp * m)(), (_vp * m)(), (C.c_int * m)()
            for k in range(m):
                t = src[live[i + k]]
                S[k], D[k] = t.data_ptr(), out[live[i + k]].data_ptr()
                RB[k] = (t.numel() // rows) * t.element_size()
            count = int(check(lib.gsicp_table_compact(rows, keep.data_ptr(), m, S, D, RB, _stream(dev)), "gsicp_table_compact"))
    return [o[:count] for o in out]


def trackable_target(xyz, rotation_raw, scaling_raw, opacity_raw, trackable_mask, opacity_th):
    """get_trackable_gaussians_tensor on the device: (points [n,3], rotations xyzw normalised [n,4], scales [n,3]) as CUDA
    tensors (the reference returns them on the CPU)."""
    dev = xyz.device
    P = xyz.shape[0]
    x, r, s, o = (t.detach().float().contiguous() for t in (xyz, rotation_raw, scaling_raw, opacity_raw))
    tm = trackable_mask.to(device=dev, dtype=torch.bool).contiguous()
    ox, orr, osc = torch.empty_like(x), torch.empty_like(r), torch.empty_like(s)
    with torch.cuda.device(dev):
        n = int(check(lib.gsicp_trackable_target(P, x.data_ptr(), r.data_ptr(), s.data_ptr(), o.data_ptr(), tm.data_ptr(),
                                                 float(opacity_th), ox.data_ptr(), orr.data_ptr(), osc.data_ptr(), _stream(dev)),
                      "gsicp_trackable_target"))
    return ox[:n], orr[:n], osc[:n]


def _inverse_sigmoid(x):
    return torch.log(x / (1 - x))


C0 = 0.28209479177387814


# ---- N4: scene.ply (scene/gaussian_model.py:269-281, 619-636, 350-385) --------------------------------------------------
def ply_attribute_names(n_dc, n_rest, n_scale=3, n_rot=4):
    """construct_list_of_attributes (gaussian_model.py:269-281): x y z nx ny nz f_dc_* f_rest_* opacity scale_* rot_*."""
    return (["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(n_dc)] + [f"f_rest_{i}" for i in range(n_rest)] +
            ["opacity"] + [f"scale_{i}" for i in range(n_scale)] + [f"rot_{i}" for i in range(n_rot)])


def write_scene_ply(path, xyz, features_dc, features_rest, opacity, scaling, rotation):
    """The reference's GaussianModel.save_ply (gaussian_model.py:619-636): one `vertex` element of float32 properties,
    binary little endian.  The attribute table is assembled on the tensors' device ([N, C] in one concatenation, features
    channel-major like the reference's transpose(1, 2).flatten) and leaves it in ONE copy; the reference builds N Python
    tuples.  Same bytes as plyfile writes for that element."""
    import os

    import numpy as np

    n = xyz.shape[0]
    dc = features_dc.detach().transpose(1, 2).reshape(n, features_dc.shape[1] * features_dc.shape[2])
    rest = features_rest.detach().transpose(1, 2).reshape(n, features_rest.shape[1] * features_rest.shape[2])
    table = torch.cat((xyz.detach(), torch.zeros_like(xyz), dc, rest, opacity.detach().reshape(n, 1), scaling.detach(),
                       rotation.detach()), dim=1).to(torch.float32).contiguous()
    names = ply_attribute_names(dc.shape[1], rest.shape[1], scaling.shape[1], rotation.shape[1])
    assert table.shape[1] == len(names)
    head = ["ply", "format binary_little_endian 1.0", f"element vertex {n}"] + [f"property float {a}" for a in names] + ["end_header"]
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    host = table.cpu().numpy().astype("<f4", copy=False)
    with open(path, "wb") as f:
        f.write(("\n".join(head) + "\n").encode("ascii"))
        f.write(np.ascontiguousarray(host).tobytes())
    return n


def read_scene_ply(path, max_sh_degree):
    """The reference's load_ply (gaussian_model.py:350-385): returns float32 tensors (CPU) in the model's layout:
    xyz [N,3], features_dc [N,1,3], features_rest [N,(D+1)^2-1,3], opacity [N,1], scaling [N,3], rotation [N,4]."""
    import numpy as np

    with open(path, "rb") as f:
        blob = f.read()
    end = blob.index(b"end_header\n") + len(b"end_header\n")
    lines = blob[:end].decode("ascii").split("\n")
    if lines[0] != "ply" or not lines[1].startswith("format binary_little_endian"):
        raise ValueError("not a binary little-endian ply file")
    n = next(int(l.split()[2]) for l in lines if l.startswith("element vertex"))
    props = [l.split() for l in lines if l.startswith("property ")]
    if any(p[1] != "float" for p in props):
        raise ValueError("scene.ply holds float properties only")
    names = [p[2] for p in props]
    tab = np.frombuffer(blob, dtype="<f4", count=n * len(names), offset=end).reshape(n, len(names))
    col = {a: i for i, a in enumerate(names)}
    pick = lambda prefix: sorted((a for a in names if a.startswith(prefix)), key=lambda a: int(a.split("_")[-1]))
    t = lambda cols: torch.from_numpy(np.ascontiguousarray(tab[:, [col[c] for c in cols]]))
    m = (max_sh_degree + 1) ** 2
    rest = pick("f_rest_")
    if len(rest) != 3 * (m - 1):
        raise ValueError(f"f_rest has {len(rest)} columns, sh degree {max_sh_degree} needs {3 * (m - 1)}")
    return {"xyz": t(["x", "y", "z"]), "f_dc": t(pick("f_dc_")).reshape(n, 3, 1).transpose(1, 2).contiguous(),
            "f_rest": t(rest).reshape(n, 3, m - 1).transpose(1, 2).contiguous(), "opacity": t(["opacity"]),
            "scaling": t(pick("scale_")), "rotation": t(pick("rot_"))}


class GaussianTable:
    """The part of the reference's GaussianModel that the mapper loop drives (scene/gaussian_model.py), with the fused
    optimizer step, one-pass pruning and the on-device target hand-over."""

    GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")

    def __init__(self, sh_degree, device="cuda"):
        self.max_sh_degree, self.active_sh_degree = sh_degree, 0
        self.device = torch.device(device)
        e = lambda *s: torch.empty(s, device=self.device)
        m = (sh_degree + 1) ** 2
        self._xyz, self._features_dc, self._features_rest = e(0, 3), e(0, 1, 3), e(0, m - 1, 3)
        self._opacity, self._scaling, self._rotation = e(0, 1), e(0, 3), e(0, 4)
        self.trackable_mask = torch.empty(0, dtype=torch.bool, device=self.device)
        self.max_radii2D, self.xyz_gradient_accum, self.denom = e(0), e(0, 1), e(0, 1)
        self.optimizer = None
        self._lr = {}

    # ---- activations (gaussian_model.py:32-46, 96-130) ----
    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_scaling(self):
        return torch.exp(self._scaling)

    @property
    def get_rotation(self):
        return torch.nn.functional.normalize(self._rotation)

    @property
    def get_opacity(self):
        return torch.sigmoid(self._opacity)

    @property
    def get_features(self):
        return torch.cat((self._features_dc, self._features_rest), dim=1)

    def save_ply(self, path):
        """gaussian_model.py:619-636."""
        return write_scene_ply(path, self._xyz, self._features_dc, self._features_rest, self._opacity, self._scaling, self._rotation)

    def load_ply(self, path):
        """gaussian_model.py:350-385: parameters from a scene.ply; the active SH degree becomes the maximum one."""
        d = {k: v.to(self.device) for k, v in read_scene_ply(path, self.max_sh_degree).items()}
        self._set(d)
        self.active_sh_degree = self.max_sh_degree
        n = self._xyz.shape[0]
        self.trackable_mask = torch.zeros(n, dtype=torch.bool, device=self.device)
        self.max_radii2D = torch.zeros(n, device=self.device)

    def params(self):
        return {"xyz": self._xyz, "f_dc": self._features_dc, "f_rest": self._features_rest, "opacity": self._opacity,
                "scaling": self._scaling, "rotation": self._rotation}

    def _set(self, d):
        (self._xyz, self._features_dc, self._features_rest, self._opacity, self._scaling, self._rotation) = (d[k] for k in self.GROUPS)

    def training_setup(self, position_lr=0.00016, feature_lr=0.0025, opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001,
                       spatial_lr_scale=1.0):
        """gaussian_model.py:217-231: six groups, Adam(lr=0, eps=1e-15)."""
        self._lr = {"xyz": position_lr * spatial_lr_scale, "f_dc": feature_lr, "f_rest": feature_lr / 20.0, "opacity": opacity_lr,
                    "scaling": scaling_lr, "rotation": rotation_lr}
        d = {k: torch.nn.Parameter(v.detach().clone().requires_grad_(True)) for k, v in self.params().items()}
        self._set(d)
        self.optimizer = FusedAdam([{"params": [d[k]], "lr": self._lr[k], "name": k} for k in self.GROUPS], lr=0.0, eps=1e-15)
        n = self._xyz.shape[0]
        self.xyz_gradient_accum = torch.zeros((n, 1), device=self.device)
        self.denom = torch.zeros((n, 1), device=self.device)

    def _rebuild(self, new_params, new_states):
        """Swap the parameters of the optimizer groups (and their Adam moments), like the reference's optimizer surgery."""
        d = {}
        for group in self.optimizer.param_groups:
            name = group["name"]
            old = group["params"][0]
            st = self.optimizer.state.pop(old, None)
            p = torch.nn.Parameter(new_params[name].requires_grad_(True))
            group["params"][0] = p
            if st is not None:
                st["exp_avg"], st["exp_avg_sq"] = new_states[name]
                self.optimizer.state[p] = st
            d[name] = p
        self._set(d)

    def add_from_pcd2_tensor(self, points, colors, rots_, scales_, z_vals_, trackable_idxs):
        """gaussian_model.py:165-203: append new Gaussians (rotations / scales from GICP, z-dependent shrink), zero moments."""
        dev = self.device
        n_new = points.shape[0]
        m = (self.max_sh_degree + 1) ** 2
        f_dc = ((colors.to(dev).float() - 0.5) / C0).reshape(n_new, 1, 3)
        f_rest = torch.zeros((n_new, m - 1, 3), device=dev)
        z = torch.clamp_min((z_vals_.to(dev).float() ** 1.5) * 2., 1.).unsqueeze(-1).repeat(1, 3)
        new = {"xyz": points.to(dev).float(), "f_dc": f_dc, "f_rest": f_rest,
               "opacity": _inverse_sigmoid(0.1 * torch.ones((n_new, 1), device=dev)),
               "scaling": torch.log(scales_.to(dev).float() / z), "rotation": rots_.to(dev).float()}
        tm = torch.zeros(n_new, dtype=torch.bool, device=dev)
        if len(trackable_idxs) != 0:
            tm[trackable_idxs] = True
        cur = self.params()
        cat = {k: torch.cat((cur[k].detach(), new[k]), dim=0) for k in self.GROUPS}
        if self.optimizer is None:
            self._set(cat)
        else:
            states = {}
            for group in self.optimizer.param_groups:
                st = self.optimizer.state.get(group["params"][0])
                k = group["name"]
                if st is not None:
                    states[k] = (torch.cat((st["exp_avg"], torch.zeros_like(new[k])), dim=0),
                                 torch.cat((st["exp_avg_sq"], torch.zeros_like(new[k])), dim=0))
                else:
                    states[k] = None
            self._rebuild(cat, states)
        n = cat["xyz"].shape[0]
        self.trackable_mask = torch.cat([self.trackable_mask, tm], dim=0)
        self.xyz_gradient_accum = torch.zeros((n, 1), device=dev)
        self.denom = torch.zeros((n, 1), device=dev)
        self.max_radii2D = torch.zeros(n, device=dev)

    def prune_points(self, mask):
        """gaussian_model.py:428-446: drop the rows where mask is True — parameters, Adam moments, accumulators and the
        trackable mask in one scan + one scatter."""
        valid = ~mask
        names, tensors = [], []
        for group in self.optimizer.param_groups:
            p = group["params"][0]
            st = self.optimizer.state.get(p)
            names.append(group["name"])
            tensors.append(p.detach())
            if st is not None:
                tensors += [st["exp_avg"], st["exp_avg_sq"]]
            else:
                tensors += [torch.zeros_like(p), torch.zeros_like(p)]
        tensors += [self.xyz_gradient_accum, self.denom, self.max_radii2D, self.trackable_mask]
        out = compact_rows(valid, tensors)
        new_p = {k: out[3 * i] for i, k in enumerate(names)}
        new_s = {k: (out[3 * i + 1], out[3 * i + 2]) for i, k in enumerate(names)}
        self._rebuild(new_p, new_s)
        self.xyz_gradient_accum, self.denom, self.max_radii2D, self.trackable_mask = out[3 * len(names):]

    def prune_large_and_transparent(self, min_opacity, extent):
        """gaussian_model.py:580-592."""
        prune = (self.get_opacity < min_opacity).squeeze(-1)
        if extent is not None:
            prune = torch.logical_or(prune, self.get_scaling.max(dim=1).values > 0.1 * extent)
        self.prune_points(prune)

    def get_trackable_gaussians_tensor(self, opacity_th):
        """gaussian_model.py:205-215, but the three tensors stay on the device."""
        return trackable_target(self._xyz, self._rotation, self._scaling, self._opacity, self.trackable_mask, opacity_th)

    def hand_over_to_tracker(self, reg, opacity_th):
        """mp_Mapper.py:171-173 + mp_Tracker.py:284-289 without the GPU -> CPU -> shared memory -> numpy round trip: the
        compacted target goes from the table straight into the GICP target buffers (device to device)."""
        pts, rots, scales = self.get_trackable_gaussians_tensor(opacity_th)
        reg.set_input_target(pts)
        reg.set_target_covariances_fromqs(rots, scales)
        return pts.shape[0]
