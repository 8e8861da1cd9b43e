"""SURVEY.md §8f row N3 on the GPU: fused Adam, one-pass pruning and the on-device tracker hand-over against PyTorch's
own ops driven the way the reference's GaussianModel drives them (scene/gaussian_model.py:217-231 optimizer groups,
:385-446 optimizer surgery on concat / prune, :205-215 trackable target)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LRS = {"xyz": 0.00016, "f_dc": 0.0025, "f_rest": 0.0025 / 20, "opacity": 0.05, "scaling": 0.005, "rotation": 0.001}
SHAPES = {"xyz": (3,), "f_dc": (1, 3), "f_rest": (15, 3), "opacity": (1,), "scaling": (3,), "rotation": (4,)}


def _params(n, dev, gen):
    return {k: torch.randn((n,) + s, generator=gen).to(dev) for k, s in SHAPES.items()}


def _make(optim_cls, init, dev):
    p = {k: torch.nn.Parameter(v.clone().requires_grad_(True)) for k, v in init.items()}
    opt = optim_cls([{"params": [p[k]], "lr": LRS[k], "name": k} for k in SHAPES], lr=0.0, eps=1e-15)
    return p, opt


def _ref_prune(opt, mask):
    """The reference's _prune_optimizer (gaussian_model.py:409-426), restated on torch.optim.Adam."""
    out = {}
    for group in opt.param_groups:
        st = opt.state.get(group["params"][0])
        if st is not None:
            st["exp_avg"], st["exp_avg_sq"] = st["exp_avg"][mask], st["exp_avg_sq"][mask]
            del opt.state[group["params"][0]]
            group["params"][0] = torch.nn.Parameter(group["params"][0][mask].requires_grad_(True))
            opt.state[group["params"][0]] = st
        else:
            group["params"][0] = torch.nn.Parameter(group["params"][0][mask].requires_grad_(True))
        out[group["name"]] = group["params"][0]
    return out


def _ref_cat(opt, ext):
    """cat_tensors_to_optimizer (gaussian_model.py:448-468), restated."""
    out = {}
    for group in opt.param_groups:
        e = ext[group["name"]]
        st = opt.state.get(group["params"][0])
        if st is not None:
            st["exp_avg"] = torch.cat((st["exp_avg"], torch.zeros_like(e)), dim=0)
            st["exp_avg_sq"] = torch.cat((st["exp_avg_sq"], torch.zeros_like(e)), dim=0)
            del opt.state[group["params"][0]]
            group["params"][0] = torch.nn.Parameter(torch.cat((group["params"][0], e), dim=0).requires_grad_(True))
            opt.state[group["params"][0]] = st
        else:
            group["params"][0] = torch.nn.Parameter(torch.cat((group["params"][0], e), dim=0).requires_grad_(True))
        out[group["name"]] = group["params"][0]
    return out


def _rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def test_fused_adam_matches_torch_adam_through_concat_and_prune(cuda):
    from gs_icp_slam_b200.map_table import FusedAdam

    gen = torch.Generator().manual_seed(0)
    init = _params(5000, cuda, gen)
    pa, oa = _make(torch.optim.Adam, init, cuda)
    pb, ob = _make(FusedAdam, init, cuda)
    launches = []
    from gs_icp_slam_b200 import _lib

    for it in range(30):
        n = pa["xyz"].shape[0]
        grads = {k: torch.randn((n,) + s, generator=gen).to(cuda) * (10.0 ** ((it % 5) - 2)) for k, s in SHAPES.items()}
        for p in (pa, pb):
            for k in p:
                p[k].grad = grads[k].clone()
        oa.step()
        l0 = _lib.launch_count()
        ob.step()
        launches.append(_lib.launch_count() - l0)
        if it == 9:  # new keyframe: concat 700 Gaussians with zero moments
            ext = _params(700, cuda, gen)
            pa, pb = _ref_cat(oa, ext), _ref_cat(ob, {k: v.clone() for k, v in ext.items()})
        if it == 19:  # pruning
            mask = (torch.rand(pa["xyz"].shape[0], generator=gen) > 0.3).to(cuda)
            pa, pb = _ref_prune(oa, mask), _ref_prune(ob, mask)
    assert set(launches) == {1}, launches  # six groups, one kernel per step
    for k in SHAPES:
        assert _rel(pb[k].detach(), pa[k].detach()) <= 1e-6, k
        sa, sb = oa.state[pa[k]], ob.state[pb[k]]
        assert _rel(sb["exp_avg"], sa["exp_avg"]) <= 1e-6 and _rel(sb["exp_avg_sq"], sa["exp_avg_sq"]) <= 1e-6, k
        assert int(sa["step"]) == int(sb["step"]) == 30


def test_compact_rows_equals_boolean_indexing(cuda):
    from gs_icp_slam_b200.map_table import compact_rows

    gen = torch.Generator().manual_seed(1)
    for n in (1, 37, 100003):
        ts = [torch.randn((n, 3), generator=gen).to(cuda), torch.randn((n, 15, 3), generator=gen).to(cuda),
              torch.randn((n,), generator=gen).to(cuda), (torch.rand(n, generator=gen) > 0.5).to(cuda),
              torch.randint(0, 255, (n, 5), generator=gen, dtype=torch.uint8).to(cuda)]
        for frac in (0.0, 0.4, 1.0):
            mask = (torch.rand(n, generator=gen) < frac).to(cuda)
            out = compact_rows(mask, ts)
            for o, t in zip(out, ts):
                assert o.dtype == t.dtype and torch.equal(o, t[mask])


def test_table_matches_reference_bookkeeping_and_hands_over_on_device(cuda):
    """GaussianTable through keyframe insertion, training steps, pruning; then the tracker target straight from the table."""
    import pygicp
    from gs_icp_slam_b200 import synthetic as S
    from gs_icp_slam_b200.map_table import GaussianTable, trackable_target

    g = S.gaussian_map(30000, 11)
    tab = GaussianTable(0, cuda)
    pts = torch.from_numpy(g["means3D"]).to(cuda)
    cols = torch.rand(len(pts), 3, device=cuda)
    rots = torch.from_numpy(g["rotations"]).to(cuda)
    scl = torch.exp(torch.from_numpy(g["scales"]).to(cuda)) if g["scales"].min() < 0 else torch.from_numpy(g["scales"]).to(cuda)
    z = torch.full((len(pts),), 1.5, device=cuda)
    tab.add_from_pcd2_tensor(pts, cols, rots, scl, z, torch.arange(0, len(pts), 2, device=cuda))
    tab.training_setup()
    gen = torch.Generator().manual_seed(2)
    for it in range(5):
        for k, p in tab.params().items():
            p.grad = torch.randn(p.shape, generator=gen).to(cuda) * 1e-2
        tab.optimizer.step()
        tab.optimizer.zero_grad(set_to_none=True)
    n0 = tab.get_xyz.shape[0]
    tab.add_from_pcd2_tensor(pts[:1000] + 0.01, cols[:1000], rots[:1000], scl[:1000], z[:1000], torch.arange(0, 1000, device=cuda))
    assert tab.get_xyz.shape[0] == n0 + 1000 and tab.trackable_mask.shape[0] == n0 + 1000
    with torch.no_grad():
        tab._opacity[::7] = -10.0  # nearly transparent
    before = {k: v.detach().clone() for k, v in tab.params().items()}
    prune = (tab.get_opacity < 0.005).squeeze(-1)
    tm_before = tab.trackable_mask.clone()
    tab.prune_large_and_transparent(0.005, None)
    for k, v in tab.params().items():
        assert torch.equal(v.detach(), before[k][~prune]), k
    assert torch.equal(tab.trackable_mask, tm_before[~prune])
    # trackable target: device result == the reference's expressions
    th = 0.09
    tp, tr, ts = tab.get_trackable_gaussians_tensor(th)
    with torch.no_grad():
        sel = torch.logical_and((tab.get_opacity > th).squeeze(-1), tab.trackable_mask)
        assert tp.is_cuda and torch.equal(tp, tab.get_xyz[sel])
        # normalize / exp evaluated in one kernel instead of torch's reduction + division kernels: 2 ulp of a unit quaternion
        assert torch.allclose(tr, tab.get_rotation[sel], rtol=0, atol=2.5e-7), float((tr - tab.get_rotation[sel]).abs().max())
        assert torch.allclose(ts, tab.get_scaling[sel], rtol=2e-6, atol=0), float((ts / tab.get_scaling[sel] - 1).abs().max())
    # hand-over: same registration result as feeding the tracker the reference's CPU copies
    cam = S.TUM
    pose = S.trajectory_pose(3, 200)
    src, trk = S.tracker_cloud(S.raycast_depth(pose, cam)[0], cam)
    res = []
    for device_path in (True, False):
        reg = pygicp.FastGICP()
        reg.set_max_correspondence_distance(0.03)
        reg.set_max_knn_distance(99999)
        if device_path:
            n_t = tab.hand_over_to_tracker(reg, th)
            assert n_t == tp.shape[0]
        else:
            reg.set_input_target(tp.cpu().numpy())
            reg.set_target_covariances_fromqs(tr.cpu().numpy().reshape(-1), ts.cpu().numpy().reshape(-1))
        reg.set_input_source(src)
        reg.set_source_filter(len(trk), S.trackable_filter(len(src), trk))
        res.append((reg.align(pose.astype(np.float32)), reg.get_source_correspondence()))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1][0], res[1][1][0])
