"""SURVEY.md §8f row N1 on the GPU: the tracker front-end kernels against the reference's own expressions
(mp_Tracker.py:394-431 set_downsample_filter + downsample_and_make_pointcloud2, :224-229 world transform, :385-392
quaternion_multiply, :374-380 eliminate_overlapped2), restated here with the same torch / numpy / scipy calls."""
import numpy as np
import pytest
import torch

from gs_icp_slam_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def _reference_cloud(depth_img, rgb_img, cam):
    W, H, s = cam["W"], cam["H"], cam["downsample"]
    h_val = s * torch.arange(0, int(H / s) + 1) - 1
    h_val[0] = 0
    h_val = h_val * W
    a, b = torch.meshgrid(h_val, torch.arange(0, W, s), indexing="ij")
    pick = ((a + b).flatten(),)
    v, u = torch.meshgrid(torch.arange(0, H), torch.arange(0, W), indexing="ij")
    u, v = u.flatten()[pick], v.flatten()[pick]
    x_pre, y_pre = (u - cam["cx"]) / cam["fx"], (v - cam["cy"]) / cam["fy"]
    colors = torch.from_numpy(rgb_img).reshape(-1, 3).float()[pick] / 255
    z = torch.from_numpy(depth_img.astype(np.float32)).flatten()[pick] / cam["depth_scale"]
    zero = torch.where(z != 0)
    filt = torch.where(z[zero] <= cam["depth_trunc"])
    z = z[zero]
    pts = torch.stack([x_pre[zero] * z, y_pre[zero] * z, z], dim=-1)
    return pts.numpy(), colors[zero].numpy(), z.numpy(), filt[0].numpy()


def test_cloud_matches_reference_expressions_and_feeds_the_tracker(cuda):
    import pygicp
    from gs_icp_slam_b200.frontend import TrackerFrontEnd

    cam = dict(S.TUM)
    pose = S.trajectory_pose(4, 200)
    depth, hit = S.raycast_depth(pose, cam)
    d16 = np.clip(np.round(depth * cam["depth_scale"]), 0, 65535).astype(np.uint16)
    d16[100:140, 200:260] = 0           # holes: dropped, raster order kept
    d16[300:330, :50] = 20000           # beyond depth_trunc: kept but untrackable
    rgb = np.clip(np.round(S.texture(hit).reshape(cam["H"], cam["W"], 3) * 255), 0, 255).astype(np.uint8)
    fe = TrackerFrontEnd(cam["W"], cam["H"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["depth_scale"], cam["depth_trunc"],
                         cam["downsample"], cuda)
    pts, cols, z, filt, trk = fe.downsample_and_make_pointcloud2(d16, rgb)
    rp, rc, rz, rf = _reference_cloud(d16, rgb, cam)
    assert pts.shape[0] == rp.shape[0] and trk.shape[0] == rf.shape[0] and trk.shape[0] < pts.shape[0]
    assert np.array_equal(trk.cpu().numpy(), rf)
    assert np.allclose(pts.cpu().numpy(), rp, rtol=2e-7, atol=0) and np.allclose(z.cpu().numpy(), rz, rtol=2e-7, atol=0)
    assert np.allclose(cols.cpu().numpy(), rc, rtol=2e-7, atol=0)
    expect = np.zeros(rp.shape[0], dtype=np.int32)
    expect[rf] = np.arange(1, len(rf) + 1)
    assert np.array_equal(filt.cpu().numpy(), expect)
    # zero-copy into the tracker == the numpy path
    tgt = S.sample_surface(20000, 5, 0.001)[0]
    res = []
    for dev_path in (True, False):
        reg = pygicp.FastGICP()
        reg.set_max_correspondence_distance(0.05)
        reg.set_max_knn_distance(99999)
        reg.set_input_target(tgt)
        reg.calculate_target_covariance_with_filter()
        if dev_path:
            reg.set_input_source(pts)
            reg.set_source_filter(trk.shape[0], filt)
        else:
            reg.set_input_source(pts.cpu().numpy())
            reg.set_source_filter(len(rf), expect)
        res.append((reg.align(pose.astype(np.float32)), reg.get_source_correspondence()))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1][0], res[1][1][0])


def test_keyframe_branch_matches_reference_expressions(cuda):
    from scipy.spatial.transform import Rotation

    from gs_icp_slam_b200.frontend import TrackerFrontEnd

    cam = dict(S.TUM)
    fe = TrackerFrontEnd(cam["W"], cam["H"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["depth_scale"], cam["depth_trunc"], 5, cuda)
    rng = np.random.default_rng(3)
    n = 12000
    pts = rng.normal(size=(n, 3)).astype(np.float32)
    rots = rng.normal(size=(n, 4)).astype(np.float32)
    rots /= np.linalg.norm(rots, axis=1, keepdims=True)
    pose = S.trajectory_pose(7, 200)
    # reference: mp_Tracker.py:224-229, 258-262, 385-392
    inv = np.linalg.inv(pose)
    T, R = inv[:3, 3], inv[:3, :3].transpose()
    ref_p = np.matmul(R, pts.transpose()).transpose() - np.matmul(R, T)
    x0, y0, z0, w0 = Rotation.from_matrix(R).as_quat()
    Q = rots
    ref_q = np.array([w0 * Q[:, 0] + x0 * Q[:, 3] + y0 * Q[:, 2] - z0 * Q[:, 1], w0 * Q[:, 1] + y0 * Q[:, 3] + z0 * Q[:, 0] - x0 * Q[:, 2],
                      w0 * Q[:, 2] + z0 * Q[:, 3] + x0 * Q[:, 1] - y0 * Q[:, 0], w0 * Q[:, 3] - x0 * Q[:, 0] - y0 * Q[:, 1] - z0 * Q[:, 2]]).T
    wp, wq = fe.to_world(torch.from_numpy(pts).to(cuda), torch.from_numpy(rots).to(cuda), pose)
    assert np.abs(wp.cpu().numpy() - ref_p).max() <= 2e-6 and np.abs(wq.cpu().numpy() - ref_q).max() <= 2e-6
    d2 = rng.uniform(0, 2e-3, size=9000).astype(np.float32)
    trk = np.sort(rng.choice(n, size=9000, replace=False)).astype(np.int32)
    keep = fe.eliminate_overlapped2(torch.from_numpy(d2).to(cuda), 1e-3, torch.from_numpy(trk).to(cuda))
    assert np.array_equal(keep.cpu().numpy(), trk[np.where(d2 > 1e-3)])
