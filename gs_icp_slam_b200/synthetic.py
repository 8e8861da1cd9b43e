"""Deterministic synthetic inputs for tests/ and bench.py (numpy only).

Scene: the inside of a 4 x 3 x 2.5 m room with two 1 m cubes on the floor.  From it we derive
  * surface point clouds (GICP configs C1 / C5),
  * Gaussian maps (means on the surfaces, flat surfels, xyzw unit quaternions) for the rasterizer
    (configs C2 / C3 / C4),
  * ray-cast RGB-D frames along a circular trajectory and the tracker's down-sampled point cloud
    (same sampling pattern as the reference's mp_Tracker.set_downsample_filter, mp_Tracker.py:394-413),
  * cameras in the reference's matrix conventions (scene/shared_objs.py:163-166,
    utils/graphics_utils.py:38-71): `viewmatrix` = (world->view)^T, `projmatrix` = (P @ world->view)^T.
"""
import math

import numpy as np

ROOM = np.array([4.0, 3.0, 2.5])
CUBES = [np.array([[0.8, 0.5, 0.0], [1.8, 1.5, 1.0]]), np.array([[2.5, 1.6, 0.0], [3.5, 2.6, 1.0]])]


def _faces(scale=1.0):
    """List of (origin, edge_u, edge_v, normal) rectangles; normals point into free space."""
    f = []
    L = ROOM * scale
    ex, ey, ez = np.eye(3)
    f.append((np.zeros(3), L[0] * ex, L[1] * ey, ez))                   # floor
    f.append((np.array([0, 0, L[2]]), L[0] * ex, L[1] * ey, -ez))       # ceiling
    f.append((np.zeros(3), L[0] * ex, L[2] * ez, ey))                   # y = 0 wall
    f.append((np.array([0, L[1], 0]), L[0] * ex, L[2] * ez, -ey))       # y = Ly wall
    f.append((np.zeros(3), L[1] * ey, L[2] * ez, ex))                   # x = 0 wall
    f.append((np.array([L[0], 0, 0]), L[1] * ey, L[2] * ez, -ex))       # x = Lx wall
    for cb in CUBES:
        lo, hi = cb[0] * scale, cb[1] * scale
        d = hi - lo
        f.append((np.array([lo[0], lo[1], hi[2]]), d[0] * ex, d[1] * ey, ez))    # top
        f.append((lo, d[0] * ex, d[2] * ez, -ey))
        f.append((np.array([lo[0], hi[1], lo[2]]), d[0] * ex, d[2] * ez, ey))
        f.append((lo, d[1] * ey, d[2] * ez, -ex))
        f.append((np.array([hi[0], lo[1], lo[2]]), d[1] * ey, d[2] * ez, ex))
    return f


def texture(points):
    """Procedural checker/gradient colour in [0,1] for world points (n,3)."""
    p = np.asarray(points, dtype=np.float64)
    chk = ((np.floor(p[:, 0] * 4) + np.floor(p[:, 1] * 4) + np.floor(p[:, 2] * 4)) % 2)
    rgb = np.stack([0.25 + 0.5 * chk, 0.2 + 0.6 * (p[:, 1] / ROOM[1] % 1.0), 0.2 + 0.6 * (p[:, 2] / ROOM[2] % 1.0)], 1)
    return np.clip(rgb, 0.0, 1.0)


def sample_surface(n, seed, noise=0.0, scale=1.0):
    """n points uniformly on the scene surfaces -> (points float64 (n,3), normals (n,3))."""
    rng = np.random.default_rng(seed)
    faces = _faces(scale)
    areas = np.array([np.linalg.norm(np.cross(u, v)) for _, u, v, _ in faces])
    which = rng.choice(len(faces), size=n, p=areas / areas.sum())
    a, b = rng.random(n), rng.random(n)
    org = np.stack([faces[w][0] for w in which])
    eu = np.stack([faces[w][1] for w in which])
    ev = np.stack([faces[w][2] for w in which])
    nrm = np.stack([faces[w][3] for w in which]).astype(np.float64)
    pts = org + a[:, None] * eu + b[:, None] * ev
    if noise > 0:
        pts = pts + rng.normal(0.0, noise, size=pts.shape)
    return pts, nrm


def rotation_about(axis, angle_rad):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + math.sin(angle_rad) * K + (1 - math.cos(angle_rad)) * (K @ K)


def gt_pose(angle_deg=2.0, trans=(0.03, -0.02, 0.01)):
    """The C1 ground-truth SE(3): rotation about (1,2,3)/|.|, small translation."""
    T = np.eye(4)
    T[:3, :3] = rotation_about([1, 2, 3], math.radians(angle_deg))
    T[:3, 3] = trans
    return T


def gicp_pair(n_target, n_source, seed_t=0, seed_s=1, noise=0.001, scale=1.0):
    """Target cloud (world) and an independently sampled source cloud moved by inv(T_gt)."""
    tgt, _ = sample_surface(n_target, seed_t, noise, scale)
    src_w, _ = sample_surface(n_source, seed_s, noise, scale)
    T = gt_pose()
    Ti = np.linalg.inv(T)
    src = src_w @ Ti[:3, :3].T + Ti[:3, 3]
    return tgt, src, T


def _mat_to_quat_xyzw(R):
    """Batch rotation matrices (n,3,3) -> unit quaternions (n,4) x,y,z,w (vectorised Shepperd)."""
    n = R.shape[0]
    m00, m11, m22 = R[:, 0, 0], R[:, 1, 1], R[:, 2, 2]
    cand = np.stack([1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22, 1 + m00 + m11 + m22], 1)
    k = np.argmax(cand, axis=1)
    s = 2.0 * np.sqrt(np.maximum(cand[np.arange(n), k], 1e-12))
    q = np.zeros((n, 4))
    # k == 3: w largest
    w = k == 3
    q[w, 3] = 0.25 * s[w]
    q[w, 0] = (R[w, 2, 1] - R[w, 1, 2]) / s[w]
    q[w, 1] = (R[w, 0, 2] - R[w, 2, 0]) / s[w]
    q[w, 2] = (R[w, 1, 0] - R[w, 0, 1]) / s[w]
    for i in range(3):
        j, l = (i + 1) % 3, (i + 2) % 3
        m = k == i
        q[m, i] = 0.25 * s[m]
        q[m, 3] = (R[m, l, j] - R[m, j, l]) / s[m]
        q[m, j] = (R[m, j, i] + R[m, i, j]) / s[m]
        q[m, l] = (R[m, l, i] + R[m, i, l]) / s[m]
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def gaussian_map(P, seed, scale=1.0, mean_size=0.02, sh_degree=0):
    """Gaussian map on the scene surfaces.  float32 arrays:
    means3D (P,3), scales (P,3) [activated], rotations (P,4) xyzw unit, opacities (P,1) [activated],
    shs (P,M,3) with the DC term from the texture."""
    rng = np.random.default_rng(seed)
    pts, nrm = sample_surface(P, seed + 1000, 0.0, scale)
    s = np.exp(rng.normal(math.log(mean_size * scale), 0.5, size=(P, 3)))
    s[:, 2] *= 0.1  # flat along the surface normal
    # surface frame: z axis = normal, x axis = any tangent rotated by a random in-plane angle
    helper = np.where(np.abs(nrm[:, [2]]) < 0.9, np.array([[0, 0, 1.0]]), np.array([[1.0, 0, 0]]))
    t1 = np.cross(nrm, helper)
    t1 /= np.linalg.norm(t1, axis=1, keepdims=True)
    t2 = np.cross(nrm, t1)
    ang = rng.uniform(0, 2 * math.pi, P)
    u = np.cos(ang)[:, None] * t1 + np.sin(ang)[:, None] * t2
    v = np.cross(nrm, u)
    R = np.stack([u, v, nrm], axis=2)  # columns
    # vectorised matrix -> quaternion for the common positive-trace case, loop for the rest
    q = _mat_to_quat_xyzw(R)
    opac = rng.uniform(0.1, 0.99, size=(P, 1))
    M = (sh_degree + 1) ** 2
    shs = np.zeros((P, M, 3))
    shs[:, 0, :] = (texture(pts / scale) - 0.5) / 0.28209479177387814
    if M > 1:
        shs[:, 1:, :] = rng.normal(0, 0.05, size=(P, M - 1, 3))
    f32 = np.float32
    return dict(means3D=pts.astype(f32), scales=s.astype(f32), rotations=q.astype(f32), opacities=opac.astype(f32),
                shs=shs.astype(f32))


# ----------------------------------------------------------------------------------------------
# cameras
# ----------------------------------------------------------------------------------------------
TUM = dict(W=640, H=480, fx=517.3, fy=516.5, cx=318.6, cy=255.3, depth_scale=5000.0, depth_trunc=3.0, downsample=5)
REPLICA = dict(W=640, H=480, fx=320.0, fy=320.0, cx=319.5, cy=239.5, depth_scale=6553.5, depth_trunc=12.0, downsample=10)


def trajectory_pose(i, n_frames, scale=1.0, radius=0.5, sweep_deg=90.0):
    """Camera-to-world pose of frame i: circle of `radius` around the room centre, yaw sweeping sweep_deg."""
    c = ROOM * scale * np.array([0.5, 0.5, 0.48])
    a = math.radians(sweep_deg) * (i / max(n_frames - 1, 1))
    pos = c + radius * scale * np.array([math.cos(a), math.sin(a), 0.0])
    yaw = a + math.pi * 0.75  # look across the room
    fwd = np.array([math.cos(yaw), math.sin(yaw), -0.12])
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.array([0, 0, 1.0]))
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, down, fwd, pos
    return c2w


def camera_matrices(c2w, cam):
    """Reference-convention camera tensors as float32 numpy: viewmatrix (4,4), projmatrix (4,4), campos (3,),
    tanfovx, tanfovy."""
    W, H = cam["W"], cam["H"]
    fovx = 2 * math.atan(W / (2 * cam["fx"]))
    fovy = 2 * math.atan(H / (2 * cam["fy"]))
    w2c = np.linalg.inv(c2w)
    znear, zfar = 0.01, 100.0
    tanx, tany = math.tan(fovx / 2), math.tan(fovy / 2)
    Pm = np.zeros((4, 4))
    Pm[0, 0] = 1.0 / tanx
    Pm[1, 1] = 1.0 / tany
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    view = np.float32(w2c).T.copy()
    proj = (view.astype(np.float64) @ Pm.T).astype(np.float32)
    campos = np.float32(c2w[:3, 3])
    return dict(viewmatrix=view, projmatrix=proj, campos=campos, tanfovx=float(tanx), tanfovy=float(tany))


def raycast_depth(c2w, cam, scale=1.0):
    """Analytic depth (z in the camera frame, float32 HxW) and world hit points (H*W,3) of the scene."""
    W, H = cam["W"], cam["H"]
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    d_cam = np.stack([(u - cam["cx"]) / cam["fx"], (v - cam["cy"]) / cam["fy"], np.ones_like(u)], -1).reshape(-1, 3)
    d = d_cam @ c2w[:3, :3].T
    o = c2w[:3, 3]
    L = ROOM * scale
    with np.errstate(divide="ignore", invalid="ignore"):
        # room: we are inside -> exit distance
        t1 = (0.0 - o) / d
        t2 = (L - o) / d
        t_room = np.min(np.maximum(t1, t2), axis=1)
        t_hit = t_room
        for cb in CUBES:
            lo, hi = cb[0] * scale, cb[1] * scale
            a, b = (lo - o) / d, (hi - o) / d
            tn = np.max(np.minimum(a, b), axis=1)
            tf = np.min(np.maximum(a, b), axis=1)
            ok = (tn <= tf) & (tn > 0)
            t_hit = np.where(ok & (tn < t_hit), tn, t_hit)
    hit = o + t_hit[:, None] * d
    return t_hit.reshape(H, W).astype(np.float32), hit  # depth along z_cam equals t because d_cam.z == 1


def downsample_indices(cam):
    """Pixel picks of mp_Tracker.set_downsample_filter (mp_Tracker.py:394-413)."""
    W, H, s = cam["W"], cam["H"], cam["downsample"]
    h_val = s * np.arange(0, int(H / s) + 1) - 1
    h_val[0] = 0
    cols = np.arange(0, W, s)
    idx = (h_val[:, None] * W + cols[None, :]).flatten()
    vv, uu = idx // W, idx % W
    return idx, (uu - cam["cx"]) / cam["fx"], (vv - cam["cy"]) / cam["fy"]


def tracker_cloud(depth, cam):
    """Camera-frame point cloud the tracker hands to pygicp (mp_Tracker.py:415-431): float64 (N,3) in
    raster order with z != 0, plus the indices of trackable points (z <= depth_trunc)."""
    idx, x_pre, y_pre = downsample_indices(cam)
    # quantise like a uint16 depth PNG
    z = np.round(depth.flatten()[idx] * cam["depth_scale"]).astype(np.uint16).astype(np.float32) / np.float32(cam["depth_scale"])
    nz = z != 0
    z = z[nz]
    pts = np.stack([x_pre[nz].astype(np.float32) * z, y_pre[nz].astype(np.float32) * z, z], -1)
    trackable = np.where(z <= cam["depth_trunc"])[0]
    return pts.astype(np.float64), trackable


def trackable_filter(n_points, trackable):
    """int32 filter of mp_Tracker.py:159-161: 0 = untrackable, else 1-based slot."""
    f = np.zeros(n_points, dtype=np.int32)
    f[trackable] = np.arange(1, len(trackable) + 1, dtype=np.int32)
    return f


def write_dataset(path, n_frames, cam=None, n_traj=200):
    """Write the synthetic RGB-D sequence to disk in the reference's Replica layout (mp_Tracker.py:341-352,
    utils/traj_utils.py:38-50): images/frameNNNNNN.jpg, depth_images/depthNNNNNN.png (uint16, depth * depth_scale),
    traj.txt (one flattened camera-to-world 4x4 per line) and caminfo.txt (configs/*/caminfo.txt format).
    Returns the path of the camera file."""
    import os

    import cv2

    cam = dict(TUM if cam is None else cam)
    os.makedirs(os.path.join(path, "images"), exist_ok=True)
    os.makedirs(os.path.join(path, "depth_images"), exist_ok=True)
    poses = []
    for i in range(n_frames):
        c2w = trajectory_pose(i, n_traj)
        depth, hit = raycast_depth(c2w, cam)
        rgb = texture(hit).reshape(cam["H"], cam["W"], 3)
        img = np.clip(np.round(rgb * 255.0), 0, 255).astype(np.uint8)
        cv2.imwrite(os.path.join(path, "images", f"frame{i:06d}.jpg"), img, [cv2.IMWRITE_JPEG_QUALITY, 95])
        d16 = np.clip(np.round(depth * cam["depth_scale"]), 0, 65535).astype(np.uint16)
        cv2.imwrite(os.path.join(path, "depth_images", f"depth{i:06d}.png"), d16)
        poses.append(c2w)
    with open(os.path.join(path, "traj.txt"), "w") as f:
        for p in poses:
            f.write(" ".join(f"{v:.12e}" for v in p.reshape(-1)) + "\n")
    cfg = os.path.join(path, "caminfo.txt")
    with open(cfg, "w") as f:
        f.write("## camera parameters\nW H fx fy cx cy depth_scale depth_trunc dataset_type\n")
        f.write(f"{cam['W']} {cam['H']} {cam['fx']} {cam['fy']} {cam['cx']} {cam['cy']} {cam['depth_scale']} {cam['depth_trunc']} replica")
    return cfg
