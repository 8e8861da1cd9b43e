"""Tracker front-end on the device: what mp_Tracker.py does per frame around pygicp, as CUDA
tensors that feed FastGICP's zero-copy entry points.  CUDA only (no CPU path).

    fe = TrackerFrontEnd(W, H, fx, fy, cx, cy, depth_scale, depth_trunc, downsample_rate)
    pts, cols, z, filt, trk = fe.downsample_and_make_pointcloud2(depth_u16, rgb_u8)     # mp_Tracker.py:415-431 (+ :159-161)
    reg.set_input_source(pts); reg.set_source_filter(len(trk), filt); pose = reg.align(prev)
    world, rots_w = fe.to_world(pts, rots, pose)                                          # mp_Tracker.py:224-229, 258-262
    new_trk = fe.eliminate_overlapped2(sq_dist, th, trk)                                  # mp_Tracker.py:267-269
"""
import ctypes as C

import numpy as np
import torch

from ._lib import check, lib

from ._lib import _sig  # noqa: E402

_vp, _i = C.c_void_p, C.c_int
_sig("gsicp_frontend_max_points", _i, [_i, _i, _i])
_sig("gsicp_frontend_cloud", _i, [_vp, _vp, _i, _i, _i] + [C.c_float] * 6 + [_vp] * 5 + [C.POINTER(_i), C.POINTER(_i), _vp])
_sig("gsicp_frontend_keyframe", _i, [_i, _vp, _vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), _vp, _vp, _vp])
_sig("gsicp_frontend_not_overlapped", _i, [_i, _vp, C.c_float, _vp, _vp, C.POINTER(_i), _vp])


class TrackerFrontEnd:
    def __init__(self, W, H, fx, fy, cx, cy, depth_scale, depth_trunc, downsample_rate, device="cuda"):
        self.W, self.H, self.step = int(W), int(H), int(downsample_rate)
        self.fx, self.fy, self.cx, self.cy = float(fx), float(fy), float(cx), float(cy)
        self.depth_scale, self.depth_trunc = float(depth_scale), float(depth_trunc)
        self.device = torch.device(device)
        n = check(lib.gsicp_frontend_max_points(self.W, self.H, self.step))
        self.max_points = n
        e = lambda *s, dt=torch.float32: torch.empty(s, dtype=dt, device=self.device)
        self._pts, self._cols, self._z = e(n, 3), e(n, 3), e(n)
        self._filt, self._trk = e(n, dt=torch.int32), e(n, dt=torch.int32)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def downsample_and_make_pointcloud2(self, depth_img, rgb_img):
        """depth uint16 [H,W], rgb uint8 [H,W,3] (numpy or tensors; host inputs are copied H2D once) ->
        (points [n,3], colors [n,3], z_values [n], filter int32 [n] (1-based slots, 0 = untrackable), trackable int32 [m])."""
        d = torch.as_tensor(np.ascontiguousarray(depth_img) if isinstance(depth_img, np.ndarray) else depth_img)
        c = torch.as_tensor(np.ascontiguousarray(rgb_img) if isinstance(rgb_img, np.ndarray) else rgb_img)
        if d.dtype not in (torch.uint16, torch.int16):
            raise TypeError("depth image must be 16-bit (as read from the dataset's PNGs)")
        if c.dtype is not torch.uint8 or tuple(c.shape) != (self.H, self.W, 3) or tuple(d.shape) != (self.H, self.W):
            raise TypeError("rgb must be uint8 [H,W,3] and depth [H,W]")
        d = d.to(self.device, non_blocking=True).contiguous()
        c = c.to(self.device, non_blocking=True).contiguous()
        n, m = _i(0), _i(0)
        with torch.cuda.device(self.device):
            check(lib.gsicp_frontend_cloud(d.data_ptr(), c.data_ptr(), self.W, self.H, self.step, self.fx, self.fy, self.cx, self.cy,
                                           self.depth_scale, self.depth_trunc, self._pts.data_ptr(), self._cols.data_ptr(),
                                           self._z.data_ptr(), self._filt.data_ptr(), self._trk.data_ptr(), C.byref(n), C.byref(m),
                                           self._stream()), "gsicp_frontend_cloud")
        n, m = n.value, m.value
        return self._pts[:n], self._cols[:n], self._z[:n], self._filt[:n], self._trk[:m]

    def to_world(self, points_cam, rots, pose_c2w):
        """mp_Tracker.py:224-229 + :258-262: with inv = inverse(pose), T = inv[:3,3], R = inv[:3,:3]^T: points = R p - R T, and
        rots = quaternion_multiply(Rotation.from_matrix(R).as_quat(), rots).  rots may be None."""
        from scipy.spatial.transform import Rotation

        inv = np.linalg.inv(np.asarray(pose_c2w, dtype=np.float64))
        T = np.ascontiguousarray(inv[:3, 3], dtype=np.float32)
        R = np.ascontiguousarray(inv[:3, :3].transpose(), dtype=np.float32)
        q = np.ascontiguousarray(Rotation.from_matrix(inv[:3, :3].transpose()).as_quat(), dtype=np.float32)
        p = points_cam.detach().float().contiguous()
        n = p.shape[0]
        out_p = torch.empty_like(p)
        r = out_r = None
        if rots is not None:
            r = rots.detach().float().contiguous().view(-1, 4)
            out_r = torch.empty_like(r)
        fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
        with torch.cuda.device(self.device):
            check(lib.gsicp_frontend_keyframe(n, p.data_ptr(), r.data_ptr() if r is not None else None, fp(R), fp(T), fp(q),
                                              out_p.data_ptr(), out_r.data_ptr() if out_r is not None else None, self._stream()),
                  "gsicp_frontend_keyframe")
        return out_p, out_r

    def eliminate_overlapped2(self, sq_distances, threshold, trackable):
        """trackable_filter[np.where(distances > threshold)] (mp_Tracker.py:267-269, 374-380) on the device."""
        d = sq_distances.detach().float().contiguous()
        t = trackable.detach().to(torch.int32).contiguous()
        out = torch.empty_like(t)
        n = _i(0)
        with torch.cuda.device(self.device):
            check(lib.gsicp_frontend_not_overlapped(t.shape[0], d.data_ptr(), float(threshold), t.data_ptr(), out.data_ptr(), C.byref(n),
                                                    self._stream()), "gsicp_frontend_not_overlapped")
        return out[:n.value]
