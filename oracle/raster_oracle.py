"""ctypes binding of oracle/libraster_oracle.so (raster_oracle.c) — numpy in, numpy out.  Test infrastructure."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libraster_oracle.so")


class _Args(C.Structure):
    _fields_ = [("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("W", C.c_int), ("H", C.c_int),
                ("tan_fovx", C.c_float), ("tan_fovy", C.c_float), ("scale_modifier", C.c_float)] + \
               [(n, C.c_void_p) for n in ("bg", "means3D", "shs", "colors", "opac", "scales", "rots", "cov_pre", "view",
                                          "proj", "campos")]


def _load():
    if not os.path.exists(_PATH):
        raise ImportError(f"{_PATH} missing: run `make -C oracle libraster_oracle.so`")
    lib = C.CDLL(_PATH)
    lib.ro_forward.restype = C.c_void_p
    lib.ro_forward.argtypes = [C.POINTER(_Args), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ro_backward.restype = C.c_int
    lib.ro_backward.argtypes = [C.POINTER(_Args), C.c_void_p, C.c_void_p] + [C.c_void_p] * 10
    lib.ro_num_rendered.restype = C.c_int
    lib.ro_num_rendered.argtypes = [C.c_void_p]
    for n in ("ro_point_list", "ro_ranges", "ro_n_contrib", "ro_final_T"):
        getattr(lib, n).restype = C.c_void_p
        getattr(lib, n).argtypes = [C.c_void_p]
    lib.ro_free.argtypes = [C.c_void_p]
    return lib


_lib = None


def _f32(x):
    return None if x is None else np.ascontiguousarray(x, dtype=np.float32)


def _p(x):
    return None if x is None else x.ctypes.data


class Result:
    pass


def forward_backward(scene, cam, H, W, bg, sh_degree=0, scale_modifier=1.0, colors_precomp=None, cov3D_precomp=None,
                     dL_dcolor=None, dL_ddepth=None):
    """Run the CPU restatement.  scene: dict(means3D, shs, opacities, scales, rotations); cam: dict(viewmatrix,
    projmatrix, campos, tanfovx, tanfovy).  Returns an object with color, depth, radii, is_used, num_rendered,
    point_list, ranges, n_contrib, final_T and (when gradients are given) the 8 gradient arrays."""
    global _lib
    if _lib is None:
        _lib = _load()
    means = _f32(scene["means3D"])
    P = means.shape[0]
    shs = None if colors_precomp is not None else _f32(scene["shs"])
    M = 0 if shs is None else shs.shape[1]
    cols = _f32(colors_precomp)
    opac = _f32(scene["opacities"]).reshape(-1)
    scales = None if cov3D_precomp is not None else _f32(scene["scales"])
    rots = None if cov3D_precomp is not None else _f32(scene["rotations"])
    covp = _f32(cov3D_precomp)
    view, proj, campos, bgc = _f32(cam["viewmatrix"]), _f32(cam["projmatrix"]), _f32(cam["campos"]), _f32(bg)
    a = _Args(P, sh_degree, M, W, H, cam["tanfovx"], cam["tanfovy"], scale_modifier, _p(bgc), _p(means), _p(shs),
              _p(cols), _p(opac), _p(scales), _p(rots), _p(covp), _p(view), _p(proj), _p(campos))
    r = Result()
    r.color = np.zeros((3, H, W), np.float32)
    r.depth = np.zeros((1, H, W), np.float32)
    r.radii = np.zeros(P, np.int32)
    r.is_used = np.zeros(P, np.uint8)
    st = _lib.ro_forward(C.byref(a), r.color.ctypes.data, r.depth.ctypes.data, r.radii.ctypes.data, r.is_used.ctypes.data)
    try:
        R = _lib.ro_num_rendered(st)
        r.num_rendered = R
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        if P > 0:
            r.point_list = np.ctypeslib.as_array(C.cast(_lib.ro_point_list(st), C.POINTER(C.c_uint32)), (max(R, 1),))[:R].astype(np.int64)
            r.ranges = np.ctypeslib.as_array(C.cast(_lib.ro_ranges(st), C.POINTER(C.c_uint32)), (tiles * 2,)).reshape(tiles, 2).astype(np.int64)
            r.n_contrib = np.ctypeslib.as_array(C.cast(_lib.ro_n_contrib(st), C.POINTER(C.c_uint32)), (H * W,)).reshape(H, W).astype(np.int64)
            r.final_T = np.ctypeslib.as_array(C.cast(_lib.ro_final_T(st), C.POINTER(C.c_float)), (H * W,)).reshape(H, W).copy()
        else:
            r.point_list = np.zeros(0, np.int64)
            r.ranges = np.zeros((tiles, 2), np.int64)
        if dL_dcolor is not None:
            gc, gd = _f32(dL_dcolor), _f32(dL_ddepth)
            z = lambda *s: np.zeros(s, np.float32)
            r.dL_dmeans2D, r.dL_dcolors, r.dL_dopacity = z(P, 3), z(P, 3), z(P, 1)
            r.dL_dmeans3D, r.dL_dcov3D, r.dL_dsh = z(P, 3), z(P, 6), z(P, max(M, 1), 3)
            r.dL_dscales, r.dL_drotations = z(P, 3), z(P, 4)
            _lib.ro_backward(C.byref(a), st, r.radii.ctypes.data, gc.ctypes.data, gd.ctypes.data,
                             r.dL_dmeans2D.ctypes.data, r.dL_dcolors.ctypes.data, r.dL_dopacity.ctypes.data,
                             r.dL_dmeans3D.ctypes.data, r.dL_dcov3D.ctypes.data, r.dL_dsh.ctypes.data,
                             r.dL_dscales.ctypes.data, r.dL_drotations.ctypes.data)
            if M == 0:
                r.dL_dsh = z(P, 0, 3)
    finally:
        _lib.ro_free(st)
    r.is_used = r.is_used.astype(bool)
    return r
