"""ctypes binding of oracle/libgicp_oracle.so (gicp_oracle.cpp).  Test infrastructure / CPU baseline.

`FastGICP` has the same method names as pygicp.FastGICP so a test can drive the CUDA implementation and
the CPU restatement with one call sequence."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libgicp_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            raise ImportError(f"{_PATH} missing: run `make -C oracle libgicp_oracle.so`")
        L = C.CDLL(_PATH)
        L.go_create.restype = C.c_void_p
        vp, i, d = C.c_void_p, C.c_int, C.c_double
        sig = {
            "go_destroy": (None, [vp]), "go_num_threads": (i, []), "go_set_num_threads": (None, [i]),
            "go_set_max_correspondence_distance": (None, [vp, d]), "go_set_max_knn_distance": (None, [vp, d]),
            "go_set_correspondence_randomness": (None, [vp, i]), "go_set_max_iterations": (None, [vp, i]),
            "go_set_input_source": (None, [vp, vp, i]), "go_set_input_target": (None, [vp, vp, i]),
            "go_set_source_filter": (None, [vp, i, vp, i]), "go_set_target_filter": (None, [vp, i, vp, i]),
            "go_calculate_target_covariance_with_filter": (i, [vp]), "go_calculate_source_covariance": (i, [vp]),
            "go_calculate_target_covariance": (i, [vp]), "go_calculate_target_covariance_withz": (i, [vp]),
            "go_set_source_z_values": (None, [vp, vp, i]), "go_set_target_z_values": (None, [vp, vp, i]),
            "go_swap_source_and_target": (None, [vp]), "go_get_fitness_score": (d, [vp, d]),
            "go_set_source_covariances_fromqs": (None, [vp, vp, vp, i]),
            "go_set_target_covariances_fromqs": (None, [vp, vp, vp, i]),
            "go_align": (i, [vp, vp, vp]), "go_has_converged": (i, [vp]), "go_get_final_hessian": (None, [vp, vp]),
            "go_get_source_correspondence": (i, [vp, vp, vp]), "go_linearize": (i, [vp, vp, vp, vp, vp]),
            "go_compute_error": (i, [vp, vp, vp]), "go_last_counts": (None, [vp, vp, vp]),
            "go_svd3": (None, [vp] * 4), "go_quat_from_matrix": (None, [vp, vp]), "go_quat_to_matrix": (None, [vp, vp]),
            "go_inverse3": (i, [vp, vp]), "go_ldlt_solve6": (None, [vp, vp, vp]), "go_so3_exp": (None, [vp, vp]),
            "go_knn": (None, [vp, i, i, vp, vp]),
        }
        for side in ("source", "target"):
            for what in ("size", "rotationsq_size", "scales_size", "cov_size"):
                sig[f"go_{side}_{what}"] = (i, [vp])
            for what in ("rotationsq", "scales", "covariances"):
                sig[f"go_get_{side}_{what}"] = (None, [vp, vp])
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def num_threads():
    return lib().go_num_threads()


def set_num_threads(n):
    """OpenMP threads of the oracle (fast_gicp: setNumThreads, fgi:36-44).  torchrun exports OMP_NUM_THREADS=1, so callers
    that want every host core (bench.py's reference arm) set it explicitly."""
    lib().go_set_num_threads(int(n))


class FastGICP:
    def __init__(self):
        self._L = lib()
        self._h = self._L.go_create()

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.go_destroy(h)

    def set_max_correspondence_distance(self, v):
        self._L.go_set_max_correspondence_distance(self._h, float(v))

    def set_max_knn_distance(self, v):
        self._L.go_set_max_knn_distance(self._h, float(v))

    def set_correspondence_randomness(self, k):
        self._L.go_set_correspondence_randomness(self._h, int(k))

    def set_max_iterations(self, n):
        self._L.go_set_max_iterations(self._h, int(n))

    def set_input_source(self, pts):
        a = np.ascontiguousarray(pts, dtype=np.float64)
        self._L.go_set_input_source(self._h, a.ctypes.data, a.shape[0])

    def set_input_target(self, pts):
        a = np.ascontiguousarray(pts, dtype=np.float64)
        self._L.go_set_input_target(self._h, a.ctypes.data, a.shape[0])

    def set_source_filter(self, n, f):
        f = np.ascontiguousarray(np.asarray(f).reshape(-1), dtype=np.int32)
        self._L.go_set_source_filter(self._h, int(n), f.ctypes.data, f.shape[0])

    def set_target_filter(self, n, f):
        f = np.ascontiguousarray(np.asarray(f).reshape(-1), dtype=np.int32)
        self._L.go_set_target_filter(self._h, int(n), f.ctypes.data, f.shape[0])

    def calculate_target_covariance_with_filter(self):
        assert self._L.go_calculate_target_covariance_with_filter(self._h) == 0

    def calculate_source_covariance(self):
        assert self._L.go_calculate_source_covariance(self._h) == 0

    def calculate_target_covariance(self):
        assert self._L.go_calculate_target_covariance(self._h) == 0

    def calculate_target_covariance_withz(self):
        assert self._L.go_calculate_target_covariance_withz(self._h) == 0

    def set_source_z_values(self, z):
        z = np.ascontiguousarray(np.asarray(z).reshape(-1), dtype=np.float32)
        self._L.go_set_source_z_values(self._h, z.ctypes.data, len(z))

    def set_target_z_values(self, z):
        z = np.ascontiguousarray(np.asarray(z).reshape(-1), dtype=np.float32)
        self._L.go_set_target_z_values(self._h, z.ctypes.data, len(z))

    def swap_source_and_target(self):
        self._L.go_swap_source_and_target(self._h)

    def get_fitness_score(self, max_range):
        return float(self._L.go_get_fitness_score(self._h, float(max_range)))

    def set_source_covariances_fromqs(self, r, s):
        r = np.ascontiguousarray(np.asarray(r).reshape(-1), dtype=np.float32)
        s = np.ascontiguousarray(np.asarray(s).reshape(-1), dtype=np.float32)
        self._L.go_set_source_covariances_fromqs(self._h, r.ctypes.data, s.ctypes.data, len(s) // 3)

    def set_target_covariances_fromqs(self, r, s):
        r = np.ascontiguousarray(np.asarray(r).reshape(-1), dtype=np.float32)
        s = np.ascontiguousarray(np.asarray(s).reshape(-1), dtype=np.float32)
        self._L.go_set_target_covariances_fromqs(self._h, r.ctypes.data, s.ctypes.data, len(s) // 3)

    def align(self, guess=None):
        g = np.eye(4, dtype=np.float32) if guess is None else np.ascontiguousarray(guess, dtype=np.float32)
        out = np.empty((4, 4), dtype=np.float32)
        self.last_iterations = self._L.go_align(self._h, g.ctypes.data, out.ctypes.data)
        if self.last_iterations < 0:
            raise RuntimeError(f"oracle align failed ({self.last_iterations})")
        return out

    def has_converged(self):
        return bool(self._L.go_has_converged(self._h))

    def get_final_hessian(self):  # LsqRegistration binding, main.cpp:170
        H = np.empty((6, 6), dtype=np.float64)
        self._L.go_get_final_hessian(self._h, H.ctypes.data)
        return H

    def _vec(self, size, get, dtype=np.float32, shape=None):
        n = getattr(self._L, size)(self._h)
        out = np.empty(n if shape is None else (n,) + shape, dtype=dtype)
        if n:
            getattr(self._L, get)(self._h, out.ctypes.data)
        return out

    def get_source_rotationsq(self):
        return self._vec("go_source_rotationsq_size", "go_get_source_rotationsq")

    def get_target_rotationsq(self):
        return self._vec("go_target_rotationsq_size", "go_get_target_rotationsq")

    def get_source_scales(self):
        return self._vec("go_source_scales_size", "go_get_source_scales")

    def get_target_scales(self):
        return self._vec("go_target_scales_size", "go_get_target_scales")

    def get_source_covariances(self):
        return self._vec("go_source_cov_size", "go_get_source_covariances", np.float64, (3, 3))

    def get_target_covariances(self):
        return self._vec("go_target_cov_size", "go_get_target_covariances", np.float64, (3, 3))

    def source_size(self):
        return self._L.go_source_size(self._h)

    def target_size(self):
        return self._L.go_target_size(self._h)

    def get_source_correspondence(self):
        n = self.source_size()
        c, s = np.empty(n, np.int32), np.empty(n, np.float32)
        if self._L.go_get_source_correspondence(self._h, c.ctypes.data, s.ctypes.data) != 0:
            raise RuntimeError("no correspondences for the current source")
        return c, s

    def linearize(self, pose):
        p = np.ascontiguousarray(pose, dtype=np.float64)
        H, b, e = np.empty((6, 6)), np.empty(6), C.c_double(0)
        assert self._L.go_linearize(self._h, p.ctypes.data, H.ctypes.data, b.ctypes.data, C.byref(e)) == 0
        return H, b, e.value

    def compute_error(self, pose):
        p = np.ascontiguousarray(pose, dtype=np.float64)
        e = C.c_double(0)
        assert self._L.go_compute_error(self._h, p.ctypes.data, C.byref(e)) == 0
        return e.value

    def last_counts(self):
        a, b = C.c_int(0), C.c_int(0)
        self._L.go_last_counts(self._h, C.byref(a), C.byref(b))
        return a.value, b.value


def knn(xyz, k):
    a = np.ascontiguousarray(xyz, dtype=np.float32)
    idx = np.empty((a.shape[0], k), np.int32)
    d2 = np.empty((a.shape[0], k), np.float32)
    lib().go_knn(a.ctypes.data, a.shape[0], k, idx.ctypes.data, d2.ctypes.data)
    return idx, d2
