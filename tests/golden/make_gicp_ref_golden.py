"""Generates tests/golden/gicp_fastgicp_ref.npz: outputs of the REFERENCE's own tracker (fast_gicp's unmodified sources +
pybind11 module, oracle/_ref/fast_gicp built by `make -C oracle ref` where /root/reference exists) on two small seeded
cases of tests/gicp_cases.py.  The oracle restatement and the CUDA tracker are checked against these vectors where the
reference build itself is absent.  Run in the build container:
    python tests/golden/make_gicp_ref_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    from oracle import ref_gicp
    from tests import gicp_cases as cases

    out = {}
    a = cases.c1(ref_gicp.FastGICP, n=4000)
    for k in ("pose", "corr", "sqd", "H", "tgt_rots", "tgt_scales", "src_rots", "src_scales"):
        out["c1_" + k] = a[k]
    a = cases.tracker_c3(ref_gicp.FastGICP, P=20000, frames=(1, 2), keyframe_at=1)
    for k in ("pose1", "pose2", "corr1", "corr2", "sqd1", "sqd2", "H2", "rots_kf", "scales_kf"):
        out["trk_" + k] = a[k]
    path = os.path.join(ROOT, "tests", "golden", "gicp_fastgicp_ref.npz")
    np.savez_compressed(path, **out)
    print(path, {k: v.shape for k, v in out.items()}, os.path.getsize(path))


if __name__ == "__main__":
    main()
