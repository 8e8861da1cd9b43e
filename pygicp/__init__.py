"""Drop-in `pygicp` module backed by libgsicp_b200.so (sm_100a CUDA), for the reference's
mp_Tracker.py / mp_Tracker_unlimit.py (`self.reg = pygicp.FastGICP()`, mp_Tracker.py:53).

Exports the class the trackers use.  The other registrations the reference binds (FastVGICP,
FastVGICPCuda, NDTCuda, align_points, downsample: submodules/fast_gicp/src/python/main.cpp:47-147,264-299)
are outside the SLAM hot path and raise NotImplementedError.
"""
from gs_icp_slam_b200.gicp import FastGICP

__version__ = "b200"


def _out_of_scope(name):
    def f(*a, **k):
        raise NotImplementedError(f"pygicp.{name} is outside the GS-ICP-SLAM hot path and is not provided by gs_icp_slam_b200")

    f.__name__ = name
    return f


align_points = _out_of_scope("align_points")
downsample = _out_of_scope("downsample")
FastVGICP = _out_of_scope("FastVGICP")

__all__ = ["FastGICP", "align_points", "downsample", "FastVGICP"]
