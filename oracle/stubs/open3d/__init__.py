"""oracle/stubs: open3d stand-in — the reference only calls o3d.io.read_image (gs_icp_slam.py:142,149; mp_Tracker.py:348,356)
on 16-bit depth PNGs and wraps the result in np.array()."""
import cv2 as _cv2

from _absorb import Absorb as _Absorb


class _IO:
    @staticmethod
    def read_image(path):
        img = _cv2.imread(str(path), _cv2.IMREAD_UNCHANGED)
        if img is None:
            raise FileNotFoundError(path)
        return img


io = _IO()
geometry = _Absorb()
utility = _Absorb()
visualization = _Absorb()
