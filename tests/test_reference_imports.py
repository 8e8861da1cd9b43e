"""Drop-in check at import level (CPU, build container only): the reference's own modules — imported unmodified from
/root/reference with this repo first on sys.path — bind OUR pygicp / diff_gaussian_rasterization / simple_knn.  Third-party
packages the reference needs but this image lacks (open3d, rerun, plyfile, cv2, ...) are replaced by empty stand-ins; they are
not on the hot path.  Skipped where the reference tree is absent (the GPU box)."""
import importlib
import os
import sys
import types
from unittest import mock

import pytest

try:  # native extension modules must not be dropped from sys.modules and re-imported (OpenCV breaks on a second import)
    import cv2  # noqa: F401
    import scipy.spatial.transform  # noqa: F401
except Exception:
    pass

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


class _Anything(types.ModuleType):
    """A module whose every attribute is a MagicMock (stand-in for an absent third-party dependency)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = mock.MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


@pytest.fixture()
def reference_on_path():
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    sys.path[:0] = [ROOT, REF]
    yield
    sys.path[:] = saved_path
    for k in list(sys.modules):
        if k not in saved_mods:
            del sys.modules[k]


def _import_with_stand_ins(name, limit=40):
    for _ in range(limit):
        try:
            return importlib.import_module(name)
        except ModuleNotFoundError as e:
            missing = e.name
            if missing is None or missing.split(".")[0] in ("pygicp", "diff_gaussian_rasterization", "simple_knn", "gs_icp_slam_b200"):
                raise
            parts = missing.split(".")
            for i in range(1, len(parts) + 1):
                sys.modules.setdefault(".".join(parts[:i]), _Anything(".".join(parts[:i])))
    raise RuntimeError(f"too many missing modules while importing {name}")


def test_reference_modules_bind_our_extensions(reference_on_path):
    import diff_gaussian_rasterization as ours_dgr
    import pygicp as ours_gicp
    from simple_knn._C import distCUDA2 as ours_dist

    assert ours_gicp.__file__.startswith(ROOT) and ours_dgr.__file__.startswith(ROOT)
    gr = _import_with_stand_ins("gaussian_renderer")      # gaussian_renderer/__init__.py:14
    assert gr.GaussianRasterizer is ours_dgr.GaussianRasterizer
    assert gr.GaussianRasterizationSettings is ours_dgr.GaussianRasterizationSettings
    gm = _import_with_stand_ins("scene.gaussian_model")   # scene/gaussian_model.py:20
    assert gm.distCUDA2 is ours_dist
    tr = _import_with_stand_ins("mp_Tracker")             # mp_Tracker.py:15,53
    assert tr.pygicp is ours_gicp and tr.pygicp.FastGICP is ours_gicp.FastGICP
    mp = _import_with_stand_ins("mp_Mapper")
    assert mp.render_3.__module__ == "gaussian_renderer"
    # the keyword names the reference's renderer passes (gaussian_renderer/__init__.py:36-51,86-94) exist on our classes
    import inspect

    assert list(ours_dgr.GaussianRasterizationSettings._fields) == [
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix", "sh_degree",
        "campos", "prefiltered", "debug"]
    assert list(inspect.signature(ours_dgr.GaussianRasterizer.__init__).parameters)[1:] == ["raster_settings"]
    assert list(inspect.signature(ours_dgr.GaussianRasterizer.forward).parameters)[1:] == [
        "means3D", "means2D", "opacities", "shs", "colors_precomp", "scales", "rotations", "cov3D_precomp"]
    # every FastGICP method the trackers call (mp_Tracker.py:53-308)
    for m in ("set_max_correspondence_distance", "set_max_knn_distance", "set_input_target", "set_input_source",
              "set_target_filter", "set_source_filter", "calculate_target_covariance_with_filter", "get_target_rotationsq",
              "get_target_scales", "get_source_rotationsq", "get_source_scales", "align", "get_source_correspondence",
              "set_target_covariances_fromqs"):
        assert callable(getattr(ours_gicp.FastGICP, m)), m
