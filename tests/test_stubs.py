"""oracle/stubs (stand-ins for viewer / metric packages the reference's scripts import): the PLY writer/reader the
reference's GaussianModel.save_ply / load_ply run on round-trips the attribute table, the depth
reader returns the 16-bit image, the swallow-everything modules swallow; and the synthetic dataset writer produces the
Replica layout the reference's loaders expect (mp_Tracker.py:341-352, utils/traj_utils.py:38-50)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUBS = os.path.join(ROOT, "oracle", "stubs")


def _with_stubs():
    if STUBS not in sys.path:
        sys.path.insert(1, STUBS)


def test_ply_roundtrip_of_the_gaussian_attribute_table(tmp_path):
    _with_stubs()
    from plyfile import PlyData, PlyElement

    # construct_list_of_attributes (gaussian_model.py:269-281) for sh_degree 0: x y z nx ny nz f_dc_0..2 opacity scale_0..2 rot_0..3
    names = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2", "opacity", "scale_0", "scale_1", "scale_2", "rot_0",
             "rot_1", "rot_2", "rot_3"]
    n = 1234
    rng = np.random.default_rng(0)
    attrs = rng.normal(size=(n, len(names))).astype(np.float32)
    el = np.empty(n, dtype=[(a, "f4") for a in names])
    el[:] = list(map(tuple, attrs))
    path = str(tmp_path / "scene.ply")
    PlyData([PlyElement.describe(el, "vertex")]).write(path)
    head = open(path, "rb").read(400).split(b"end_header")[0].decode()
    assert head.startswith("ply\nformat binary_little_endian 1.0\nelement vertex 1234\nproperty float x\n")
    back = PlyData.read(path)
    v = back.elements[0]
    assert [p.name for p in v.properties] == names and v.count == n
    for i, a in enumerate(names):
        assert np.array_equal(np.asarray(v[a]), attrs[:, i])
    assert np.array_equal(np.asarray(back["vertex"]["opacity"]), attrs[:, 9])


def test_dataset_writer_and_depth_reader(tmp_path):
    _with_stubs()
    import open3d as o3d
    import rerun as rr

    from gs_icp_slam_b200 import synthetic as S

    cfg = S.write_dataset(str(tmp_path / "data"), 2)
    lines = open(cfg).read().splitlines()
    assert lines[2].split() == ["640", "480", "517.3", "516.5", "318.6", "255.3", "5000.0", "3.0", "replica"]
    d = np.array(o3d.io.read_image(str(tmp_path / "data" / "depth_images" / "depth000001.png")))
    assert d.dtype == np.uint16 and d.shape == (480, 640) and d.max() > 1000
    depth, _ = S.raycast_depth(S.trajectory_pose(1, 200), S.TUM)
    assert np.abs(d.astype(np.float32) / 5000.0 - depth).max() <= 1.01e-4
    poses = np.loadtxt(str(tmp_path / "data" / "traj.txt")).reshape(-1, 4, 4)
    assert poses.shape[0] == 2 and np.allclose(poses[1], S.trajectory_pose(1, 200))
    assert sorted(os.listdir(tmp_path / "data" / "images")) == ["frame000000.jpg", "frame000001.jpg"]
    rr.init("x")
    rr.log("a/b", rr.Points3D([[0, 0, 0]], colors=[1, 2, 3]))  # swallowed
