"""Phase breakdown of align_lm_kernel at the C3 tracker shape (12 416-point frames against a 300k-Gaussian target):
GSICP_LM_MARKS=1 makes block 0 record %globaltimer at every phase boundary; the library prints the deltas (us) to stderr as
    start | per outer iteration: L, barrier, reduce | per trial: solve, E, barrier, reduce | ... | publish
Also prints the wall time of align() around it (host spin included).   python tools/prof_align.py [frames]"""
import os
import sys
import time

os.environ.setdefault("GSICP_LM_MARKS", "1")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import pygicp  # noqa: E402
from gs_icp_slam_b200 import synthetic as S  # noqa: E402

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 6
cam = S.TUM
g = S.gaussian_map(300000, 3)
r = pygicp.FastGICP()
r.set_max_correspondence_distance(0.03)
r.set_max_knn_distance(99999)
r.set_input_target(g["means3D"].astype(np.float64))
r.set_target_covariances_fromqs(g["rotations"].reshape(-1), g["scales"].reshape(-1))
pose = S.trajectory_pose(0, 200).astype(np.float32)
for f in range(1, n_frames + 1):
    pts, tr = S.tracker_cloud(S.raycast_depth(S.trajectory_pose(f, 200), cam)[0], cam)
    r.set_input_source(pts)
    r.set_source_filter(len(tr), S.trackable_filter(len(pts), tr))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pose = np.array(r.align(pose))
    t1 = time.perf_counter()
    print(f"frame {f}: align wall {1e6 * (t1 - t0):.1f} us, iterations {r.last_iterations}", flush=True)
