"""X1: the reference's UNMODIFIED SLAM drivers (gs_icp_slam_unlimit.py -> mp_Tracker_unlimit.py +
mp_Mapper.py, two processes) run end to end on the synthetic sequence on top of this repo's drop-in packages."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_unmodified_drivers_complete_the_sequence(cuda):
    ref = os.path.join(ROOT, "oracle", "_ref", "gs_icp_slam", "mp_Mapper.py")
    if not os.path.isfile(ref) and not os.path.isfile("/root/reference/mp_Mapper.py"):
        pytest.skip("reference SLAM scripts not installed (oracle/install_ref_slam.sh needs /root/reference)")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_slam.py"), "--impl", "ours", "--frames", "50"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    line = [x for x in p.stdout.splitlines() if x.startswith("{")]
    assert line, p.stdout[-2000:] + p.stderr[-2000:]
    d = json.loads(line[-1])
    assert d["rc"] == 0 and d["system_fps"] is not None, d
    assert d["system_fps"] > 5.0
    assert d["ate_rmse_cm"] is not None and d["ate_rmse_cm"] < 5.0, d  # the synthetic trajectory is tracked to centimetres
