#!/usr/bin/env python
"""X1: run the reference's UNMODIFIED SLAM drivers — gs_icp_slam_unlimit.py
spawning mp_Tracker_unlimit.py and mp_Mapper.py as two processes — on the synthetic RGB-D sequence, with either
  --impl ours       this repo's drop-in packages (pygicp, diff_gaussian_rasterization, simple_knn) first on PYTHONPATH
  --impl reference  the reference's own extensions (oracle/_ref/site, oracle/_ref/fast_gicp)
and report the tracker's own figures: "System FPS" = num_images / (t_end - total_start_time) (mp_Tracker.py:113,333) and
ATE RMSE.  The drivers come from oracle/_ref/gs_icp_slam (installed unmodified by oracle/install_ref_slam.sh; the GPU box
has no /root/reference); viewer / metric packages absent from the image are stood in by oracle/stubs.
    python tools/run_slam.py --impl ours --frames 50
Prints one JSON line."""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=50)
    ap.add_argument("--timeout", type=int, default=900)
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--dataset", default=None, help="reuse a dataset directory written by a previous run")
    a = ap.parse_args()
    ref = "/root/reference" if os.path.isfile("/root/reference/mp_Mapper.py") else os.path.join(ROOT, "oracle", "_ref", "gs_icp_slam")
    if not os.path.isfile(os.path.join(ref, "mp_Mapper.py")):
        print(json.dumps({"impl": a.impl, "unavailable": "reference SLAM scripts not installed (oracle/install_ref_slam.sh)"}))
        return 0
    from gs_icp_slam_b200 import synthetic as S

    work = a.dataset or tempfile.mkdtemp(prefix="gsicp_slam_")
    data = os.path.join(work, "data")
    if not os.path.isfile(os.path.join(data, "traj.txt")):
        t0 = time.time()
        cfg = S.write_dataset(data, a.frames, S.TUM)
        print(f"[run_slam] wrote {a.frames} frames to {data} in {time.time() - t0:.1f}s", file=sys.stderr)
    cfg = os.path.join(data, "caminfo.txt")
    out = os.path.join(work, f"out_{a.impl}")
    stubs = os.path.join(ROOT, "oracle", "stubs")
    if a.impl == "ours":
        path = [ROOT, stubs, ref]
    else:
        site = os.path.join(ROOT, "oracle", "_ref", "site")
        fg = os.path.join(ROOT, "oracle", "_ref", "fast_gicp")
        if not os.path.isdir(site) or not os.path.isdir(fg):
            print(json.dumps({"impl": a.impl, "unavailable": "oracle/_ref/site or oracle/_ref/fast_gicp missing"}))
            return 0
        path = [site, fg, stubs, ref]
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(path), OMP_WAIT_POLICY="passive", PYTHONUNBUFFERED="1")
    # TUM settings of the reference (tum.sh:135-142): keyframe_th 0.81, overlapped 1e-3 / 1e-3, max_corr 0.03,
    # trackable opacity 0.09, downsample 5
    cmd = [sys.executable, os.path.join(ref, "gs_icp_slam_unlimit.py"), "--dataset_path", data, "--config", cfg, "--output_path", out,
           "--keyframe_th", "0.81", "--knn_maxd", "99999.0", "--overlapped_th", "1e-3", "--overlapped_th2", "1e-3",
           "--max_correspondence_distance", "0.03", "--trackable_opacity_th", "0.09", "--downsample_rate", "5"]
    t0 = time.time()
    try:
        p = subprocess.run(cmd, cwd=ref, env=env, capture_output=True, text=True, timeout=a.timeout)
        rc, so, se = p.returncode, p.stdout, p.stderr
    except subprocess.TimeoutExpired as ex:
        rc, so, se = -9, (ex.stdout or b"").decode("utf8", "replace") if isinstance(ex.stdout, bytes) else (ex.stdout or ""), \
            (ex.stderr or b"").decode("utf8", "replace") if isinstance(ex.stderr, bytes) else (ex.stderr or "")
    wall = time.time() - t0
    fps = re.search(r"System FPS:\s*([0-9.]+)", so)
    ate = re.search(r"ATE RMSE:\s*([0-9.eE+-]+)", so)
    psnr = re.search(r"PSNR:\s*([0-9.]+)", so)
    res = {"impl": a.impl, "drivers": "gs_icp_slam_unlimit.py -> mp_Tracker_unlimit.py + mp_Mapper.py (unmodified, two processes)",
           "frames": a.frames, "rc": rc, "wall_s": wall, "system_fps": float(fps.group(1)) if fps else None,
           "ate_rmse_cm": float(ate.group(1)) if ate else None, "psnr": float(psnr.group(1)) if psnr else None,
           "extensions": "this repo's drop-ins (libgsicp_b200.so)" if a.impl == "ours" else "reference (oracle/_ref/site, oracle/_ref/fast_gicp)"}
    if fps is None:
        res["stdout_tail"], res["stderr_tail"] = so[-1500:], se[-2500:]
    print(json.dumps(res))
    if not a.keep and not a.dataset:
        shutil.rmtree(work, ignore_errors=True)
    return 0 if fps is not None else 1


if __name__ == "__main__":
    sys.exit(main())
