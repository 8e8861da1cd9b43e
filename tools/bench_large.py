"""Large-config scaling benches (BASELINE configs C4 and C5), one process per GPU under torchrun:
  C4: 1280x960, 1M Gaussians, rasterizer fwd+bwd only, tiles sharded over the ranks, gradient all-reduce
  C5: GICP on a 2M x 2M point pair, source points sharded, all-reduce of the 28-double normal equations
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_large.py [c4|c5] [iters]
Prints one JSON line per config (rank 0): ms per iteration (max over ranks) — strong scaling, total work fixed."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from gs_icp_slam_b200 import rasterizer as R  # noqa: E402
from gs_icp_slam_b200 import sharding  # noqa: E402
from gs_icp_slam_b200 import synthetic as S  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
which = sys.argv[1] if len(sys.argv) > 1 else "c4"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t.item()))
    return float(np.median(ts))


if which == "c4":
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer

    W, H, P = 1280, 960, 1000000
    cam = dict(S.TUM)
    cam.update(W=W, H=H, fx=cam["fx"] * 2, fy=cam["fy"] * 2, cx=cam["cx"] * 2, cy=cam["cy"] * 2)
    g = S.gaussian_map(P, 4, scale=2.0)
    cm = S.camera_matrices(S.trajectory_pose(3, 20, scale=2.0), cam)
    t = {k: torch.from_numpy(v).to(dev).requires_grad_(True) for k, v in g.items()}
    c = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in cm.items()}
    m2 = torch.zeros_like(t["means3D"], requires_grad=True)
    gen = torch.Generator(device="cpu").manual_seed(5)
    gcol, gdep = torch.randn((3, H, W), generator=gen).to(dev), torch.randn((1, H, W), generator=gen).to(dev)
    R.set_tile_shard(world, rank)
    cb_time = [0.0, 0]
    if world > 1:
        inner = sharding.make_raster_allreduce(dev)

        def timed_cb(ptr, count, stream):
            t0 = time.perf_counter()
            inner(ptr, count, stream)
            cb_time[0] += time.perf_counter() - t0
            cb_time[1] = count

        R.set_allreduce(timed_cb)
    mask = sharding.tile_owner_mask(H, W, world, rank, dev)
    rs = GaussianRasterizationSettings(H, W, c["tanfovx"], c["tanfovy"], torch.zeros(3, device=dev), 1.0, c["viewmatrix"],
                                       c["projmatrix"], 0, c["campos"], False, False)
    info = {}

    def it():
        depth, color, radii, used = GaussianRasterizer(rs)(means3D=t["means3D"], means2D=m2, opacities=t["opacities"],
                                                           shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
        info["R"] = color.grad_fn.num_rendered
        ((color * gcol * mask).sum() + (depth * gdep * mask).sum()).backward()
        for k in t:
            t[k].grad = None
        m2.grad = None

    ms = timed(it, iters)
    from gs_icp_slam_b200 import _lib

    _lib.prof_reset()
    _lib.prof_enable(True)
    for _ in range(iters):
        it()
    torch.cuda.synchronize()
    _lib.prof_enable(False)
    kern = {k: round(v[0] / max(v[1], 1) * 1e3, 1) for k, v in _lib.prof_read().items() if v[1]}
    # phase split (forward / loss + backward), host-timed with a sync on both sides
    ph = {"fwd": 0.0, "bwd": 0.0}
    cb_time[0] = 0.0
    for _ in range(iters):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        depth, color, radii, used = GaussianRasterizer(rs)(means3D=t["means3D"], means2D=m2, opacities=t["opacities"],
                                                           shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ((color * gcol * mask).sum() + (depth * gdep * mask).sum()).backward()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        ph["fwd"] += (t1 - t0) * 1e3 / iters
        ph["bwd"] += (t2 - t1) * 1e3 / iters
        for k in t:
            t[k].grad = None
        m2.grad = None
    if rank == 0:
        print(json.dumps({"kernels_us": kern, "phase_ms": ph, "allreduce_cb_ms": cb_time[0] * 1e3 / iters,
                          "allreduce_floats": cb_time[1]}))
        print(json.dumps({"config": "C4 1280x960, 1M Gaussians, raster fwd+bwd, tile-sharded", "n_gpus": world, "ms_per_iter": ms,
                          "iters_per_s": 1e3 / ms, "tile_instances_this_rank": info["R"]}))
else:
    import pygicp

    n = int(os.environ.get("C5_POINTS", 2000000))
    tgt, src, T = S.gicp_pair(n, n, 6, 7, 0.001, scale=5.0)
    reg = pygicp.FastGICP()
    reg.set_max_correspondence_distance(0.25)
    reg.set_max_knn_distance(99999)
    if world > 1:
        reg.set_shard(world, rank, sharding.make_gicp_allreduce(dev))
    reg.set_input_target(tgt)
    reg.calculate_target_covariance()
    reg.set_input_source(src)
    reg.calculate_source_covariance()
    pose = np.eye(4)
    res = {}

    def it():
        res["H"], res["b"], res["e"] = reg.linearize(pose)

    ms = timed(it, iters)
    t0 = time.time()
    out = reg.align(np.eye(4))
    dt = (time.time() - t0) * 1e3
    if rank == 0:
        print(json.dumps({"config": f"C5 GICP {n}x{n} points, source-sharded", "n_gpus": world, "linearize_ms": ms,
                          "align_ms": dt, "lm_iterations": reg.last_iterations, "pose_err": float(np.abs(out - T).max())}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
