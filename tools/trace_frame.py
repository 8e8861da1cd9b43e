"""Timeline of the mapper iteration (torch.profiler, CPU + CUDA): where the host leaves the GPU idle.  Development tool.
Writes gpurun_out/trace_mapper.json (chrome trace) and prints per-phase host time and GPU busy time per frame."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile, record_function  # noqa: E402

import bench  # noqa: E402

n = 12
cam, gmap, frames = bench.make_sequence(n + 1, bench.MAP_P)
eng = bench.Ours(cam, gmap, frames, torch.device("cuda:0"), 1, 0, loss="ssim_fused")
for i in range(4):
    eng.mapper_part(i, True)
torch.cuda.synchronize()


def one(i):
    f = frames[i + 1]
    c, m = f["d_cam"], eng.map
    with record_function("P1_settings+forward"):
        rs = eng.Settings(cam["H"], cam["W"], c["tanfovx"], c["tanfovy"], eng.bg, 1.0, c["viewmatrix"], c["projmatrix"], 0, c["campos"], False, False)
        depth, color, radii, is_used = eng.Rasterizer(rs)(means3D=m["means3D"], means2D=eng.means2D, opacities=m["opacities"],
                                                          shs=m["shs"], scales=m["scales"], rotations=m["rotations"])
    with record_function("P2_loss"):
        loss = eng.fused.mapping_loss(color, depth, f["d_rgb"], f["d_depth"])
    with record_function("P3_backward"):
        loss.backward()
    with record_function("P4_item"):
        lv = float(loss.item())
    with record_function("P5_cleargrads"):
        for k in m:
            m[k].grad = None
        eng.means2D.grad = None
    return lv


with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(4, n):
        with record_function("FRAME"):
            one(i)
        torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace("gpurun_out/trace_mapper.json")
ev = json.load(open("gpurun_out/trace_mapper.json"))["traceEvents"]
frames_ev = sorted([e for e in ev if e.get("name") == "FRAME" and e.get("ph") == "X"], key=lambda e: e["ts"])
kern = sorted([e for e in ev if e.get("cat") == "kernel"], key=lambda e: e["ts"])
for fe in frames_ev[2:6]:
    t0, t1 = fe["ts"], fe["ts"] + fe["dur"]
    print(f"FRAME host {fe['dur']:.0f} us")
    for ph in ("P1_settings+forward", "P2_loss", "P3_backward", "P4_item", "P5_cleargrads"):
        for e in ev:
            if e.get("name") == ph and e.get("ph") == "X" and t0 <= e["ts"] <= t1:
                print(f"   {ph:22s} start +{e['ts'] - t0:7.0f}  dur {e['dur']:7.0f}")
    ks = [k for k in kern if t0 <= k["ts"] <= t1 + 2000]
    busy = sum(k["dur"] for k in ks)
    print(f"   GPU kernels {len(ks)}, busy {busy:.0f} us, first +{ks[0]['ts'] - t0:.0f}, last end +{ks[-1]['ts'] + ks[-1]['dur'] - t0:.0f}")
    prev = None
    for k in ks:
        gap = 0 if prev is None else k["ts"] - (prev["ts"] + prev["dur"])
        print(f"      +{k['ts'] - t0:7.0f} {k['dur']:7.1f} us  gap {gap:6.1f}  {k['name'][:60]}")
        prev = k
os._exit(0)
