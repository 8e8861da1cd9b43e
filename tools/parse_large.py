"""Summarise the JSON lines of bench.py --config c1|c4|c5 logs:  python tools/parse_large.py gpurun_out/large_*_TAG.log"""
import json
import sys

for path in sys.argv[1:]:
    d = None
    for line in open(path, errors="replace"):
        line = line.strip()
        if line.startswith("{") and '"metric"' in line:
            try:
                d = json.loads(line)
            except Exception:
                pass
    if d is None:
        print(path, "no JSON line")
        continue
    kern = {k: round(v["ms_per_launch"] * v["launches_per_step"], 3) for k, v in d.get("kernels", {}).items()}
    extra = {k: d[k] for k in ("lm_iterations", "pose_error_vs_ground_truth", "gpu_launches", "tile_instances_this_rank") if k in d}
    print(path.split("/")[-1], "n_gpus", d["n_gpus"], round(d["ms_per_step"], 4), "ms/step", round(d["value"], 1), d["unit"], kern, extra)
