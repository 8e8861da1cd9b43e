"""Summarise an ncu report (--set full) as text for profiles/:  python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt
With `--traffic profiles/ncu_traffic.json [tag]` the per-launch DRAM bytes and warp-instruction counts of this repo's hot
kernels are also written to the JSON file bench.py reads for `roofline.traffic` (mean over the captured launches)."""
import json
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_static", "smem static"), ("launch__shared_mem_per_block_dynamic", "smem dynamic"),
    ("launch__occupancy_limit_registers", "occ limit regs (blocks)"), ("launch__occupancy_limit_shared_mem", "occ limit smem (blocks)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp inst"),
    ("smsp__thread_inst_executed_pred_on_per_inst_executed.ratio", "pred-on threads / warp inst"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "warps stalled long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "warps stalled short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "warps stalled barrier / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "warps stalled wait / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "warps stalled math pipe throttle / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "warps stalled mio throttle / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "warps stalled lg throttle / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "warps stalled not selected / issue"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "warps stalled no instruction / issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "warps stalled branch resolving / issue"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "warps stalled dispatch stall / issue"),
    ("smsp__average_warps_issue_stalled_selected_per_issue_active.ratio", "warps stalled selected / issue"),
]


TRAFFIC_KERNELS = {"render_backward_kernel": "render_backward", "render_forward_kernel": "render_forward",
                   "gaussian_backward_kernel": "gaussian_backward", "preprocess_kernel": "preprocess", "tile_sort_kernel": "tile_sort",
                   "align_lm_kernel": "gicp_align", "linearize_kernel": "gicp_linearize", "knn_kernel": "gicp_covariance"}


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main(path, traffic_path=None, tag=""):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    print(f"# {path}: ncu --set full --clock-control none --import-source on (one row per profiled launch)")
    acc = {}
    for r in rows[2:]:
        d = dict(zip(head, r))
        u = dict(zip(head, units))
        for frag, name in TRAFFIC_KERNELS.items():
            if frag in d["Kernel Name"]:
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                rd, wr = num(d.get("dram__bytes_read.sum", "")), num(d.get("dram__bytes_write.sum", ""))
                if rd is not None and wr is not None:
                    b = rd * scale.get(u.get("dram__bytes_read.sum", "byte"), 1.0) + wr * scale.get(u.get("dram__bytes_write.sum", "byte"), 1.0)
                    a = acc.setdefault(name, {"dram": [], "inst": []})
                    a["dram"].append(b)
                    i = num(d.get("smsp__inst_executed.sum", ""))
                    if i is not None:
                        a["inst"].append(i)
        print(f"\n== {d['Kernel Name'][:110]}")
        for k, label in KEYS:
            if k in d and d[k] != "":
                print(f"   {label:34s} {d[k]:>16s} {u.get(k, '')}")


    if traffic_path:
        try:
            out = json.load(open(traffic_path))
        except Exception:
            out = {}
        for name, a in acc.items():
            out[name] = {"dram_bytes": sum(a["dram"]) / len(a["dram"]),
                         "warp_instructions": (sum(a["inst"]) / len(a["inst"])) if a["inst"] else None,
                         "launches": len(a["dram"]),
                         "source": f"ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per launch ({tag or path})"}
        json.dump(out, open(traffic_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[2] == "--traffic":
        main(sys.argv[1], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        main(sys.argv[1])
