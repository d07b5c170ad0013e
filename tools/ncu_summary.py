"""Summarise an .ncu-rep (one kernel launch) into the handful of numbers DESIGN.md and
bench.py's roofline.traffic cite.
Usage: python tools/ncu_summary.py rep [rep...]                       -> JSON summary on stdout
       python tools/ncu_summary.py --traffic SHA rep [rep...]        -> also rewrites profiles/traffic.json: the DRAM
           bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of every capture whose file name is
           <prefix>prof_<bench workload name>.ncu-rep, stamped with SHA (the git revision of the build that was
           profiled) and the date; entries of workloads that were not re-captured are kept and listed as older."""
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_dispatch_stall"]

UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}


def summarise(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS or "dmma" in h.lower():
            try:
                x = float(v.replace(",", ""))
            except ValueError:
                continue
            if u in UNIT and ("bytes" in h or "duration" in h):
                x *= UNIT[u]
                u = "B" if "bytes" in h else "s"
            d[h] = [x, u]
    if "dram__bytes_read.sum" in d:
        d["dram_traffic_bytes"] = d["dram__bytes_read.sum"][0] + d["dram__bytes_write.sum"][0]
    return d


if __name__ == "__main__":
    import datetime
    import os
    import re
    args = sys.argv[1:]
    sha = None
    if args and args[0] == "--traffic":
        sha, args = args[1], args[2:]
    res = {p: summarise(p) for p in args}
    print(json.dumps(res, indent=1))
    if sha:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        path = os.path.join(root, "profiles", "traffic.json")
        try:
            with open(path) as f:
                traffic = json.load(f)
        except Exception:
            traffic = {}
        old_src = traffic.pop("_source", None)
        fresh = []
        for p_, d in res.items():
            m = re.search(r"prof_(.+)\.ncu-rep$", os.path.basename(p_))
            if m and "dram_traffic_bytes" in d:
                traffic[m.group(1)] = d["dram_traffic_bytes"]
                fresh.append(m.group(1))
        traffic["_source"] = {"git_sha": sha, "date": datetime.date.today().isoformat(),
                              "how": "ncu --set full --clock-control none, one launch per workload, dram__bytes_read.sum + dram__bytes_write.sum",
                              "captured_at_this_sha": sorted(fresh), "older_entries_from": old_src}
        with open(path, "w") as f:
            json.dump(traffic, f, indent=1)
