#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers the design decisions rest on."""
import csv, subprocess, sys, json
def main(path, out=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
            "launch__occupancy_limit_registers", "sm__cycles_elapsed.max", "sm__inst_executed.sum",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
            "smsp__inst_executed.avg.per_cycle_active", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
            "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum",
            "smsp__cycles_active.avg"]
    result = []
    for vals in rows[2:]:
        d = {}
        for i, h in enumerate(hdr):
            if h in keys:
                d[h] = vals[i] + (" " + units[i] if units[i] else "")
            elif h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                # warps stalled for this reason per issued instruction (the "Warp State" chart of the ncu UI)
                try:
                    if float(vals[i]) >= 0.1:
                        name = h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
                        d.setdefault("stalled_warps_per_issue", {})[name] = round(float(vals[i]), 3)
                except ValueError:
                    pass
        result.append(d)
    text = json.dumps(result, indent=1)
    print(text)
    if out:
        open(out, "w").write(text)
if __name__ == "__main__":
    main(*sys.argv[1:])
