#!/usr/bin/env python
"""What the time of a training program is made of, from hardware counters (north_star item 3).

Input: the ncu metric lists of a sample of the stride-1 training sweeps (tools/runs/r2_training.sh: every 8th
program of fitcache / fitddr under ``ncu --metrics gpu__time_duration.sum, lts__t_bytes.sum, dram__bytes.sum,
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum, smsp__inst_executed.sum``).  Per kernel launch it regresses the
duration on those counters (non-negative least squares with a constant = launch latency), which splits the model's
"cache" cost into a shared-memory term (wavefronts), an L2 term (bytes) and an issue term (warp instructions) and
its "memory" cost into HBM bytes.

    python tools/fit_counters.py gpurun_out/r2_fitcache_counters.csv gpurun_out/r2_fitddr_counters.csv
"""
import csv
import json
import sys

import numpy as np

METRICS = ["gpu__time_duration.sum", "lts__t_bytes.sum", "dram__bytes.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum"]
SCALE = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6,
         "Gbyte": 1e9, "": 1.0, "inst": 1.0}


def read(path):
    """ncu --csv metric list -> [{metric: value}] per kernel launch (durations in ms, bytes in bytes)."""
    launches = {}
    with open(path) as handle:
        rows = [line for line in handle if line.startswith('"')]
    for rec in csv.DictReader(rows):
        name = rec["Metric Name"]
        if name in METRICS:
            value = float(rec["Metric Value"].replace(",", "")) * SCALE.get(rec["Metric Unit"], 1.0)
            launches.setdefault(int(rec["ID"]), {"kernel": rec["Kernel Name"]})[name] = value
    return [v for _, v in sorted(launches.items()) if all(m in v for m in METRICS)]


def fit(samples):
    from scipy.optimize import nnls
    t = np.array([s[METRICS[0]] for s in samples])
    cols = np.stack([np.array([s[m] for s in samples]) for m in METRICS[1:]] + [np.ones(len(samples))], axis=1)
    scale = cols.max(axis=0)
    coef, _ = nnls(cols / scale, t)
    coef = coef / scale
    pred = cols @ coef
    r2 = 1.0 - float(((t - pred) ** 2).sum() / ((t - t.mean()) ** 2).sum())
    share = {m: float((cols[:, n] * coef[n]).sum() / pred.sum()) for n, m in enumerate(METRICS[1:] + ["launch"])}
    return {"samples": len(samples), "R2": round(r2, 4),
            "ms_per_L2_GB": coef[0] * 1e9, "ms_per_HBM_GB": coef[1] * 1e9, "ms_per_G_shared_wavefronts": coef[2] * 1e9,
            "ms_per_G_warp_instructions": coef[3] * 1e9, "launch_ms": coef[4],
            "implied_GBps": {"L2": (1.0 / (coef[0] * 1e6)) if coef[0] else None,
                             "HBM": (1.0 / (coef[1] * 1e6)) if coef[1] else None},
            "share_of_predicted_time": {k: round(v, 3) for k, v in share.items()}}


def main():
    out = {}
    for path in sys.argv[1:]:
        samples = read(path)
        if samples:
            out[path.split("/")[-1]] = fit(samples)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
