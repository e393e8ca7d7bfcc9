#!/usr/bin/env python
"""Model-vs-measurement study: does the analytical model, re-parameterised
for the B200, rank implementation variants like the GPU does?

  plan    solve the MILP (absinthe_b200/optimizer.py, HiGHS, in parallel) for a spread of fusion groupings -> variants JSON (tiling + estimate)
  measure (GPU box) time every variant of a variants JSON -> Spearman(estimate, measured)
"""
import argparse
import json
import os
import random
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def plan(args):
    from absinthe_b200 import programs as P
    from absinthe_b200.driver import model_program, solve_all
    domain = tuple(int(v) for v in args.domain.split("x"))
    seq = P.KERNELS[args.kernel][1]()
    rng = random.Random(args.seed)
    assignments = [[0] * len(seq), list(range(len(seq)))]
    while len(assignments) < args.count:
        a = [0]
        for _ in seq[1:]:
            a.append(a[-1] + rng.choice([0, 0, 1]))
        if a not in assignments:
            assignments.append(a)
    os.makedirs(args.work, exist_ok=True)
    jobs = []
    for n, a in enumerate(assignments):
        name = "%s-%s" % (args.kernel, "-".join(map(str, a)))
        jobs.append(("v%d" % n, os.path.join(args.work, ""), model_program(args.kernel, name, groups=list(zip(seq, a)),
                                                                         domain=domain)))
    variants = []
    for (key, program), a in zip(solve_all(jobs), assignments):
        if "TILING" not in program or "OBJECTIVE" not in program:
            print("skip (no incumbent within the time limit)", a, program.get("ERROR", ""))
            continue
        tiles = [[g["GROUPS"][0][k] for k in ("NX", "NY", "NZ")] for g in program["TILING"]["GROUPS"]]
        variants.append({"assignment": a, "tiles": tiles, "estimate": float(program["OBJECTIVE"])})
        print(program["VARIANT"], tiles, program["OBJECTIVE"], flush=True)
    with open(args.out, "w") as handle:
        json.dump({"kernel": args.kernel, "domain": list(domain), "machine": P.base_program(args.kernel)["MACHINE"],
                   "memory": P.base_program(args.kernel)["MEMORY"], "cache": P.base_program(args.kernel)["CACHE"],
                   "variants": variants}, handle, indent=1)


def _compile(job):
    from absinthe_b200.generator import build_variant
    try:
        build_variant(job)
        return None
    except Exception as exc:
        return str(exc)[-300:]


def build(args):
    """Compile every variant of a study (no GPU needed): the cache travels to the GPU box with the snapshot."""
    from concurrent.futures import ProcessPoolExecutor
    from absinthe_b200 import programs as P
    study = json.load(open(args.variants))
    kernel, domain = study["kernel"], tuple(study["domain"])
    seq = P.KERNELS[kernel][1]()
    jobs = []
    for item in study["variants"]:
        program = P.make_program(kernel, P.split_sequence(seq, item["assignment"]), [tuple(t) for t in item["tiles"]],
                                 domain=domain)
        program["CUDA"] = json.loads(args.cuda) if args.cuda else {}
        jobs.append(program)
    with ProcessPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        errors = [e for e in pool.map(_compile, jobs) if e]
    print(len(jobs), "variants,", len(errors), "failures", errors[:3])


def measure(args):
    import torch
    from absinthe_b200 import programs as P
    from absinthe_b200.fit import spearman
    from absinthe_b200.launcher import Variant, make_periodic
    study = json.load(open(args.variants))
    kernel, domain = study["kernel"], tuple(study["domain"])
    seq = P.KERNELS[kernel][1]()
    cuda = json.loads(args.cuda) if args.cuda else {}
    tensors, est, got = {}, [], []
    gen = torch.Generator(device="cuda").manual_seed(3)
    for item in study["variants"]:
        groups = P.split_sequence(seq, item["assignment"])
        program = P.make_program(kernel, groups, [tuple(t) for t in item["tiles"]], domain=domain)
        program["CUDA"] = dict(cuda)
        try:
            v = Variant.from_program(program)
        except Exception as exc:
            print("skip", item["assignment"], str(exc)[-200:])
            continue
        dims = (v.X, v.Y, v.Z, v.HX, v.HY, v.HZ)
        for name in v.inputs + v.outputs:
            if name not in tensors:
                t = torch.rand(v.shape, dtype=torch.float64, device="cuda", generator=gen) + 0.5
                make_periodic(t, dims)
                tensors[name] = t
        v.bind([tensors[n] for n in v.inputs], [tensors[n] for n in v.outputs])
        total, halo = v.run(args.runs, flush=False)
        v.close()
        item["measured_ms"] = statistics.median(total)
        est.append(item["estimate"])
        got.append(item["measured_ms"])
        print(item["assignment"], item["tiles"], "est %.4f  measured %.4f ms" % (item["estimate"], item["measured_ms"]),
              flush=True)
    study["spearman"] = spearman(est, got)
    print("Spearman(estimate, measured) over %d variants: %.3f" % (len(est), study["spearman"]))
    if args.out:
        with open(args.out, "w") as handle:
            json.dump(study, handle, indent=1)


def main():
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    a = sub.add_parser("plan")
    a.add_argument("--kernel", required=True)
    a.add_argument("--domain", default="1024x1024x80")
    a.add_argument("--count", type=int, default=16)
    a.add_argument("--seed", type=int, default=1)
    a.add_argument("--work", default="/tmp/model_study")
    a.add_argument("--out", required=True)
    b = sub.add_parser("measure")
    b.add_argument("--variants", required=True)
    b.add_argument("--runs", type=int, default=4)
    b.add_argument("--cuda", default="")
    b.add_argument("--out", default="")
    c = sub.add_parser("build")
    c.add_argument("--variants", required=True)
    c.add_argument("--cuda", default="")
    args = ap.parse_args()
    {"plan": plan, "measure": measure, "build": build}[args.cmd](args)


if __name__ == "__main__":
    main()
