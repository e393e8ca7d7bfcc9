#!/usr/bin/env python
"""Static SASS opcode counts (cuobjdump -sass) of the cubins bench.py runs, the fused column kernel and tuned advection:
the evidence that tile movement is TMA (UTMALDG), synchronisation is mbarrier-based (SYNCS) and what the store path
looks like.  Writes profiles/r02_sass_opcodes.json."""
import collections
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def kernels(cubin):
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True, check=True).stdout
    out, name = {}, None
    for line in sass.splitlines():
        line = line.strip()
        if line.startswith("Function :"):
            name = line.split(":", 1)[1].strip()
            out[name] = collections.Counter()
        elif name and line.startswith("/*") and "*/" in line:
            body = line.split("*/", 1)[1].strip()
            if not body or body.startswith("/*"):
                continue
            tok = body.split()
            op = tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]
            out[name][op.rstrip(";").split(".")[0]] += 1
    return out


def main():
    from ab_test import candidate
    from absinthe_b200.generator import build_variant
    column = {"_clean": True, "_assignment": [0] * 9, "STREAM_K": "regs", "SMEM_BUDGET": 232448, "FUSE_HALO": True,
              "WAIT_SLEEP": 100, "_shapes": [[57, 9, 80]], "BLOCK": [1, 1],
              "COLUMN": {"WARPS": [2, 9], "UNROLL": 1, "SLOTS": 3, "MERGE": True, "CLUSTER": [2, 1], "SLACK": 2}}
    result = {"_comment": "static SASS opcode counts (cuobjdump -sass, tools/sass_opcodes.py) of the cubins bench.py runs, "
                          "the fused column kernel in clusters of two (DESIGN.md section 3b) and tuned advection; UTMALDG = "
                          "cp.async.bulk.tensor (TMA) loads, SYNCS = mbarrier operations, UCGABAR = cluster barrier; there is "
                          "no UTMASTG: outputs leave through STG.64 (DESIGN.md section 3b)"}
    for label, kernel, extra in (("diffusion", "diffusion", {}), ("fastwaves", "fastwaves", {}),
                                 ("advection", "advection", {}), ("fastwaves-column-cluster2", "fastwaves", column)):
        _, cubin = build_variant(candidate(kernel, extra))
        for name, ops in kernels(cubin).items():
            result["%s:%s" % (label, name)] = {
                "instructions": sum(ops.values()), "top": dict(ops.most_common(14)),
                **{k: sum(v for op, v in ops.items() if op.startswith(k)) for k in
                   ("UTMALDG", "UTMASTG", "SYNCS", "SHFL", "STG", "R2UR", "UCGABAR")}}
    path = os.path.join(ROOT, "profiles", "r02_sass_opcodes.json")
    json.dump(result, open(path, "w"), indent=1)
    print(json.dumps({k: v for k, v in result.items() if k != "_comment"}, indent=None)[:1500])


if __name__ == "__main__":
    main()
