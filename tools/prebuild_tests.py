"""Compile every variant the GPU tests use, in parallel, so the in-tree cache travels to the GPU box."""
import os, sys
from concurrent.futures import ProcessPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

def build(program):
    from absinthe_b200.generator import build_variant
    try:
        build_variant(program); return None
    except Exception as exc:
        return "%s %s: %s" % (program.get("VARIANT"), program.get("CUDA"), str(exc)[-1500:])

def main():
    from absinthe_b200 import programs as P
    from common import zoo
    import test_parity_gpu as T
    jobs = []
    for name, program in zoo():
        jobs.append(P.clone(program))
        q = P.clone(program); q["CUDA"] = {"LOAD": "ldg"}; jobs.append(q)
        for _, cuda in T.OPTIONS:
            q = P.clone(program); q["CUDA"] = dict(cuda); jobs.append(q)
        q = P.clone(program); q["CUDA"] = {"INLINE": T.PARTIAL_INLINE[program["NAME"]], "BLOCK": (2, 2), "FUSE_HALO": True}; jobs.append(q)
        q = P.clone(q); q["CUDA"]["SHUFFLE"] = True; jobs.append(q)
        q = P.clone(q); q["CUDA"].update(STREAM_K=True, BLOCK=(2, 1)); jobs.append(q)
    for _, program in T.SHIPPED:
        jobs.append(P.clone(program))
        q = P.clone(program); q["CUDA"] = {"INLINE": "all", "BLOCK": (2, 1), "SHUFFLE": True, "FUSE_HALO": True}; jobs.append(q)
    with ProcessPoolExecutor(max_workers=os.cpu_count()) as pool:
        errors = [e for e in pool.map(build, jobs) if e]
    print(len(jobs), "variants,", len(errors), "failures")
    for e in errors[:5]: print(e)
    return 1 if errors else 0

if __name__ == "__main__":
    sys.exit(main())
