#!/usr/bin/env python
"""force_lab.py -- times variants of the plain-LJ pair kernel back to back on ONE resident LJ-1M system (development tool).

The variants are compiled into libemdee.so (EMDEE_LJ_VARIANTS in engine.cu) and selected at run time through
EmDeeX_tune("force_variant" | "carveout"): same list, same coordinates, same step loop for every variant, so the numbers
are comparable inside one GPU call. Prints one line per (variant, carveout): kernel ms per launch (CUDA events from the
library's timer ring), ms per MD step, list builds.

    python tools/force_lab.py [--variants 0,1,2] [--carveouts -1,0] [--steps 40] [--ncell 63]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from emdee_b200 import api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="0")
    ap.add_argument("--carveouts", default="-1")
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--ncell", type=int, default=63)
    args = ap.parse_args()
    lib = api.load()
    R, P, L = bench.make_workload(args.ncell)
    s = bench.build_system(lib, R, P, L, 1)
    s.set_kernel_timing(True)
    for _ in range(args.warmup):
        bench.md_step(s)
    N = R.shape[0]
    print(f"# {N} atoms, {args.steps} MD steps per variant after {args.warmup} warm-up; columns: variant carveout | "
          f"force kernel ms/launch | build kernel ms/launch | ms/step | step minus kernels | builds | U", flush=True)
    for c in [int(x) for x in args.carveouts.split(",")]:
        for v in [int(x) for x in args.variants.split(",")]:
            lib.EmDeeX_tune(s.md, b"force_variant", v)
            lib.EmDeeX_tune(s.md, b"carveout", c)
            for _ in range(3):
                bench.md_step(s)
            s.synchronize()
            st0, b0 = s.stats(), s.md.Builds
            t0 = time.perf_counter()
            for _ in range(args.steps):
                bench.md_step(s)
            s.synchronize()
            dt = time.perf_counter() - t0
            st1 = s.stats()
            fl = max(st1.force_launches - st0.force_launches, 1)
            bl = st1.build_launches - st0.build_launches
            fms = (st1.force_ms - st0.force_ms) / fl
            bms = (st1.build_ms - st0.build_ms) / max(bl, 1)
            step = 1e3 * dt / args.steps
            print(f"variant {v:3d} carveout {c:4d} | force {fms:.4f} ms | build {bms:.4f} ms | step {step:.4f} ms | "
                  f"force launches/step {fl / args.steps:.2f} | builds {s.md.Builds - b0} | U {s.md.Energy.Potential:.6f}", flush=True)
    lib.EmDeeX_tune(s.md, b"force_variant", 0)
    lib.EmDeeX_tune(s.md, b"carveout", -1)
    # per-kernel device time of the default step (CUDA events from the library's ring), in-situ (warm caches)
    k0 = s.kernel_times()
    b0 = s.md.Builds
    t0 = time.perf_counter()
    for _ in range(2 * args.steps):
        bench.md_step(s)
    s.synchronize()
    dt = time.perf_counter() - t0
    k1 = s.kernel_times()
    nst = 2 * args.steps
    line = []
    tot = 0.0
    for kind in s.KERNEL_KINDS:
        ms, n = k1[kind][0] - k0[kind][0], k1[kind][1] - k0[kind][1]
        if n:
            line.append(f"{kind} {ms / n * 1e3:.1f} us x {n / nst:.2f}/step")
            tot += ms
    print(f"# default step, {nst} steps, {s.md.Builds - b0} builds: {1e3 * dt / nst:.4f} ms/step wall; kernels: " + "; ".join(line) +
          f"; timed kernels sum {tot / nst:.4f} ms/step", flush=True)
    s.finalize()


if __name__ == "__main__":
    main()
