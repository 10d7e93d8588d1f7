#!/usr/bin/env python
"""spce_lab.py -- times the variants of the typed pair kernel (EmDeeX_tune "typed_variant") on ONE resident SPC/E box
(NIST sample x n^3, rigid bodies, LJ shifted-force on O + coul_damped_square_smoothed): force evaluations on the same
coordinates, kernel ms from the library's CUDA-event ring, energies printed so that the variants can be compared.

    python tools/spce_lab.py [--variants 0,1,2,3] [--replicas 8] [--evals 10]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as cm  # noqa: E402
from emdee_b200 import api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="0,1,2,3")
    ap.add_argument("--replicas", type=int, default=8)
    ap.add_argument("--evals", type=int, default=10)
    args = ap.parse_args()
    lib = api.load()
    s, c = cm.spce_sample_system(lib, lambda l: l.EmDee_coul_damped_square_smoothed(0.2, 1.0), replicas=args.replicas, threads=1)
    s.set_kernel_timing(True)
    print(f"# {c['N']} atoms; columns: variant | force kernel ms | U | W", flush=True)
    for v in [int(x) for x in args.variants.split(",")]:
        lib.EmDeeX_tune(s.md, b"typed_variant", v)
        for _ in range(2):
            s.upload("coordinates", c["R"])
            s.compute_forces()
        k0 = s.kernel_times()["force"]
        for _ in range(args.evals):
            s.upload("coordinates", c["R"])
            s.compute_forces()
        k1 = s.kernel_times()["force"]
        print(f"typed variant {v} | force {(k1[0] - k0[0]) / max(k1[1] - k0[1], 1):.4f} ms | U {s.md.Energy.Potential!r} | W {s.md.Virial.Total!r}", flush=True)
    lib.EmDeeX_tune(s.md, b"typed_variant", 0)
    s.finalize()


if __name__ == "__main__":
    main()
