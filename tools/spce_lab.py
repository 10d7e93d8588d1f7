#!/usr/bin/env python
"""spce_lab.py -- times launch shapes of the typed (SPC/E) pair kernel back to back on one resident system (development tool).
    python tools/spce_lab.py [--variants 0,1] [--replicas 8]
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
    ap.add_argument("--variants", default="0,1")
    ap.add_argument("--replicas", type=int, default=8)
    ap.add_argument("--evals", type=int, default=6)
    args = ap.parse_args()
    lib = api.load()
    s, c = cm.spce_sample_system(lib, lambda l: l.EmDee_coul_damped_square_smoothed(0.2, 1.0), replicas=args.replicas, threads=1)
    s.set_kernel_timing(True)
    s.md.Options.Compute = True
    print(f"# {c['R'].size // 3} atoms; columns: variant | force kernel ms | U | W", flush=True)
    for v in [int(x) for x in args.variants.split(",")] * 2:
        lib.EmDeeX_tune(s.md, b"force_variant", v)
        s.upload("coordinates", c["R"]); s.compute_forces()
        s.synchronize()
        st0 = s.stats()
        for _ in range(args.evals):
            s.upload("coordinates", c["R"]); s.compute_forces()
        s.synchronize()
        st1 = s.stats()
        fl = max(st1.force_launches - st0.force_launches, 1)
        print(f"typed variant {v} | force {(st1.force_ms - st0.force_ms) / fl:.4f} ms | U {s.md.Energy.Potential!r} | W {s.md.Virial.Total!r}", flush=True)
    lib.EmDeeX_tune(s.md, b"force_variant", 0)
    s.finalize()


if __name__ == "__main__":
    main()
