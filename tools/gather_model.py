"""Count, on the CPU, what one warp-wide gather of k_pair_forces touches under different lane->work mappings.

The force kernel is bound by L1 data-pipe wavefronts. This script rebuilds the product's
data layout for an LJ liquid (extended cell grid with ghost shell, entries sorted by cell, rows ascending in the
sorted index) and counts, per warp-gather, the distinct 32-byte sectors, the distinct 128-byte lines, and the
lines summed over 4-lane groups (the way a 256-bit-per-lane request is likely split), for

  * default  : 32 lanes = 32 consecutive sorted entries, lane follows its own row (tile layout)
  * rows G   : G lanes share one atom and read G consecutive row entries (EMDEE_ROWS=G)

It predicts the RATIO of data-pipe work between the mappings under each hardware hypothesis; tools/lsu_probe.cu
tells which hypothesis holds. CPU only (uses the oracle to melt the lattice); nothing here is shipped.

    python tools/gather_model.py [--ncell 16] [--steps 150]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as cm  # noqa: E402


def melted_box(ncell, steps):
    R, L = cm.fcc_lj_box(ncell, rho=0.8442, jitter=0.05, seed=1)
    lib = cm.oracle(fast=True)
    s = lib.system(os.cpu_count() or 1, 1, 2.5, 0.3, R.shape[0], None, None, None)
    s.set_pair_model(1, 1, lib.EmDee_pair_lj_cut(1.0, 1.0), 0.0)
    s.upload("box", [L])
    s.upload("coordinates", R)
    s.random_momenta(1.44, True, 1234)
    for _ in range(steps):
        s.boost(1.0, 0.0, 0.0025)
        s.displace(1.0, 0.0, 0.005)
        s.boost(1.0, 0.0, 0.0025)
    R = s.download("coordinates")
    s.finalize()
    return R, L


def build_layout(R, L, xRc):
    """Sorted entries (real + ghost images in a 2-cell shell) and, per real entry, its ascending neighbor row."""
    M = max(int(np.floor(2 * L / xRc)), 5)
    Mx = M + 4
    Rs = R / L
    Rs = Rs - np.floor(Rs)
    cell = np.minimum((Rs * M).astype(np.int64), M - 1)            # (N,3) home cell
    ent_pos, ent_cell, ent_real = [], [], []
    for sx in (-1, 0, 1):
        for sy in (-1, 0, 1):
            for sz in (-1, 0, 1):
                shift = np.array([sx, sy, sz])
                c = cell + shift * M + 2                           # extended-grid coordinates
                ok = np.all((c >= 0) & (c < Mx), axis=1)
                ent_pos.append(Rs[ok] + shift)
                ent_cell.append(c[ok])
                ent_real.append(np.full(ok.sum(), sx == 0 and sy == 0 and sz == 0))
    pos = np.concatenate(ent_pos)
    c = np.concatenate(ent_cell)
    real = np.concatenate(ent_real)
    lin = (c[:, 2] * Mx + c[:, 1]) * Mx + c[:, 0]
    order = np.argsort(lin, kind="stable")
    pos, c, real, lin = pos[order], c[order], real[order], lin[order]
    ncell = Mx ** 3
    start = np.searchsorted(lin, np.arange(ncell + 1))
    r2max = (xRc / L) ** 2
    rows = [None] * len(pos)
    for e in np.nonzero(real)[0]:
        cx, cy, cz = c[e]
        out = []
        for dz in range(-2, 3):
            for dy in range(-2, 3):
                base = ((cz + dz) * Mx + (cy + dy)) * Mx + cx
                lo, hi = start[base - 2], start[base + 3]          # x-run of 5 cells: contiguous entries
                if hi > lo:
                    d = pos[lo:hi] - pos[e]
                    m = (d * d).sum(1) < r2max
                    idx = np.nonzero(m)[0] + lo
                    out.append(idx[idx != e])
        rows[e] = np.concatenate(out) if out else np.zeros(0, np.int64)
    return pos, real, rows


def count(gathers):
    """gathers: iterable of int arrays (entry indices read by the active lanes of one warp-gather, lane order)."""
    n = sectors = lines = quad = pairs = oct16 = 0
    for g, lanes in gathers:
        n += 1
        pairs += len(g)
        sectors += len(np.unique(g))
        lines += len(np.unique(g >> 2))
        for q in range(8):                                          # 4-lane groups of a 256-bit request
            sel = g[(lanes >> 2) == q]
            if len(sel):
                quad += len(np.unique(sel >> 2))
        for q in range(4):                                          # 16-byte records: 8-lane groups, 8 records per 128-byte line
            sel = g[(lanes >> 3) == q]
            if len(sel):
                oct16 += len(np.unique(sel >> 3))
    return dict(gathers=n, pairs=pairs, sectors=sectors / n, lines=lines / n, quadlines=quad / n, oct16=oct16 / n,
                lanes_active=pairs / (32.0 * n))


def default_mapping(real, rows, nwarps, rng):
    Next = len(real)
    tiles = rng.choice(np.arange(Next // 32), size=min(nwarps, Next // 32), replace=False)
    for t in tiles:
        es = np.arange(32 * t, 32 * t + 32)
        rr = [rows[e] if real[e] else np.zeros(0, np.int64) for e in es]
        mx = max(len(r) for r in rr)
        for k in range(mx):
            lanes = np.array([i for i in range(32) if k < len(rr[i])])
            yield np.array([rr[i][k] for i in lanes]), lanes


def rows_mapping(real, rows, G, nwarps, rng):
    apw = 32 // G
    Next = len(real)
    warps = rng.choice(np.arange(Next // apw), size=min(nwarps, Next // apw), replace=False)
    for w in warps:
        es = np.arange(apw * w, apw * w + apw)
        rr = [rows[e] if real[e] else np.zeros(0, np.int64) for e in es]
        mx = max(len(r) for r in rr)
        if mx == 0:
            continue
        for k0 in range(0, mx, G):
            g, lanes = [], []
            for a in range(apw):
                seg = rr[a][k0:k0 + G]
                g.extend(seg)
                lanes.extend(a * G + np.arange(len(seg)))
            if g:
                yield np.array(g), np.array(lanes)


def cluster_mapping(real, rows, C, nwarps, rng):
    """C consecutive entries share one UNION row; lane (a, s) handles atom a and union slot s of 32/C per
    gather, computing only if the slot's entry is in atom a's own row (mask). Yields also the useful pairs."""
    J = 32 // C
    Next = len(real)
    warps = rng.choice(np.arange(Next // C), size=min(nwarps, Next // C), replace=False)
    for w in warps:
        es = np.arange(C * w, C * w + C)
        rr = [rows[e] if real[e] else np.zeros(0, np.int64) for e in es]
        if max(len(r) for r in rr) == 0:
            continue
        union = np.unique(np.concatenate(rr))
        member = [np.isin(union, r) for r in rr]
        for k0 in range(0, len(union), J):
            seg = union[k0:k0 + J]
            useful = sum(int(m[k0:k0 + J].sum()) for m in member)
            # every lane of a slot column reads the same entry: the request touches len(seg) sectors
            yield seg, np.arange(len(seg)) * C, useful


def count_cluster(gathers):
    n = sectors = lines = useful = 0
    for seg, lanes, u in gathers:
        n += 1
        sectors += len(seg)
        lines += len(np.unique(seg >> 2))
        useful += u
    return dict(gathers=n, sectors=sectors / n, lines=lines / n, lanes_active=useful / (32.0 * n))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ncell", type=int, default=16)
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--warps", type=int, default=600)
    args = ap.parse_args()
    R, L = melted_box(args.ncell, args.steps)
    pos, real, rows = build_layout(R, L, 2.8)
    nreal = int(real.sum())
    avg = sum(len(rows[e]) for e in np.nonzero(real)[0]) / nreal
    print(f"# {nreal} atoms, {len(real)} sorted entries (ghost shell included), {avg:.1f} list entries per atom (full list)")
    print(f"# per warp-gather: distinct 32B sectors | distinct 128B lines | lines summed over 4-lane groups | active lanes")
    rng = np.random.default_rng(0)
    res = {"default (lane = atom)": count(default_mapping(real, rows, args.warps, rng))}
    for G in (2, 4, 8, 16, 32):
        res[f"rows G={G:<2d} (G lanes per atom)"] = count(rows_mapping(real, rows, G, args.warps * (32 // G), rng))
    base = res["default (lane = atom)"]
    for C in (2, 4, 8):
        r = count_cluster(cluster_mapping(real, rows, C, args.warps * 4, rng))
        r["quadlines"] = float("nan")
        res[f"cluster {C} atoms x {32 // C:<2d} slots"] = r
    print("# last column: 16-byte records (EMDEE_REC16): 128-byte line segments summed over 8-lane groups, per pair, relative to the "
          "default mapping's quad-lines with 32-byte records")
    for name, r in res.items():
        per_pair = {k: r[k] / (32 * r["lanes_active"]) for k in ("sectors", "lines", "quadlines")}
        bp = {k: base[k] / (32 * base["lanes_active"]) for k in ("sectors", "lines", "quadlines")}
        print(f"{name:28s} sectors {r['sectors']:5.1f}  lines {r['lines']:5.1f}  quad-lines {r['quadlines']:5.1f}  "
              f"active {100 * r['lanes_active']:4.0f}%   per pair vs default: sectors x{per_pair['sectors'] / bp['sectors']:.2f} "
              f"lines x{per_pair['lines'] / bp['lines']:.2f} quad-lines x{per_pair['quadlines'] / bp['quadlines']:.2f}"
              + (f"   rec16 oct-lines {r['oct16']:5.1f} x{r['oct16'] / (32 * r['lanes_active']) / bp['quadlines']:.2f}" if 'oct16' in r else ""))


if __name__ == "__main__":
    main()
