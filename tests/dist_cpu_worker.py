"""world_size-2 gloo worker (CPU): host-side logic of the N>1 path -- slab arithmetic against the library's own
helper, the NCCL-id broadcast plumbing, pair gathering, max-over-ranks timing -- driven with the CPU oracle as
the compute backend of each rank's replica."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("EMDEE_QUIET", "1")

import ctypes as C  # noqa: E402

import common as cm  # noqa: E402
from emdee_b200 import dist as edist  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = cm.product()          # loads without a GPU; only host-only entry points are called here
    # 1. slab arithmetic: python mirror == library helper, slabs tile [0, M) without gaps
    for M in (5, 6, 37, 75, 302):
        edges = []
        for r in range(world):
            z0, z1 = C.c_int(), C.c_int()
            lib.EmDeeX_slab_range(M, r, world, C.byref(z0), C.byref(z1))
            assert (z0.value, z1.value) == edist.slab_range(M, r, world)
            edges.append((z0.value, z1.value))
        assert edges[0][0] == 0 and edges[-1][1] == M
        assert all(edges[k][1] == edges[k + 1][0] for k in range(world - 1))
    assert edist.cells_per_dim(105.815, 2.5, 0.3) == 75
    # 2. id broadcast plumbing (a fake 128-byte id)
    payload = bytes(range(128)) if rank == 0 else bytes(128)
    got = edist.broadcast_bytes(payload, 128, 0)
    assert got == bytes(range(128))
    # 3. each rank owns the pairs whose lower atom lies in its slab: the gathered union is the oracle's set
    orc = cm.oracle()
    s, c = cm.lj_sample_system(orc, lambda l, e, sg: l.EmDee_pair_lj_cut(e, sg))
    pairs = s.pairs()
    M = edist.cells_per_dim(c["L"], c["Rc"], c["Rs"])
    z0, z1 = edist.slab_range(M, rank, world)
    zs = c["R"][:, 2] / c["L"]
    layer = np.minimum((M * (zs - np.floor(zs))).astype(int), M - 1)
    mine = pairs[(layer[pairs[:, 0]] >= z0) & (layer[pairs[:, 0]] < z1)]
    allp = edist.gather_pairs(mine)
    assert np.array_equal(allp, pairs)
    # 4. timing reduction
    assert edist.max_over_ranks(1.0 + rank) == float(world)
    s.finalize()
    dist.barrier()
    if rank == 0:
        print("[dist-cpu] ALL OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
