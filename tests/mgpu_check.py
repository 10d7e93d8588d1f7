"""Run under torchrun with one rank per GPU (tests/test_multi_gpu.py launches it; tests/test_emulated_multi_rank.py
runs the same script on the CPU through the kernel emulator): the slab-decomposed CUDA path
must give the single-process oracle's results -- pair sets bit-exact (union over ranks), forces, energies,
virial, rebuild counts -- on a static configuration and along a short trajectory."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("EMDEE_QUIET", "1")

import common as cm  # noqa: E402
from emdee_b200 import dist as edist  # noqa: E402


def build_lj(lib, comm, ncell=None, charged=False):
    if ncell is None:
        ncell = int(os.environ.get("EMDEE_MGPU_NCELL", "14"))   # 14 -> 16 cell layers; 8 ranks need >= 21 (25 layers)
    R, L = cm.fcc_lj_box(ncell, rho=0.8442, jitter=0.06, seed=5)
    N = R.shape[0]
    s = lib.system(2, 1, 2.5, 0.3, N, None, None, None)
    if comm:
        edist.init_comm(lib, s)
    s.set_pair_model(1, 1, lib.EmDee_pair_lj_cut(1.0, 1.0), 1.0 if charged else 0.0)
    if charged:
        s.set_coul_model(lib.EmDee_shifted_force(lib.EmDee_coul_cut()))
        q = np.where(np.arange(N) % 2 == 0, 0.5, -0.5)
        s.upload("charges", q)
    s.upload("box", np.array([L]))
    s.upload("coordinates", R)
    return s


def build_spce(lib, comm, replicas=2):
    def pre(lib_, s_):
        if comm:
            edist.init_comm(lib_, s_)
    # spce_sample_system creates the system and sets models before any upload; hook the comm in between
    orig = cm.api.System.set_pair_model
    state = {"done": False}

    def patched(self, *a, **k):
        if not state["done"]:
            pre(self.lib, self)
            state["done"] = True
        return orig(self, *a, **k)

    cm.api.System.set_pair_model = patched
    try:
        s, c = cm.spce_sample_system(lib, lambda l: l.EmDee_coul_damped_square_smoothed(0.2, 1.0), replicas=replicas)
    finally:
        cm.api.System.set_pair_model = orig
    return s


def pinned(shape):
    """Page-locked float64 host array (numpy view); on the CPU emulator any array will do."""
    if os.environ.get("EMDEE_MGPU_EMULATED") == "1":
        return np.empty(shape)
    t = torch.empty(shape, dtype=torch.float64).pin_memory()
    _keep.append(t)
    return t.numpy()


_keep = []


def compare(tag, sp, so, rank, ftol=1e-10, stol=1e-12):
    F = sp.download("forces")          # collective on the product side
    pairs = edist.gather_pairs(sp.pairs())
    if rank == 0:
        assert np.array_equal(pairs, so.pairs()), f"{tag}: pair sets differ ({pairs.shape} vs {so.pairs().shape})"
        err = cm.rel_force_error(F, so.download("forces"))
        assert err <= ftol, f"{tag}: force error {err:.3e}"
        sc = max(abs(so.md.Energy.Potential), abs(so.md.Energy.Coulomb), abs(so.md.Virial.Total))
        for a, b, nm in [(sp.md.Energy.Potential, so.md.Energy.Potential, "U"), (sp.md.Energy.Coulomb, so.md.Energy.Coulomb, "Ucoul"),
                         (sp.md.Virial.Total, so.md.Virial.Total, "W"), (sp.md.Virial.Body, so.md.Virial.Body, "Wbody")]:
            assert abs(a - b) <= stol * sc, f"{tag}: {nm} {a!r} vs {b!r}"
        print(f"[mgpu] {tag}: ok  pairs={pairs.shape[0]} force_err={err:.2e} U={sp.md.Energy.Potential:.6f}", flush=True)


def main():
    rank, world, local = edist.env_rank_world()
    emulated = os.environ.get("EMDEE_MGPU_EMULATED") == "1"
    if emulated:
        # CPU run: kernels through tests/cusim, collectives through the shared-memory NCCL stand-in (EMDEE_NCCL_LIB is
        # set by the launching test), python-side plumbing over gloo
        dist.init_process_group("gloo")
        lib = cm.emulated()
    else:
        torch.cuda.set_device(local)
        os.environ["EMDEE_DEVICE"] = str(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        lib = cm.product()
    dev = "cpu" if emulated else "cuda"
    orc = cm.oracle() if rank == 0 else None

    for charged in ((False,) if os.environ.get("EMDEE_MGPU_SKIP_CHARGED") == "1" else (False, True)):
        sp = build_lj(lib, True, charged=charged)
        so = build_lj(orc, False, charged=charged) if rank == 0 else None
        compare(f"lj charged={charged} static", sp, so, rank)
        sp.random_momenta(1.3, True, 777)
        if rank == 0:
            so.random_momenta(1.3, True, 777)
        for step in range(40):
            for s in ([sp, so] if rank == 0 else [sp]):
                s.md.Options.Compute = (step % 8 == 7)
                s.boost(1.0, 0.0, 0.0025)
                s.displace(1.0, 0.0, 0.005)
                s.boost(1.0, 0.0, 0.0025)
        b = torch.tensor([sp.md.Builds], device=dev)
        dist.all_reduce(b, op=dist.ReduceOp.MAX)
        assert int(b.item()) == sp.md.Builds
        if rank == 0:
            assert sp.md.Builds == so.md.Builds and sp.md.Builds > 2, (sp.md.Builds, so.md.Builds)
            assert abs(sp.md.Kinetic.Total - so.md.Kinetic.Total) <= 1e-10 * abs(so.md.Kinetic.Total)
        compare(f"lj charged={charged} after 40 steps ({sp.md.Builds} builds)", sp, so, rank, ftol=1e-8, stol=1e-10)
        Rp = sp.download("coordinates")
        if rank == 0:
            assert np.abs(Rp - so.download("coordinates")).max() < 1e-10
        if charged and os.environ.get("EMDEE_MGPU_SKIP_VERLET") != "1":
            # EmDee_verlet_step with the shadow-Hamiltonian bookkeeping (s0 of migrating atoms is all-reduced)
            for step in range(12):
                for s in ([sp, so] if rank == 0 else [sp]):
                    s.md.Options.Compute = True
                    s.verlet_step(0.005)
            if rank == 0:
                assert sp.md.Builds == so.md.Builds
                for grp, nm in (("Energy", "ShadowPotential"), ("Kinetic", "ShadowKinetic"), ("Kinetic", "Total"), ("Energy", "Potential")):
                    a, b = getattr(getattr(sp.md, grp), nm), getattr(getattr(so.md, grp), nm)
                    assert abs(a - b) <= 1e-9 * abs(b), f"verlet_step {grp}.{nm}: {a!r} vs {b!r}"
                print(f"[mgpu] verlet_step with shadow terms ok ({sp.md.Builds} builds)", flush=True)
        if not charged:
            # EmDee_rdf over the slab-decomposed list (collective: histogram all-reduced). The two trajectories differ
            # at rounding level after 40 steps, so a pair sitting on a bin edge may move one bin: compare counts
            gp = sp.rdf(40, 2.5, [1], [1])
            if rank == 0:
                go = so.rdf(40, 2.5, [1], [1])
                assert np.abs(gp - go).max() <= 2e-3 * np.abs(go).max(), np.abs(gp - go).max()
                assert abs(gp.sum() - go.sum()) <= 1e-6 * go.sum()
                print(f"[mgpu] rdf ok (max |dg| = {np.abs(gp - go).max():.2e})", flush=True)
        if not charged:
            # the host-buffer pattern of bench.py's e2e arm: every step a full set of coordinates comes from the host
            # (all ranks upload the same array), forces are computed and downloaded (collective)
            base = Rp.copy()
            rng = np.random.default_rng(17)
            for k in range(5):
                base = base + rng.normal(0.0, 0.05, base.shape)      # large enough to force rebuilds + migration
                for s in ([sp, so] if rank == 0 else [sp]):
                    s.upload("coordinates", base)
                    s.compute_forces()
                compare(f"lj host-buffer step {k}", sp, so, rank)
            if rank == 0:
                assert sp.md.Builds == so.md.Builds
            # local I/O (EmDeeX_tune "local_io"): every rank reads only its owned + halo atoms from its own host array and
            # writes back only the forces of the atoms it owns; the union over the ranks must be the oracle's force array
            lib.EmDeeX_tune(sp.md, b"local_io", 1)
            N = base.shape[0]
            h2d0, d2h0 = sp.io_bytes()
            for k in range(6):
                base = base + rng.normal(0.0, 0.03, base.shape)      # continuous motion: well under one cell layer per upload
                hb = pinned((N, 3))
                hb[:] = base
                sp.upload("coordinates", hb)
                sp.compute_forces()
                if rank == 0:
                    so.upload("coordinates", base)
                    so.compute_forces()
                Fl = pinned((N, 3))        # local I/O needs page-locked host memory (pageable arrays take the collective path)
                Fl[:] = np.nan
                sp.lib.EmDee_download(sp.md, b"forces", Fl.ctypes.data_as(cm.api._dp))
                mine = torch.from_numpy(np.isfinite(Fl[:, 0]).astype(np.float64)).to(dev)
                Fsum = torch.from_numpy(np.nan_to_num(Fl)).to(dev)
                dist.all_reduce(mine)
                dist.all_reduce(Fsum)
                assert float(mine.min()) == 1.0 and float(mine.max()) == 1.0, f"every atom must be written by exactly one rank: min {float(mine.min())} max {float(mine.max())} sum {float(mine.sum())} N {N} local finite {int(np.isfinite(Fl[:, 0]).sum())} {int(np.isfinite(Fl).sum())}"
                if rank == 0:
                    err = cm.rel_force_error(Fsum.cpu().numpy(), so.download("forces"))
                    assert err <= 1e-10, f"local I/O step {k}: force error {err:.3e}"
                    assert abs(sp.md.Energy.Potential - so.md.Energy.Potential) <= 1e-12 * abs(so.md.Energy.Potential)
            h2d1, d2h1 = sp.io_bytes()
            assert (h2d1 - h2d0) < 6 * 24 * N and (d2h1 - d2h0) < 6 * 24 * N, "local I/O must move less than the full arrays"
            if rank == 0:
                assert sp.md.Builds == so.md.Builds
                print(f"[mgpu] local I/O ok: {(h2d1 - h2d0) / 6 / (24 * N):.2f} of the coordinates up, "
                      f"{(d2h1 - d2h0) / 6 / (24 * N):.2f} of the forces down per rank and step ({sp.md.Builds} builds)", flush=True)
            lib.EmDeeX_tune(sp.md, b"local_io", 0)
        sp.finalize()
        if rank == 0:
            so.finalize()

    nrep = 2 if world <= 3 else 3     # M = 5*nrep cell layers; every rank needs at least three
    if 5 * nrep >= 3 * world:
        sp = build_spce(lib, True, nrep)
        so = build_spce(orc, False, nrep) if rank == 0 else None
        compare(f"spce {nrep}^3 replicas (rigid bodies, body virial)", sp, so, rank)
        # rigid-body dynamics on several ranks: body state is replicated, forces and torques of the members each rank
        # owns are all-reduced per kick (Engine::boost_all), every rank moves every body
        c = cm.load_fixture("NIST_spce_sample")
        nsteps = int(os.environ.get("EMDEE_MGPU_BODY_STEPS", "6"))
        for s in (([sp, so] if rank == 0 else [sp]) if nsteps > 0 else []):
            s.random_momenta(c["kB"] * c["Temp"], True, 4242)
            for step in range(nsteps):
                s.md.Options.Compute = (step == nsteps - 1)
                s.boost(1.0, 0.0, 0.5)
                s.displace(1.0, 0.0, 1.0)
                s.boost(1.0, 0.0, 0.5)
        Rp, Pp = (sp.download("coordinates"), sp.download("momenta")) if nsteps > 0 else (None, None)
        if rank == 0 and nsteps > 0:
            assert np.abs(Rp - so.download("coordinates")).max() < 1e-9
            assert np.abs(Pp - so.download("momenta")).max() <= 1e-8 * np.abs(Pp).max()
            assert sp.md.Builds == so.md.Builds
            for a, b, nm in [(sp.md.Kinetic.Total, so.md.Kinetic.Total, "K"), (sp.md.Kinetic.Rotational, so.md.Kinetic.Rotational, "Krot")]:
                assert abs(a - b) <= 1e-9 * abs(b), f"spce dynamics: {nm} {a!r} vs {b!r}"
        if nsteps > 0:
            compare(f"spce {nrep}^3 replicas after {nsteps} rigid-body NVE steps", sp, so, rank, ftol=1e-8, stol=1e-9)
        sp.finalize()
    elif rank == 0:
        print(f"[mgpu] spce skipped: {world} ranks would need more than {nrep}^3 replicas", flush=True)
    # ---- Ewald (coul_long + kspace_ewald) and bonded terms on the slab-decomposed path ---------------------------
    import test_oracle_bonded as t_bonded
    import test_oracle_ewald as t_ewald
    ncell = {2: 6, 3: 8, 4: 10}.get(world)     # M = floor(2L/3.1) cell layers with L = 2*ncell: at least three per rank
    if ncell is not None and os.environ.get("EMDEE_MGPU_SKIP_EWALD") != "1":
        rng = np.random.default_rng(9)
        g = 2 * ncell
        Rsalt = (np.stack(np.meshgrid(np.arange(g), np.arange(g), np.arange(g), indexing="ij"), axis=-1).reshape(-1, 3) + 0.25
                 + rng.normal(scale=0.06, size=(g ** 3, 3)))

        def salt(lib, comm):
            orig = cm.api.System.set_pair_model
            state = {"done": not comm}

            def patched(self, *a, **k):
                if not state["done"]:
                    edist.init_comm(self.lib, self)
                    state["done"] = True
                return orig(self, *a, **k)
            cm.api.System.set_pair_model = patched
            try:
                return t_ewald.rock_salt(lib, ncell=ncell, R=Rsalt, accuracy=1e-4)[0]
            finally:
                cm.api.System.set_pair_model = orig
        sp = salt(lib, True)
        so = salt(orc, False) if rank == 0 else None
        compare(f"ewald rock salt {g}^3 ions", sp, so, rank, ftol=1e-9, stol=1e-11)
        for s in ([sp, so] if rank == 0 else [sp]):
            s.random_momenta(0.05, True, 3)
            for _ in range(4):
                s.boost(1.0, 0.0, 0.01)
                s.displace(1.0, 0.0, 0.02)
                s.boost(1.0, 0.0, 0.01)
        compare("ewald rock salt after 4 steps", sp, so, rank, ftol=1e-8, stol=1e-10)
        sp.finalize()
    if world == 2:
        def water(lib, comm):
            orig = cm.api.System.set_pair_model
            state = {"done": not comm}

            def patched(self, *a, **k):
                if not state["done"]:
                    edist.init_comm(self.lib, self)
                    state["done"] = True
                return orig(self, *a, **k)
            cm.api.System.set_pair_model = patched
            try:
                return _flexible(lib)
            finally:
                cm.api.System.set_pair_model = orig

        def _flexible(lib):
            L = 18.0                      # M = floor(2L/5.8) = 6 cell layers: three per rank
            rng = np.random.default_rng(3)
            nmol = 125
            grid = np.array([[i, j, k] for i in range(5) for j in range(5) for k in range(5)], dtype=float) * 3.5 + 0.7
            mol = np.array([[0.0, 0.0, 0.0], [0.8, 0.58, 0.0], [-0.8, 0.58, 0.0]])
            R = (grid[:, None, :] + mol[None, :, :]).reshape(-1, 3) + rng.normal(scale=0.05, size=(3 * nmol, 3))
            return t_bonded.flexible_water(lib, R=R, L=L, nmol=nmol)[0]
        sp = water(lib, True)
        so = water(orc, False) if rank == 0 else None
        compare("flexible water (bonds + angles), static", sp, so, rank)
        if rank == 0:
            assert abs(sp.md.Energy.Bond - so.md.Energy.Bond) <= 1e-12 * abs(so.md.Energy.Potential)
            assert abs(sp.md.Energy.Angle - so.md.Energy.Angle) <= 1e-12 * abs(so.md.Energy.Potential)
        for s in ([sp, so] if rank == 0 else [sp]):
            s.random_momenta(0.0006, True, 13)
            for _ in range(6):
                s.boost(1.0, 0.0, 0.25)
                s.displace(1.0, 0.0, 0.5)
                s.boost(1.0, 0.0, 0.25)
        compare("flexible water after 6 steps", sp, so, rank, ftol=1e-8, stol=1e-10)
        sp.finalize()
    dist.barrier()
    if rank == 0:
        print("[mgpu] ALL OK", flush=True)
    dist.destroy_process_group()


def long_bond_scenario():
    """EMDEE_MGPU_LONG_BOND=1: a harmonic bond longer than Rc + skin on the slab-decomposed path. The partner of an owned atom
    then lies outside the rank's halo and its coordinates go stale: every rank must stop with the library's message (Error in
    bonded force computation ...) instead of computing a wrong force. The caller checks exit code and stderr."""
    rank, world, local = edist.env_rank_world()
    if os.environ.get("EMDEE_MGPU_EMULATED") == "1":
        dist.init_process_group("gloo")
        lib = cm.emulated()
    else:
        torch.cuda.set_device(local)
        os.environ["EMDEE_DEVICE"] = str(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        lib = cm.product()
    c = cm.load_fixture("NIST_spce_sample")
    nmol, L = 125, 18.0
    rng = np.random.default_rng(3)
    grid = np.array([[i, j, k] for i in range(5) for j in range(5) for k in range(5)], dtype=float) * 3.5 + 0.7
    mol = np.array([[0.0, 0.0, 0.0], [0.8, 0.58, 0.0], [-0.8, 0.58, 0.0]])
    R = (grid[:, None, :] + mol[None, :, :]).reshape(-1, 3) + rng.normal(scale=0.05, size=(3 * nmol, 3))
    types = np.tile(np.array([1, 2, 2], dtype=np.int32), nmol)
    s = lib.system(2, 1, 5.0, 0.8, 3 * nmol, types, c["mass"], None)
    edist.init_comm(lib, s)
    eps = c["epsilon"] / c["mvv2e"]
    s.set_pair_model(1, 1, lib.EmDee_shifted_force(lib.EmDee_pair_lj_cut(eps[0], c["sigma"][0])), c["kCoul"])
    s.set_pair_model(2, 2, lib.EmDee_pair_none(), c["kCoul"])
    s.set_coul_model(lib.EmDee_shifted_force(lib.EmDee_coul_cut()))
    bond = lib.EmDee_bond_harmonic(0.9, 1.0)
    for m in range(nmol):
        o = 3 * m + 1
        s.lib.EmDee_add_bond(s.md, o, o + 1, bond)
        s.lib.EmDee_add_bond(s.md, o, o + 2, bond)
    s.lib.EmDee_add_bond(s.md, 1, 3 * 2 + 1, bond)   # O of molecule 0 - O of molecule 2: 7 A along z > Rc + skin = 5.8 A
    s.upload("charges", np.tile(np.array([-0.8476, 0.4238, 0.4238]), nmol))
    s.upload("box", np.array([L]))
    s.upload("coordinates", R)          # first box + coordinates: forces are computed here -> the library stops
    s.compute_forces()
    print("[mgpu] long bond was NOT rejected", flush=True)


if __name__ == "__main__":
    if os.environ.get("EMDEE_MGPU_LONG_BOND") == "1":
        long_bond_scenario()
    else:
        main()
