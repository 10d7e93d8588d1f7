"""Edge cases of the list build and the force dispatch that the model-variety tests do not reach: list-capacity
overflow and regrowth, systems smaller than a warp, more atom types than the shared-memory table holds
(MAX_SMEM_TYPES = 16), empty cells everywhere, a cluster that leaves most of the box empty.
Each scenario runs twice: on the CPU through the kernel emulator (tests/cusim), and -- marked `gpu`, file name
sorting late -- on the device. Same bars as tests/test_gpu_parity.py.
"""
import numpy as np
import pytest

import common as cm
from test_gpu_parity import assert_state_parity


def clump(lib):
    """Very non-uniform density: a dense droplet in a big box. The first capacity guess (from the MEAN density)
    is far too small, so the build overflows and regrows (engine.cu: `if (!hflags[1]) break;`)."""
    rng = np.random.default_rng(3)
    L = 30.0
    n_side = 9
    g = (np.arange(n_side) - n_side / 2) * 1.05
    drop = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3) + 15.0 + rng.uniform(-0.05, 0.05, (n_side ** 3, 3))
    gas = rng.uniform(0, L, (40, 3))
    gas = gas[np.all(np.abs(gas - 15.0) > 6.5, axis=1) | (np.linalg.norm(gas - 15.0, axis=1) > 9.0)]
    R = np.concatenate([drop, gas])
    s = lib.system(2, 1, 2.5, 0.4, R.shape[0], None, None, None)
    s.set_pair_model(1, 1, lib.EmDee_shifted_force(lib.EmDee_pair_lj_cut(1.0, 1.0)), 0.0)
    s.upload("box", [L])
    s.upload("coordinates", R)
    return s


def tiny(n):
    def build(lib):
        rng = np.random.default_rng(n)
        L = 14.0
        g = np.arange(4) * 1.1                      # the first n sites of a jittered 4x4x4 lattice (no overlaps)
        sites = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
        R = 5.0 + sites[:n] + rng.uniform(-0.05, 0.05, (n, 3))
        s = lib.system(1, 1, 2.5, 0.3, n, None, None, None)
        s.set_pair_model(1, 1, lib.EmDee_pair_lj_cut(1.0, 1.0), 0.0)
        s.upload("box", [L])
        s.upload("coordinates", R)
        return s
    return build


def many_types(lib):
    """20 types: the interaction table no longer fits the shared-memory staging and is read from global memory."""
    R, L = cm.fcc_lj_box(6, rho=0.8, jitter=0.07, seed=13)
    N = R.shape[0]
    nt = 20
    types = (np.arange(N) % nt + 1).astype(np.int32)
    q = np.where(np.arange(N) % 2 == 0, 0.3, -0.3)
    s = lib.system(2, 1, 2.5, 0.3, N, types, np.linspace(1.0, 3.0, nt), None)
    for t in range(1, nt + 1):
        eps, sig = 0.5 + 0.05 * t, 0.9 + 0.01 * t
        model = lib.EmDee_pair_lj_cut(eps, sig) if t % 3 else lib.EmDee_pair_softcore_cut(eps, sig, 0.8)
        s.set_pair_model(t, t, model, 1.0)
    s.set_pair_model(1, 2, lib.EmDee_pair_none(), 1.0)
    s.set_coul_model(lib.EmDee_coul_damped_smoothed(0.3, 0.4))
    s.upload("charges", q)
    s.upload("box", [L])
    s.upload("coordinates", R)
    return s


SCENARIOS = {"clump": clump, "one atom": tiny(1), "two atoms": tiny(2), "five atoms": tiny(5), "31 atoms": tiny(31),
             "33 atoms": tiny(33), "20 types": many_types}


def check(product_lib, name):
    build = SCENARIOS[name]
    sp, so = build(product_lib), build(cm.oracle())
    assert_state_parity(sp, so)
    if name == "clump":
        st = sp.stats()
        assert st.build_launches >= 2        # at least one overflow + regrow
    if name not in ("one atom",):
        for s in (sp, so):
            s.random_momenta(0.5, True, 31)
        for _ in range(15):
            for s in (sp, so):
                s.boost(1.0, 0.0, 0.002)
                s.displace(1.0, 0.0, 0.004)
                s.boost(1.0, 0.0, 0.002)
        assert sp.md.Builds == so.md.Builds
        assert abs(sp.md.Energy.Potential - so.md.Energy.Potential) <= 1e-9 * max(abs(so.md.Energy.Potential), 1.0)
    sp.finalize(), so.finalize()


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_edge_case_on_emulator(name):
    check(cm.emulated(), name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SCENARIOS))
def test_edge_case_on_gpu(name):
    check(cm.product(), name)
