"""The GPU parity tests, run on the CPU through the CUDA execution-model emulator (tests/cusim).

The emulated library is the product's own abi.cpp + engine.cu + engine_*.cuh compiled for the host (threads of a
block as fibers; barriers, shuffles, ballots, atomics and shared memory emulated). Every test below is the SAME
function the `-m gpu` run executes on the B200 (tests/test_gpu_parity.py and friends), with `common.product`
pointed at the emulated library, so the kernels' logic -- indexing, loop bounds, masks, reductions, the
criterion, the opt-in paths -- is checked against the oracle without a device. What this cannot show:
performance, memory-model races, PTX-level behaviour (the few inline-PTX helpers have host stand-ins).
The product itself still refuses to run without a GPU (tests/test_abi_surface.py).
"""
import itertools
import os

import pytest

import common as cm
import test_gpu_parity as t_par
import test_zz_box_rescale as t_box
import test_zz_rdf as t_rdf
import test_zzz_bonded as t_bd
import test_zzz_config_c1 as t_c1
import test_zzz_ewald as t_ew
import test_zzzz_phase_space as t_ps
import test_zzz_rigid_bodies as t_rb
import test_zz_single_type_coulomb as t_stc


def _cases(module, skip=()):
    """(id, function, kwargs) for every test function of a GPU test module, parametrize marks expanded."""
    out = []
    for name in sorted(dir(module)):
        fn = getattr(module, name)
        if not name.startswith("test_") or not callable(fn) or name in skip:
            continue
        marks = [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]
        axes = []
        for m in marks:
            names = [n.strip() for n in m.args[0].split(",")] if isinstance(m.args[0], str) else list(m.args[0])
            vals = [v if isinstance(v, (tuple, list)) and len(names) > 1 else (v,) for v in m.args[1]]
            axes.append([dict(zip(names, v)) for v in vals])
        for combo in itertools.product(*axes) if axes else [()]:
            kw = {}
            for d in combo:
                kw.update(d)
            ident = name + ("[" + "-".join(str(getattr(v, "__name__", v)) for v in kw.values()) + "]" if kw else "")
            out.append(pytest.param(fn, kw, id=f"{module.__name__}::{ident}"))
    return out


# test_kernels_actually_ran reads device statistics whose launch counts the emulator also keeps: included.
# test_reference_kat_replay_on_gpu (100 steps x 3 models) is covered here by one model to bound the run time.
CASES = (_cases(t_par)
         + _cases(t_box, skip=("test_scenario_on_the_oracle_builds",))
         + _cases(t_stc)
         + _cases(t_rdf)
         + _cases(t_rb)
         + _cases(t_ps)
         + _cases(t_bd)
         + _cases(t_ew)
         + _cases(t_c1)
         )


@pytest.fixture(autouse=True)
def _use_emulated_library(monkeypatch):
    lib = cm.emulated()
    monkeypatch.setattr(cm, "product", lambda: lib)
    monkeypatch.setenv("EMDEE_TEST_REPLAY_STEPS", "20")
    monkeypatch.setenv("EMDEE_TEST_REPLICAS", "1")
    monkeypatch.setenv("EMDEE_TEST_C1_ATOMS", "1000")


@pytest.mark.parametrize("fn,kwargs", CASES)
def test_on_emulator(fn, kwargs, monkeypatch):
    import inspect
    if "monkeypatch" in inspect.signature(fn).parameters:
        kwargs = dict(kwargs, monkeypatch=monkeypatch)
    fn(**kwargs)


def test_scheduling_order_and_fma_contraction_do_not_matter():
    """Three re-runs of a selection of the cases above, as concurrent subprocesses:
    * CUSIM_ORDER=reverse / random:11 -- the emulator runs the threads of a block (and the blocks of a grid) in reverse
      or in a fresh random order every scheduling round. Parity must hold regardless: a failure means a missing barrier
      or a grid-wide finish that depends on which block comes last.
    * CUSIM_FMA=1 -- nvcc contracts a*b+c into fused multiply-adds by default; the regular emulator build does not
      (-ffp-contract=off), so a tolerance that only holds without FMAs would first fail on the B200. A second emulator
      build with contraction on (-O2 -march=x86-64-v3 -ffp-contract=fast) runs the parity tests of the kernels written
      after the last GPU session at their GPU tolerances."""
    import subprocess
    import sys
    order_sel = ("two_types or dynamics_with_rebuilds or kat_replay_on_gpu or spce_single_point or typed or rdf or degenerate "
                 "or verlet_step_with_shadow or next_to_rigid or rock_salt or share_phase_space")
    fma_sel = "nve_trajectory or verlet_step_with_shadow or bonded or rock_salt or testfortran or body_frames"
    cm.emulated()   # build the regular library once, before the children need it
    runs = [({"CUSIM_ORDER": "reverse"}, order_sel), ({"CUSIM_ORDER": "random:11"}, order_sel), ({"CUSIM_FMA": "1"}, fma_sel)]
    procs = []
    for extra, sel in runs:
        env = dict(os.environ, **extra)
        procs.append((extra, subprocess.Popen([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-k",
                                               f"test_on_emulator and ({sel})"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                              text=True, env=env, cwd=cm.ROOT)))
    for extra, p in procs:
        out, _ = p.communicate(timeout=1800)
        assert p.returncode == 0, f"{extra}: " + out[-3000:]

