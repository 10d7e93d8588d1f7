#!/usr/bin/env python
"""bench.py -- atom-steps/s of EmDee's nonbonded hot path (neighbor-list maintenance + pair
forces/energy/virial) on the synthetic LJ box BASELINE.json names, through the C ABI.

  python bench.py --gpus N --steps K --warmup W            product (CUDA, emdee_b200/lib/libemdee.so)
  python bench.py --impl reference --gpus N --steps K ...  the reference ALGORITHM on host cores
                                                           (CPU oracle port: the Fortran reference cannot
                                                           be built in this image)

A "step" is one velocity-Verlet step of the reference's own test loop (reference
test/common/contained.f90:63-72): EmDee_boost, EmDee_displace, EmDee_boost -- i.e. exactly one
EmDee_compute_forces (rebuild check, list rebuild when triggered, forces + energy + virial with
Options.Compute = true) plus the two tiny momentum/position updates that produce the next configuration.

  value : atoms x steps / time with the state RESIDENT in HBM (only the scalars come back each call)
  e2e   : same metric with HOST buffers: each step uploads that step's coordinates from pinned host
          memory (EmDee_upload), runs EmDee_compute_forces and downloads the forces (EmDee_download)

One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("EMDEE_QUIET", "1")

from emdee_b200 import api  # noqa: E402

METRIC = "atom-steps/s (forces+nlist)"
UNIT = "atom-steps/s"

# workload: BASELINE.json configs[3] -- synthetic LJ fluid, rho* = 0.8442, Rc = 2.5 sigma, energy +
# virial every step; skin and dt follow SURVEY.md section 8(d) (LAMMPS in.lj convention).
RHO, RC, SKIN, DT, TSTAR = 0.8442, 2.5, 0.3, 0.005, 1.44
NCELL_DEFAULT = 63   # fcc 63^3 x 4 = 1 000 188 atoms


def make_workload(ncell, seed=86245):
    N = 4 * ncell ** 3
    L = (N / RHO) ** (1.0 / 3.0)
    a = L / ncell
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]]) + 0.25
    g = np.arange(ncell)
    cells = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    R = ((cells[:, None, :] + base[None, :, :]) * a).reshape(-1, 3)
    rng = np.random.default_rng(seed)
    R = R + rng.uniform(-0.05, 0.05, size=R.shape)
    P = np.random.default_rng(seed + 1).normal(0.0, np.sqrt(TSTAR), size=R.shape)
    P -= P.mean(axis=0)
    return np.ascontiguousarray(R), np.ascontiguousarray(P), L


def build_system(lib, R, P, L, threads, before_upload=None):
    s = lib.system(threads, 1, RC, SKIN, R.shape[0], None, None, None)
    if before_upload is not None:
        before_upload(s)          # multi-GPU: hand the system its communicator before any upload
    s.set_pair_model(1, 1, lib.EmDee_pair_lj_cut(1.0, 1.0), 0.0)
    s.upload("box", np.array([L]))
    s.upload("coordinates", R)
    s.upload("momenta", P)
    s.md.Options.Compute = True
    return s


def md_step(s):
    s.boost(1.0, 0.0, 0.5 * DT)
    s.displace(1.0, 0.0, DT)
    s.boost(1.0, 0.0, 0.5 * DT)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._thr = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for k, nme in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def oracle_lib():
    """The CPU restatement of the reference algorithm, -Ofast build (test infrastructure; used here only as
    the timed CPU baseline / reference arm)."""
    path = os.path.join(ROOT, "oracle", "_build", "libemdee_oracle_fast.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return api.EmDeeLib(path)


def cpu_run(ncell, steps, warmup, threads):
    """Times the reference algorithm (oracle port) on `threads` host cores: same workload, same step."""
    lib = oracle_lib()
    R, P, L = make_workload(ncell)
    s = build_system(lib, R, P, L, threads)
    for _ in range(warmup):
        md_step(s)
    t0 = time.perf_counter()
    for _ in range(steps):
        md_step(s)
    dt = time.perf_counter() - t0
    N = R.shape[0]
    out = {"value": N * steps / dt, "seconds": dt, "steps": steps, "N": N, "builds": int(s.md.Builds),
           "U": s.md.Energy.Potential}
    s.finalize()
    return out


def reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = int(os.environ.get("EMDEE_CPU_THREADS", cores))
    ncell = args.ncell if world == 1 else int(round(args.ncell * world ** (1.0 / 3.0)))
    N = 4 * ncell ** 3
    # bounded sample: the same box as the product arm at this N, a few steps (~0.2-1 s each per million atoms)
    steps = max(1, min(args.steps, 20 if world == 1 else max(2, 20 // world)))
    warmup = max(1, min(args.warmup, 3))
    r = cpu_run(ncell, steps, warmup, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * r["seconds"] / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(ncell, N, 1, "host cores (no GPU)"),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{steps} velocity-Verlet steps of the {N}-atom LJ box after {warmup} warm-up "
                                   f"({r['builds']} list builds); reference algorithm restated in C++/OpenMP "
                                   f"(-Ofast), the Fortran reference cannot be built in this image"},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(ncell, N, world, where):
    return {"workload": f"synthetic LJ fluid, fcc {ncell}^3 x 4 = {N} atoms in one cubic box, rho*=0.8442, Rc=2.5, skin=0.3, "
                        f"pair_lj_cut(1,1), T*=1.44, dt=0.005, energy+virial every step (BASELINE.json configs[3])",
            "atoms_per_gpu": N // world, "atoms_total": N, "rc": RC, "skin": SKIN, "dt": DT,
            "parallelism": where,
            "l2_policy": "inputs larger than L2: the neighbor list streamed every step (>300 MB at 1M atoms) exceeds the 126 MB L2"}


def spce_line_multi_gpu(args, world):
    """Informational: the resident rigid-body NVE arm of spce_line on `world` GPUs (one rank per GPU, z-slabs). Body state
    is replicated on every rank and the bodies' (F, tau) are all-reduced per kick. The box stays
    cubic, so the replica count per dimension is the nearest integer to replicas * world^(1/3) (weak scaling, approximately)."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import common as cm
    from emdee_b200 import dist as edist
    rank, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    emulated = os.environ.get("EMDEE_MGPU_EMULATED") == "1"   # CPU rehearsal of this function (tests/cusim + fake NCCL + gloo)
    if emulated:
        dist.init_process_group("gloo")
        lib, dev = cm.emulated(), "cpu"
    else:
        torch.cuda.set_device(local_rank)
        os.environ["EMDEE_DEVICE"] = str(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        lib, dev = api.load(), "cuda"
    sync = (lambda: None) if emulated else torch.cuda.synchronize
    n = int(round(args.replicas * world ** (1.0 / 3.0)))
    K, W = min(args.steps, 30), 3
    orig = cm.api.System.set_pair_model
    state = {"done": False}

    def hooked(self, *a, **k):          # the communicator must exist before the first upload
        if not state["done"]:
            edist.init_comm(self.lib, self)
            state["done"] = True
        return orig(self, *a, **k)
    cm.api.System.set_pair_model = hooked
    try:
        s, c = cm.spce_sample_system(lib, lambda l: l.EmDee_coul_damped_square_smoothed(0.2, 1.0), replicas=n, threads=1)
    finally:
        cm.api.System.set_pair_model = orig
    N = c["N"]
    s.random_momenta(c["kB"] * c["Temp"], True, 86245)

    def nve(k):
        for _ in range(k):
            s.boost(1.0, 0.0, 0.5)
            s.displace(1.0, 0.0, 1.0)
            s.boost(1.0, 0.0, 0.5)
    nve(W)
    E0 = s.md.Energy.Potential + s.md.Kinetic.Total
    b0 = s.md.Builds
    dist.barrier()
    sync()
    t0 = time.perf_counter()
    nve(K)
    sync()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        sec = float(dt.item())
        print(json.dumps({"metric": METRIC, "informational": True, "n_gpus": world,
                          "workload": f"SPC/E NIST sample x {n}^3 = {N} atoms in one cubic box, Rc=10 A, skin=2 A, rigid bodies, "
                                      "coul_damped_square_smoothed(0.2,1.0); resident NVE with the device rigid-body integrator, "
                                      "z-slab decomposition, replicated body state",
                          "value": N * K / sec, "unit": "atom-steps/s", "ms_per_step": 1e3 * sec / K, "steps": K,
                          "builds": s.md.Builds - b0,
                          "energy_drift_rel": abs(s.md.Energy.Potential + s.md.Kinetic.Total - E0) / abs(s.md.Kinetic.Total)}), flush=True)
    s.finalize()
    dist.barrier()
    dist.destroy_process_group()


def spce_line(args):
    """Informational (not the contract line): SPC/E water, NIST sample replicated n^3 times, rigid bodies,
    LJ shifted-force on O + pair_none on H + coul_damped_square_smoothed(0.2, 1.0) (reference test/test_coul_*.f90).
    Two arms: (1) host buffers: a step = upload the configuration rigidly drifted a little further,
    EmDee_compute_forces, read the scalars (e2e by nature); (2) resident: NVE with the device-resident rigid-body
    integrator (EmDee_boost / EmDee_displace / EmDee_boost, exact free-rotor rotation), nothing crosses PCIe but the
    per-step scalars. The CPU port runs the same two loops."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import common as cm
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        spce_line_multi_gpu(args, world)
        return
    lib = api.load()
    n = args.replicas
    K, W = min(args.steps, 30), 3
    res = {}
    for tag, thelib, steps in (("gpu", lib, K), ("cpu", oracle_lib(), 6)):
        s, c = cm.spce_sample_system(thelib, lambda l: l.EmDee_coul_damped_square_smoothed(0.2, 1.0), replicas=n,
                                     threads=os.cpu_count() or 1)
        N, mol = c["N"], c["molecule"]
        calls = {"k": 0}
        drift = 0.17 * np.ones(3) / np.sqrt(3.0)
        def advance():
            # the whole configuration drifts rigidly by 0.17 A per call: the physical state (energies, forces) is
            # unchanged, atoms keep crossing cells, and the displacement criterion (skin = 2 A) fires every ~6 calls
            calls["k"] += 1
            return c["R"] + calls["k"] * drift
        for _ in range(W if tag == "gpu" else 1):
            s.upload("coordinates", advance())
            s.compute_forces()
        if tag == "gpu":
            s.set_kernel_timing(True)
            st0 = s.stats()
        b0 = s.md.Builds
        t0 = time.perf_counter()
        for _ in range(steps):
            s.upload("coordinates", advance())
            s.compute_forces()
        dt = time.perf_counter() - t0
        res[tag] = {"atom_steps_per_s": N * steps / dt, "ms_per_step": 1e3 * dt / steps, "builds": s.md.Builds - b0,
                    "steps": steps, "U": s.md.Energy.Potential}
        if tag == "gpu":
            st1 = s.stats()
            fl = st1.force_launches - st0.force_launches
            res[tag]["force_kernel_ms"] = (st1.force_ms - st0.force_ms) / max(fl, 1)
            res[tag]["build_kernel_ms"] = (st1.build_ms - st0.build_ms) / max(st1.build_launches - st0.build_launches, 1)
            res[tag]["list_entries_per_atom_half"] = st1.list_entries / 2.0 / N
            res[tag]["interacting_per_atom_half"] = st1.interacting / 2.0 / N
        # ---- resident arm: rigid-body NVE on the device (dt = 1 fs, 298 K) ----
        s.upload("coordinates", c["R"])
        s.random_momenta(c["kB"] * c["Temp"], True, 86245)
        dt_fs, nres = 1.0, (steps if tag == "gpu" else 4)
        def nve(k):
            for _ in range(k):
                s.boost(1.0, 0.0, 0.5 * dt_fs)
                s.displace(1.0, 0.0, dt_fs)
                s.boost(1.0, 0.0, 0.5 * dt_fs)
        nve(W if tag == "gpu" else 1)
        E0 = s.md.Energy.Potential + s.md.Kinetic.Total
        b0 = s.md.Builds
        t0 = time.perf_counter()
        nve(nres)
        dtr = time.perf_counter() - t0
        res[tag]["resident"] = {"atom_steps_per_s": N * nres / dtr, "ms_per_step": 1e3 * dtr / nres, "steps": nres,
                                "builds": s.md.Builds - b0,
                                "energy_drift_rel": abs(s.md.Energy.Potential + s.md.Kinetic.Total - E0) / abs(s.md.Kinetic.Total)}
        s.finalize()
    print(json.dumps({"metric": METRIC, "workload": f"SPC/E NIST sample x {n}^3 = {N} atoms, Rc=10 A, skin=2 A, rigid bodies, "
                      "coul_damped_square_smoothed(0.2,1.0) as in reference test/test_coul_damped_smoothed.f90:45; arm 1: upload + compute_forces per step; arm 'resident': NVE with the "
                      "device-resident rigid-body integrator", "informational": True,
                      "gpu": res["gpu"], "cpu_port": res["cpu"], "cores": os.cpu_count()}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--ncell", type=int, default=NCELL_DEFAULT, help="fcc cells per dimension (atoms = 4*ncell^3)")
    ap.add_argument("--cpu-steps", type=int, default=4, help="steps of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="lj", choices=["lj", "spce"],
                    help="lj: the contract workload (BASELINE.json configs[3]); spce: informational line for the "
                         "SPC/E n^3 replica box (force evaluations on uploaded configurations)")
    ap.add_argument("--replicas", type=int, default=8, help="spce: replicas per dimension of the NIST sample")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    if args.workload == "spce":
        spce_line(args)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    os.environ["EMDEE_DEVICE"] = str(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lib = api.load()
    from emdee_b200 import dist as edist
    W = max(args.warmup, 3)
    K = args.steps
    # Multi-GPU (weak scaling): ONE box of about world x 1M atoms (the reference only knows cubic boxes), cut
    # into z-slabs of cell layers, one per rank; ghost positions are exchanged every step over NCCL, energies
    # all-reduced, rebuilds re-bin after an all-reduce of the owned coordinates.
    ncell = args.ncell if world == 1 else int(round(args.ncell * world ** (1.0 / 3.0)))
    R, P, L = make_workload(ncell)
    N = R.shape[0]                      # atoms of the whole job
    s = build_system(lib, R, P, L, 1, before_upload=(lambda sy: edist.init_comm(lib, sy)) if world > 1 else None)
    s.set_kernel_timing(True)

    # ---- resident arm ---------------------------------------------------------------------------
    for _ in range(W):
        md_step(s)
    st0 = s.stats()
    builds0 = s.md.Builds
    # Device timing: CUDA events on the stream the library launches its kernels on (EmDeeX_stream). Every step
    # ends with a host-visible result (EmDee_boost returns the kinetic energy), so events on torch's idle default
    # stream bracket the same region; they are recorded too and used only if the foreign stream cannot be wrapped.
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dv0, dv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        ev_stream = torch.cuda.ExternalStream(s.stream(), device=torch.device("cuda", local_rank))
    except Exception:   # pragma: no cover
        ev_stream = None

    def mark(lib_ev, def_ev):
        nonlocal ev_stream
        def_ev.record()
        if ev_stream is not None:
            try:
                lib_ev.record(ev_stream)
            except Exception:   # pragma: no cover
                ev_stream = None

    barrier()
    with ClockSampler(local_rank) as clocks:
        t0 = time.perf_counter()
        mark(ev0, dv0)
        for _ in range(K):
            md_step(s)
        mark(ev1, dv1)
        barrier()
        wall = time.perf_counter() - t0
    dev_ms, ev_where = dv0.elapsed_time(dv1), "torch default stream (steps are host-synchronous)"
    if ev_stream is not None:
        try:
            lib_ms = ev0.elapsed_time(ev1)
            if 0.5 * dev_ms <= lib_ms <= 1.5 * dev_ms:   # both bracket the same host-synchronous region
                dev_ms, ev_where = lib_ms, "library stream (EmDeeX_stream)"
        except Exception:   # pragma: no cover
            pass
    st1 = s.stats()
    builds = s.md.Builds - builds0
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    value = N * K / (dev_ms_max * 1e-3)

    force_launches = st1.force_launches - st0.force_launches
    force_ms = (st1.force_ms - st0.force_ms) / max(force_launches, 1)
    build_launches = st1.build_launches - st0.build_launches
    build_ms = (st1.build_ms - st0.build_ms) / max(build_launches, 1)
    launches = st1.launches - st0.launches
    n_local = N / world                          # atoms per rank (list statistics below are rank 0's)
    C_half = st1.list_entries / 2.0 / n_local    # neighbor-list entries per atom (half-list count, as SURVEY 8(d))
    P_half = st1.interacting / 2.0 / n_local     # entries with r < Rc per atom

    # ---- e2e arm: host buffers in the timed region -------------------------------------------------
    nframes = int(max(6, min(24, 1.0e9 // (24 * N))))   # at most ~1 GB of pinned frames per rank
    frames = torch.empty((nframes, N, 3), dtype=torch.float64).pin_memory()
    fout = torch.empty((N, 3), dtype=torch.float64).pin_memory()
    fr = frames.numpy()
    for k in range(nframes):
        md_step(s)
        fr[k] = s.download("coordinates")
    order = list(range(nframes)) + list(range(nframes - 2, 0, -1))   # ping-pong: displacements evolve like a trajectory
    e2eK = min(K, 50)
    import ctypes as C
    fptr = C.c_void_p(fout.data_ptr())

    def e2e_step(k):
        frame = frames[order[k % len(order)]]
        lib.EmDee_upload(C.byref(s.md), b"coordinates", C.c_void_p(frame.data_ptr()))
        lib.EmDee_compute_forces(C.byref(s.md))
        lib.EmDee_download(s.md, b"forces", fptr)

    for k in range(3):
        e2e_step(k)
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    for k in range(e2eK):
        e2e_step(3 + k)
    ev1.record()
    barrier()
    e2e_wall = time.perf_counter() - t0
    t = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = N * e2eK / float(t.item())

    # ---- rooflines for the dominant kernel (pair forces) -------------------------------------------
    hbm_peak, peak_src = measured_peaks()
    bytes_per_atom = 56.0 + 4.0 * C_half                     # 24 r + 24 F + 8 offsets + 4*C list
    flops_per_atom = 15.0 * C_half + 22.0 * P_half + 6.0
    ach_gbs = bytes_per_atom * n_local / (force_ms * 1e-3) / 1e9 if force_ms > 0 else None
    fp64_peak = lib.EmDeeX_measure_fp64_tflops() if rank == 0 else None
    ach_tf = flops_per_atom * n_local / (force_ms * 1e-3) / 1e12 if force_ms > 0 else None

    traffic = None
    tpath = os.path.join(ROOT, "profiles", "force_kernel_traffic.json")
    if world == 1 and ncell == NCELL_DEFAULT and os.path.exists(tpath):
        traffic = json.load(open(tpath))["dram_bytes_per_launch"]   # from the committed ncu --set full capture

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        r = cpu_run(ncell, args.cpu_steps, 1, cores)
        cpu_baseline = {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{args.cpu_steps} velocity-Verlet steps of the same {N}-atom box after 1 warm-up, "
                                  f"{r['seconds']:.1f} s; reference algorithm restated in C++/OpenMP (-Ofast)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(ncell, N, world, "single GPU" if world == 1 else
                                      f"z-slab decomposition over {world} GPUs (NCCL halo of ghost positions per step, "
                                      f"all-reduced energies, no reverse force exchange)"),
            "timing": {"device_ms_total": dev_ms_max, "events_on": ev_where, "wall_s": wall, "list_builds_in_timed_region": int(builds),
                       "force_kernel_ms": force_ms, "build_kernel_ms": build_ms,
                       "force_kernel_share_of_step": force_ms / (dev_ms_max / K) if K else None},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 24 * N, "d2h_bytes_per_step": 24 * N + 40,   # per rank (SPMD: every rank moves the full arrays)
                    "steps": e2eK, "what": "EmDee_upload(coordinates, pinned host) + EmDee_compute_forces + "
                                           "EmDee_download(forces, pinned host) per step, wall clock, max over ranks"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_pair_forces", "achieved": ach_gbs, "peak": hbm_peak,
                         "unit": "GB/s", "frac": (ach_gbs / hbm_peak) if ach_gbs else None, "traffic": traffic,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_atom": bytes_per_atom, "list_entries_per_atom_half": C_half,
                         "interacting_per_atom_half": P_half,
                         "note": "FP64-issue-bound kernel: see roofline_fp64; HBM fraction reported because "
                                 "MEASURED_PEAKS.json carries only HBM and bf16 peaks"},
            "roofline_fp64": {"bound": "fp64", "kernel": "k_pair_forces", "achieved": ach_tf, "peak": fp64_peak,
                              "unit": "TFLOP/s", "frac": (ach_tf / fp64_peak) if (ach_tf and fp64_peak) else None,
                              "algorithmic_flops_per_atom": flops_per_atom,
                              "peak_source": "DFMA microbenchmark run in this process (EmDeeX_measure_fp64_tflops)"},
            "clocks": clocks.summary(),
            "state": {"U": s.md.Energy.Potential, "W": s.md.Virial.Total, "K": s.md.Kinetic.Total,
                      "builds_total": int(s.md.Builds)},
        }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    s.finalize()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
