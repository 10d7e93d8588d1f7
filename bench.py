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


def build_system(lib, R, P, L, threads, before_upload=None, workload="lj"):
    s = lib.system(threads, 1, RC, SKIN, R.shape[0], None, None, None)
    if before_upload is not None:
        before_upload(s)          # multi-GPU: hand the system its communicator before any upload
    if workload == "lj_coul_sf":
        # BASELINE.json configs[4] / SURVEY 8(d): charges +0.5/-0.5 alternating by fcc basis index (neutral), kCoul = 1,
        # coul_sf through the reference's own setter order (quirk Q1 applies to both arms alike)
        s.set_pair_model(1, 1, lib.EmDee_pair_lj_cut(1.0, 1.0), 1.0)
        s.set_coul_model(lib.EmDee_coul_sf())
        s.upload("charges", np.where(np.arange(R.shape[0]) % 2 == 0, 0.5, -0.5))
    else:
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


def cpu_run(ncell, steps, warmup, threads, workload="lj"):
    """Times the reference algorithm (oracle port) on `threads` host cores: same workload, same step."""
    lib = oracle_lib()
    R, P, L = make_workload(ncell)
    s = build_system(lib, R, P, L, threads, workload=workload)
    U0, W0 = s.md.Energy.Potential, s.md.Virial.Total   # step-0 state (the first upload evaluates the forces)
    for _ in range(warmup):
        md_step(s)
    t0 = time.perf_counter()
    for _ in range(steps):
        md_step(s)
    dt = time.perf_counter() - t0
    N = R.shape[0]
    out = {"value": N * steps / dt, "seconds": dt, "steps": steps, "N": N, "builds": int(s.md.Builds),
           "U": s.md.Energy.Potential, "W": s.md.Virial.Total, "K": s.md.Kinetic.Total, "U0": U0, "W0": W0}
    s.finalize()
    return out


def total_ncell(args, world):
    """fcc cells per dimension of the whole job: weak scaling grows the box with the rank count (about `--ncell`^3 x 4
    atoms per GPU, cubic box), strong scaling keeps the `--ncell` box whatever the rank count."""
    base = args.ncell
    if args.atoms_per_gpu:
        base = max(8, int(round((args.atoms_per_gpu / 4.0) ** (1.0 / 3.0))))
    if world == 1 or args.scaling == "strong":
        return base
    return int(round(base * world ** (1.0 / 3.0)))


def parallelism(world):
    return "single GPU" if world == 1 else (f"z-slab decomposition over {world} GPUs: per step halo positions + rebuild-criterion "
                                            f"state to the neighbors / peers and small all-reduces of energies and kinetic sums "
                                            f"(the library's own kernels over NVLink peer memory, else NCCL), no reverse force exchange")


def workload_config(args, ncell, N, world):
    what = {"lj": "pair_lj_cut(1,1) (BASELINE.json configs[3])",
            "lj_coul_sf": "pair_lj_cut(1,1) + coul_sf, charges +-0.5 by basis index, kCoul=1 (BASELINE.json configs[4])"}[args.workload]
    return {"workload": f"synthetic fluid, fcc {ncell}^3 x 4 = {N} atoms in one cubic box, rho*=0.8442, Rc=2.5, skin=0.3, "
                        f"{what}, T*=1.44, dt=0.005, energy+virial every step",
            "atoms_per_gpu": N // world, "atoms_total": N, "rc": RC, "skin": SKIN, "dt": DT,
            "parallelism": parallelism(world),
            "l2_policy": "inputs larger than L2: the neighbor list streamed every step (>300 MB per million atoms) exceeds the 126 MB L2"}


def reference_arm(args, rank, world):
    """The reference ALGORITHM on the box's host cores (C++/OpenMP restatement, oracle/): same box, same step as the product
    arm at this rank count; each run a bounded sample (the step count below) so that it ends within minutes."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = int(os.environ.get("EMDEE_CPU_THREADS", cores))
    ncell = total_ncell(args, world)
    N = 4 * ncell ** 3
    # bounded sample: about 0.2-0.4 s per step and million atoms on 16-32 cores
    budget_steps = max(2, int(20e6 // N))
    steps = max(1, min(args.steps, budget_steps))
    warmup = max(1, min(args.warmup, 3))
    r = cpu_run(ncell, steps, warmup, threads, args.workload)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * r["seconds"] / steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, ncell, N, world),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{steps} velocity-Verlet steps of the {N}-atom box after {warmup} warm-up "
                                   f"({r['builds']} list builds) on host cores, no GPU; reference algorithm restated in "
                                   f"C++/OpenMP (-Ofast), the Fortran reference cannot be built in this image"},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "state": {"U": r["U"], "W": r["W"], "K": r["K"], "U_step0": r["U0"], "W_step0": r["W0"], "builds_total": r["builds"]},
    }
    print(json.dumps(line), flush=True)


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


def spce_measure(args, cpu=True):
    """SPC/E water, NIST sample replicated n^3 times (n = 8: 1 152 000 atoms), rigid bodies, LJ shifted-force on O +
    pair_none on H + coul_damped_square_smoothed(0.2, 1.0) (reference test/test_coul_damped_smoothed.f90:45).
    Two arms: (1) host buffers: a step = upload the configuration rigidly drifted a little further,
    EmDee_compute_forces, read the scalars (e2e by nature); (2) resident: NVE with the device-resident rigid-body
    integrator (EmDee_boost / EmDee_displace / EmDee_boost, exact free-rotor rotation), nothing crosses PCIe but the
    per-step scalars. The CPU port runs the same two loops on a bounded sample. One GPU."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import common as cm
    lib = api.load()
    n = args.replicas
    K, W = min(args.steps, 30), 3
    res = {}
    arms = [("gpu", lib, K)] + ([("cpu", oracle_lib(), 4)] if cpu else [])
    for tag, thelib, steps in arms:
        s, c = cm.spce_sample_system(thelib, lambda l: l.EmDee_coul_damped_square_smoothed(0.2, 1.0), replicas=n,
                                     threads=os.cpu_count() or 1)
        N = c["N"]
        calls = {"k": 0}
        drift = 0.17 * np.ones(3) / np.sqrt(3.0)
        def advance():
            # the whole configuration drifts rigidly by 0.17 A per call: the physical state (energies, forces) is
            # unchanged, atoms keep crossing cells, and the displacement criterion (skin = 2 A) fires every ~6 calls
            calls["k"] += 1
            return c["R"] + calls["k"] * drift
        # ---- resident arm: rigid-body NVE on the device (dt = 1 fs, 298 K) ----
        # (run first: it also brings the device back to its working clocks after the CPU baseline of the LJ line, during which
        # the GPU sat idle -- the host-buffer loop below, with the device waiting on the host between launches, does not)
        res[tag] = {}
        if tag == "gpu":
            s.set_kernel_timing(True)
        s.upload("coordinates", c["R"])
        s.random_momenta(c["kB"] * c["Temp"], True, 86245)
        dt_fs, nres = 1.0, (steps if tag == "gpu" else 3)
        def nve(k):
            for _ in range(k):
                s.boost(1.0, 0.0, 0.5 * dt_fs)
                s.displace(1.0, 0.0, dt_fs)
                s.boost(1.0, 0.0, 0.5 * dt_fs)
        nve(10 if tag == "gpu" else 1)
        E0 = s.md.Energy.Potential + s.md.Kinetic.Total
        b0 = s.md.Builds
        if tag == "gpu":
            s.synchronize()
            st0 = s.stats()
        t0 = time.perf_counter()
        nve(nres)
        dtr = time.perf_counter() - t0
        if tag == "gpu":
            # the pair kernel's time is taken here, where the GPU is busy back to back: in the host-buffer loop below the
            # device waits on the host between launches and the event times wander with its clocks (5.6 to 35 ms were seen)
            s.synchronize()
            st1 = s.stats()
            fl = st1.force_launches - st0.force_launches
            res[tag]["force_kernel_ms"] = (st1.force_ms - st0.force_ms) / max(fl, 1)
        res[tag]["resident"] = {"atom_steps_per_s": N * nres / dtr, "ms_per_step": 1e3 * dtr / nres, "steps": nres,
                                "builds": s.md.Builds - b0,
                                "energy_drift_rel": abs(s.md.Energy.Potential + s.md.Kinetic.Total - E0) / abs(s.md.Kinetic.Total)}
        next_frame = advance
        if tag == "gpu":
            # the GPU arm uploads from pinned host frames prepared beforehand, like the LJ e2e arm: the timed step is
            # EmDee_upload + EmDee_compute_forces, not the numpy arithmetic that makes the synthetic frame
            import torch
            frames = torch.empty((W + steps, N, 3), dtype=torch.float64).pin_memory()
            fr = frames.numpy()
            for k in range(W + steps):
                fr[k] = advance()
            it = iter(range(W + steps))
            next_frame = lambda: fr[next(it)]   # noqa: E731
        for _ in range(W if tag == "gpu" else 1):
            s.upload("coordinates", next_frame())
            s.compute_forces()
        if tag == "gpu":
            st0 = s.stats()
        b0 = s.md.Builds
        t0 = time.perf_counter()
        for _ in range(steps):
            s.upload("coordinates", next_frame())
            s.compute_forces()
        dt = time.perf_counter() - t0
        res[tag].update({"atom_steps_per_s": N * steps / dt, "ms_per_step": 1e3 * dt / steps, "builds": s.md.Builds - b0,
                         "steps": steps, "U": s.md.Energy.Potential, "N": N})
        if tag == "gpu":
            st1 = s.stats()
            fl = st1.force_launches - st0.force_launches
            res[tag]["force_kernel_ms_host_loop"] = (st1.force_ms - st0.force_ms) / max(fl, 1)
            res[tag]["build_kernel_ms"] = (st1.build_ms - st0.build_ms) / max(st1.build_launches - st0.build_launches, 1)
            res[tag]["list_entries_per_atom_half"] = st1.list_entries / 2.0 / N
            res[tag]["interacting_per_atom_half"] = st1.interacting / 2.0 / N
        s.finalize()
    res["workload"] = (f"SPC/E NIST sample x {n}^3 = {res['gpu']['N']} atoms, Rc=10 A, skin=2 A, rigid bodies, "
                       "coul_damped_square_smoothed(0.2,1.0) as in reference test/test_coul_damped_smoothed.f90:45")
    return res


def spce_block(args, cpu=True):
    """The SPC/E half of the headline metric ("LJ & SPC/E"), shaped like the contract line: value = resident rigid-body NVE,
    e2e = the host-buffer arm, roofline of the pair kernel, CPU port beside it."""
    r = spce_measure(args, cpu)
    g = r["gpu"]
    N = g["N"]
    C_half, P_half = g["list_entries_per_atom_half"], g["interacting_per_atom_half"]
    # 15 per listed pair (separation + cutoff test); per interacting pair 1 sqrt + ~24 for the damped, smoothed
    # Coulomb term (1 exp, 2 div) on the charged pairs (all of them in water) and 22 + 4 for the shifted-force LJ term on the
    # O-O pairs (1/9 of the pairs)
    flops_per_atom = 15.0 * C_half + P_half * (25.0 + 26.0 / 9.0) + 6.0
    fm = g["force_kernel_ms"]
    peak = api.load().EmDeeX_measure_fp64_tflops()
    ach = flops_per_atom * N / (fm * 1e-3) / 1e12 if fm > 0 else None
    out = {"workload": r["workload"], "value": g["resident"]["atom_steps_per_s"], "unit": UNIT,
           "ms_per_step": g["resident"]["ms_per_step"], "steps": g["resident"]["steps"], "builds": g["resident"]["builds"],
           "energy_drift_rel": g["resident"]["energy_drift_rel"],
           "what": "value: resident NVE with the device rigid-body integrator (EmDee_boost / EmDee_displace / EmDee_boost); "
                   "e2e: EmDee_upload(coordinates, pinned host frames) + EmDee_compute_forces per step",
           "e2e": {"value": g["atom_steps_per_s"], "unit": UNIT, "ms_per_step": g["ms_per_step"], "h2d_bytes_per_step": 24 * N,
                   "d2h_bytes_per_step": 40, "steps": g["steps"], "builds": g["builds"]},
           "timing": {"force_kernel_ms": fm, "build_kernel_ms": g["build_kernel_ms"],
                      "force_kernel_ms_host_loop": g.get("force_kernel_ms_host_loop")},
           "roofline": {"bound": "fp64", "kernel": "k_pair_forces_typed", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                        "frac": (ach / peak) if (ach and peak) else None, "algorithmic_flops_per_atom": flops_per_atom,
                        "list_entries_per_atom_half": C_half, "interacting_per_atom_half": P_half,
                        "peak_source": "DFMA microbenchmark run in this process"},
           "state": {"U": g["U"]}}
    if "cpu" in r:
        c = r["cpu"]
        out["cpu_baseline"] = {"value": c["resident"]["atom_steps_per_s"], "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                               "sample": f"{c['resident']['steps']} rigid-body NVE steps of the same box (resident arm); host-buffer arm "
                                         f"{c['atom_steps_per_s']:.4g} atom-steps/s over {c['steps']} evaluations; U = {c['U']!r}",
                               "e2e_value": c["atom_steps_per_s"]}
    return out


def spce_line(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        spce_line_multi_gpu(args, world)
        return
    r = spce_measure(args)
    print(json.dumps({"metric": METRIC, "workload": r["workload"] + "; arm 1: upload + compute_forces per step; arm 'resident': NVE "
                      "with the device-resident rigid-body integrator", "informational": True,
                      "gpu": r["gpu"], "cpu_port": r.get("cpu"), "cores": os.cpu_count()}), flush=True)


def parity_block(lib, s, args, world, rank, dist, ncell, R, P, L, before_upload):
    """Step-0 parity of the distributed (or single-GPU) product against the CPU oracle on rank 0, outside the timed region:
    potential energy, virial and the number of neighbor pairs (pair SETS are compared bit for bit in tests/). Boxes the
    oracle cannot hold (more than ~10M atoms) are checked on an 8M-atom box cut into the same number of slabs."""
    import torch
    sub = None
    if R.shape[0] > 10_000_000:
        sub = 126
        Rc_, Pc_, Lc_ = make_workload(sub)
        sp = build_system(lib, Rc_, Pc_, Lc_, 1, before_upload=before_upload, workload=args.workload)
    else:
        sp, Rc_, Pc_, Lc_ = s, R, P, L
    U, W = sp.md.Energy.Potential, sp.md.Virial.Total
    npairs = torch.tensor([sp.pair_count()], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(npairs)
    out = None
    if rank == 0:
        so = build_system(oracle_lib(), Rc_, Pc_, Lc_, os.cpu_count() or 1, workload=args.workload)
        Uo, Wo, no = so.md.Energy.Potential, so.md.Virial.Total, so.pair_count()
        so.finalize()
        out = {"U_rel": abs(U - Uo) / abs(Uo), "W_rel": abs(W - Wo) / abs(Wo), "pairs_equal": int(npairs.item()) == int(no),
               "pairs": int(npairs.item()), "U": U, "U_oracle": Uo, "atoms": int(Rc_.shape[0]),
               "what": "step-0 potential energy, virial and neighbor-pair count vs the CPU oracle (strict reference arithmetic) "
                       "on rank 0" + (f"; {R.shape[0]}-atom box exceeds the oracle: checked on fcc {sub}^3 x 4 over the same ranks" if sub else "")}
    if sub is not None:
        sp.finalize()
    if world > 1:
        dist.barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--ncell", type=int, default=NCELL_DEFAULT, help="fcc cells per dimension per GPU (atoms = 4*ncell^3)")
    ap.add_argument("--atoms-per-gpu", type=int, default=0, help="alternative to --ncell: atoms per GPU (rounded to an fcc box)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the box grows with the rank count; strong: the --ncell box is cut into more slabs")
    ap.add_argument("--cpu-steps", type=int, default=4, help="steps of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-spce", action="store_true", help="skip the SPC/E block of the line")
    ap.add_argument("--no-parity", action="store_true", help="skip the step-0 oracle check")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer arm (development runs)")
    ap.add_argument("--workload", default="lj", choices=["lj", "lj_coul_sf", "spce"],
                    help="lj: the contract workload (BASELINE.json configs[3]); lj_coul_sf: configs[4] (LJ + coul_sf, charged); "
                         "spce: the SPC/E n^3 replica box alone (informational line)")
    ap.add_argument("--replicas", type=int, default=8, help="spce: replicas per dimension of the NIST sample")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if args.workload == "spce":
            args.workload = "lj"
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
    # Multi-GPU: ONE cubic box (the reference only knows cubic boxes) cut into z-slabs of cell layers, one per rank
    # (DESIGN.md section 7). Weak scaling: about 4*ncell^3 atoms per rank; strong scaling: the same box for every rank count.
    ncell = total_ncell(args, world)
    R, P, L = make_workload(ncell)
    N = R.shape[0]                      # atoms of the whole job
    before_upload = (lambda sy: edist.init_comm(lib, sy)) if world > 1 else None
    s = build_system(lib, R, P, L, 1, before_upload=before_upload, workload=args.workload)
    torch.cuda.reset_peak_memory_stats()
    parity = None
    if not args.no_parity:
        parity = parity_block(lib, s, args, world, rank, dist, ncell, R, P, L, before_upload)
    s.set_kernel_timing(True)

    # ---- resident arm ---------------------------------------------------------------------------
    for _ in range(W):
        md_step(s)
    st0 = s.stats()
    kt0 = s.kernel_times()
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
        trace = [] if os.environ.get("EMDEE_BENCH_TRACE") else None   # development: host wall time of every step
        for _ in range(K):
            if trace is not None:
                tb, b_before = time.perf_counter(), s.md.Builds
            md_step(s)
            if trace is not None:
                trace.append((time.perf_counter() - tb, int(s.md.Builds - b_before)))
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
    kt1 = s.kernel_times()
    builds = s.md.Builds - builds0
    if trace is not None:
        ts = sorted(((w, k, b) for k, (w, b) in enumerate(trace)), reverse=True)[:12]
        plain = sorted(w for w, b in trace if b == 0)
        rb = sorted(w for w, b in trace if b != 0)
        sys.stderr.write(f"[bench trace r{rank}] median step without rebuild {1e3 * plain[len(plain) // 2]:.3f} ms, with rebuild "
                         f"{1e3 * rb[len(rb) // 2] if rb else 0:.3f} ms ({len(rb)} of {len(trace)}); slowest: " +
                         ", ".join(f"#{k}:{1e3 * w:.2f}ms{'R' if b else ''}" for w, k, b in ts) + "\n")
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
    # in-situ device time per kernel kind and step (CUDA events from the library's ring; warm caches, unlike ncu)
    per_step = {k: (kt1[k][0] - kt0[k][0]) / K for k in kt1 if kt1[k][1] > kt0[k][1]}
    comm_mode = int(lib.EmDeeX_comm_mode(s.md))
    state_res = {"U": s.md.Energy.Potential, "W": s.md.Virial.Total, "K": s.md.Kinetic.Total, "builds_total": int(s.md.Builds)}
    peak_mem = torch.cuda.max_memory_allocated()   # torch's own share; the library's buffers are reported by the driver query
    free_b, total_b = torch.cuda.mem_get_info()
    mem_used = total_b - free_b

    # ---- e2e arm: host buffers in the timed region -------------------------------------------------
    e2e = None
    if not args.no_e2e:
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
        if world > 1:
            # several GPUs: every rank moves only its own slab across PCIe (owned + halo coordinates up, owned forces down,
            # zero-copy from / to its pinned arrays); assembling a full force array, if wanted, is the caller's gather
            lib.EmDeeX_tune(s.md, b"local_io", 1)

        def e2e_step(k):
            frame = frames[order[k % len(order)]]
            lib.EmDee_upload(C.byref(s.md), b"coordinates", C.c_void_p(frame.data_ptr()))
            lib.EmDee_compute_forces(C.byref(s.md))
            lib.EmDee_download(s.md, b"forces", fptr)

        for k in range(3):
            e2e_step(k)
        io0 = s.io_bytes()
        barrier()
        t0 = time.perf_counter()
        for k in range(e2eK):
            e2e_step(3 + k)
        barrier()
        e2e_wall = time.perf_counter() - t0
        io1 = s.io_bytes()
        t = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": N * e2eK / float(t.item()), "unit": UNIT,
               "h2d_bytes_per_step": (io1[0] - io0[0]) // e2eK, "d2h_bytes_per_step": (io1[1] - io0[1]) // e2eK + 40,
               "steps": e2eK, "ms_per_step": 1e3 * float(t.item()) / e2eK,
               "what": "EmDee_upload(coordinates, pinned host) + EmDee_compute_forces + EmDee_download(forces, pinned host) per step, "
                       "wall clock, max over ranks; bytes are per rank, counted by the library" +
                       ("; several GPUs: each rank moves its own slab only (EmDeeX_tune local_io)" if world > 1 else "")}

    # ---- rooflines for the dominant kernel (pair forces) -------------------------------------------
    hbm_peak, peak_src = measured_peaks()
    coul = args.workload == "lj_coul_sf"
    bytes_per_atom = 56.0 + 4.0 * C_half + (12.0 if coul else 0.0)   # 24 r + 24 F + 8 offsets + 4*C list (+ charge, type)
    flops_per_atom = 15.0 * C_half + (22.0 + (13.0 if coul else 0.0)) * P_half + 6.0     # SURVEY 8(d); coul_sf: +1 sqrt +11 +1 div
    ach_gbs = bytes_per_atom * n_local / (force_ms * 1e-3) / 1e9 if force_ms > 0 else None
    fp64_peak = lib.EmDeeX_measure_fp64_tflops() if rank == 0 else None
    ach_tf = flops_per_atom * n_local / (force_ms * 1e-3) / 1e12 if force_ms > 0 else None

    traffic = None
    tpath = os.path.join(ROOT, "profiles", "force_kernel_traffic.json")
    if world == 1 and ncell == NCELL_DEFAULT and args.workload == "lj" and os.path.exists(tpath):
        traffic = json.load(open(tpath))["dram_bytes_per_launch"]   # from the committed ncu --set full capture

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        r = cpu_run(ncell, args.cpu_steps, 1, cores, args.workload)
        cpu_baseline = {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{args.cpu_steps} velocity-Verlet steps of the same {N}-atom box after 1 warm-up, "
                                  f"{r['seconds']:.1f} s; reference algorithm restated in C++/OpenMP (-Ofast)"}
    spce = None
    if rank == 0 and world == 1 and not args.no_spce and args.workload == "lj":
        s.finalize()
        s = None
        spce = spce_block(args, cpu=not args.no_cpu_baseline)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, ncell, N, world),
            "timing": {"device_ms_total": dev_ms_max, "events_on": ev_where, "wall_s": wall, "list_builds_in_timed_region": int(builds),
                       "force_kernel_ms": force_ms, "build_kernel_ms": build_ms,
                       "force_kernel_share_of_step": force_ms / (dev_ms_max / K) if K else None,
                       "kernel_ms_per_step": per_step, "device_memory_used_bytes": int(mem_used),
                       "comm_mode": {0: "single GPU", 1: "NCCL per step", 2: "NVLink peer kernels per step, NCCL at rebuilds"}[comm_mode]},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_pair_forces", "achieved": ach_gbs, "peak": hbm_peak,
                         "unit": "GB/s", "frac": (ach_gbs / hbm_peak) if ach_gbs else None, "traffic": traffic,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_atom": bytes_per_atom, "list_entries_per_atom_half": C_half,
                         "interacting_per_atom_half": P_half,
                         "note": "the kernel is bound by the L1 gather path and FP64 issue, not by HBM: see roofline_fp64; HBM "
                                 "fraction reported because MEASURED_PEAKS.json carries only HBM and bf16 peaks"},
            "roofline_fp64": {"bound": "fp64", "kernel": "k_pair_forces", "achieved": ach_tf, "peak": fp64_peak,
                              "unit": "TFLOP/s", "frac": (ach_tf / fp64_peak) if (ach_tf and fp64_peak) else None,
                              "algorithmic_flops_per_atom": flops_per_atom,
                              "peak_source": "DFMA microbenchmark run in this process (EmDeeX_measure_fp64_tflops)"},
            "clocks": clocks.summary(),
            "state": state_res,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if parity is not None:
            line["parity"] = parity
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        if spce is not None:
            line["spce"] = spce
        print(json.dumps(line), flush=True)
    if s is not None:
        s.finalize()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
