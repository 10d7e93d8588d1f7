"""Generates tests/golden/*.npz from the reference's own test fixtures (run in the build container,
where /root/reference is mounted; the GPU box has no /root/reference, hence the committed copies).

  NIST_lj_sample.npz    <- reference test/NIST_lj_sample.mol  + test/lj_sample.inp
  NIST_spce_sample.npz  <- reference test/NIST_spce_sample.mol + test/spce_sample.inp

The reader follows reference test/common/mConfig.f90:33-58 (only the first box length is used) and
test/common/contained.f90:25-52 (input file: value lines alternate with bracketed title lines).

Also stores the reference's pinned known-answer triples [Potential, Virial, Potential+Kinetic]:
  test/test_pair_lj_cut.f90:46, test/test_pair_lj_sf.f90:46, test/test_pair_lj_smoothed.f90:46,
  test/test_pair_lj_shifted_smoothed.f90:46 (stale, kept for documentation).
"""
import os
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def read_inp(path):
    vals = []
    with open(path) as f:
        lines = [ln.strip() for ln in f.readlines()]
    for k in range(1, len(lines), 2):
        if lines[k]:
            vals.append(lines[k].split()[0])
    keys = ["config", "Rc", "Rs", "seed", "Dt", "Nsteps", "Nprop", "Temp", "mvv2e", "Pconv", "kB", "kCoul"]
    return dict(zip(keys, vals))


def read_mol(path):
    with open(path) as f:
        lines = f.readlines()
    ntypes = int(lines[3].split()[0])
    mass = np.zeros(ntypes)
    eps = np.zeros(ntypes)
    sig = np.zeros(ntypes)
    p = 5
    for _ in range(ntypes):
        t = lines[p].split()
        mass[int(t[0]) - 1], eps[int(t[0]) - 1], sig[int(t[0]) - 1] = float(t[1]), float(t[2]), float(t[3])
        p += 1
    L = float(lines[p + 1].split()[0])
    N = int(lines[p + 3].split()[0])
    p += 5
    mol = np.zeros(N, dtype=np.int32)
    typ = np.zeros(N, dtype=np.int32)
    q = np.zeros(N)
    R = np.zeros((N, 3))
    for i in range(N):
        t = lines[p + i].split()
        a = int(t[0]) - 1
        mol[a], typ[a], q[a] = int(t[1]), int(t[2]), float(t[3])
        R[a] = [float(t[4]), float(t[5]), float(t[6])]
    return dict(mass=mass, epsilon=eps, sigma=sig, L=L, N=N, molecule=mol, atomType=typ, Q=q, R=R)


def main():
    for mol, inp, out in [("NIST_lj_sample.mol", "lj_sample.inp", "NIST_lj_sample.npz"),
                          ("NIST_spce_sample.mol", "spce_sample.inp", "NIST_spce_sample.npz")]:
        d = read_mol(os.path.join(REF, "test", mol))
        p = read_inp(os.path.join(REF, "test", inp))
        np.savez_compressed(
            os.path.join(HERE, out), **d, Rc=float(p["Rc"]), Rs=float(p["Rs"]), seed=int(p["seed"]),
            Dt=float(p["Dt"]), Temp=float(p["Temp"]), mvv2e=float(p["mvv2e"]), kB=float(p["kB"]),
            kCoul=float(p["kCoul"]))
        print(out, d["N"], d["L"], p)
    np.savez(os.path.join(HERE, "reference_kats.npz"),
             lj_cut=np.array([-4379.8688080569782, -795.05995365595675, -3393.4159288242085]),
             lj_sf=np.array([-3898.7815412526024, 86.533120925720581, -2912.1605359154992]),
             lj_square_smoothed_skin1=np.array([-4246.3171425451010, -965.82316498190426, -3259.1347688423789]),
             lj_shifted_square_smoothed_STALE=np.array([-4180.0071205201639, -617.08138531797272, -3192.7954182277017]))


if __name__ == "__main__":
    main()
