"""ctypes binding of the EmDee C ABI (include/emdee.h), the Python mirror of the reference's
Fortran interface module (reference src/emdee_header.f03:84-320 + generated constructors).

The function names, argument order/meaning and the by-value ``tEmDee`` struct are the reference's.
``EmDeeLib(path)`` binds ANY shared library that exports that ABI; the package itself only ever
binds the CUDA product (``emdee_b200/lib/libemdee.so``, see ``load()``). Tests bind the CPU oracle
through the same class so that parity tests drive both implementations with identical calls.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(_PKG_DIR, "lib", "libemdee.so")


class tTime(C.Structure):  # reference src/EmDeeCode.f90:44-49
    _fields_ = [("Pair", C.c_double), ("Motion", C.c_double), ("Neighbor", C.c_double), ("Total", C.c_double)]


class tEnergy(C.Structure):  # reference src/EmDeeData.f90:40-49
    _fields_ = [("Potential", C.c_double), ("Dispersion", C.c_double), ("Coulomb", C.c_double),
                ("Bond", C.c_double), ("Angle", C.c_double), ("Dihedral", C.c_double),
                ("ShadowPotential", C.c_double), ("UpToDate", C.c_bool)]


class tKinetic(C.Structure):  # reference src/EmDeeData.f90:51-59
    _fields_ = [("Total", C.c_double), ("TransPart", C.c_double * 3), ("Rotational", C.c_double),
                ("RotPart", C.c_double * 3), ("ShadowKinetic", C.c_double),
                ("ShadowRotational", C.c_double), ("UpToDate", C.c_bool)]


class tVirial(C.Structure):  # reference src/EmDeeData.f90:61-64
    _fields_ = [("Total", C.c_double), ("Body", C.c_double)]


class tOpts(C.Structure):  # reference src/EmDeeCode.f90:36-42 (the IMPLEMENTED layout)
    _fields_ = [("Translate", C.c_bool), ("Rotate", C.c_bool), ("RotationMode", C.c_int),
                ("AutoBodyUpdate", C.c_bool), ("Compute", C.c_bool)]


class tEmDee(C.Structure):  # reference src/EmDeeCode.f90:51-61
    _fields_ = [("Builds", C.c_int), ("Time", tTime), ("Energy", tEnergy), ("Kinetic", tKinetic),
                ("Virial", tVirial), ("DoF", C.c_int), ("RotDoF", C.c_int), ("Data", C.c_void_p),
                ("Options", tOpts)]


assert C.sizeof(tEmDee) == 240 and tEmDee.Data.offset == 216 and tEmDee.Options.offset == 224


class tEmDeeXStats(C.Structure):  # include/emdee_ext.h
    _fields_ = [("launches", C.c_longlong), ("force_launches", C.c_longlong), ("force_ms", C.c_double),
                ("build_launches", C.c_longlong), ("build_ms", C.c_double), ("list_entries", C.c_longlong),
                ("interacting", C.c_longlong), ("cells_per_dim", C.c_int), ("device", C.c_int)]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_mdp = C.POINTER(tEmDee)

# name -> (restype, argtypes); every symbol include/emdee.h declares
ABI = {
    "EmDee_system": (tEmDee, [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, _ip, _dp, _ip]),
    "EmDee_memory_address": (C.c_void_p, [tEmDee, C.c_char_p]),
    "EmDee_share_phase_space": (None, [tEmDee, _mdp]),
    "EmDee_layer_based_parameters": (None, [tEmDee, C.c_double, _ip, _ip]),
    "EmDee_set_pair_model": (None, [tEmDee, C.c_int, C.c_int, C.c_void_p, C.c_double]),
    "EmDee_set_pair_multimodel": (None, [tEmDee, C.c_int, C.c_int, C.POINTER(C.c_void_p), _dp]),
    "EmDee_set_kspace_model": (None, [tEmDee, C.c_void_p]),
    "EmDee_set_coul_model": (None, [tEmDee, C.c_void_p]),
    "EmDee_set_coul_multimodel": (None, [tEmDee, C.POINTER(C.c_void_p)]),
    "EmDee_ignore_pair": (None, [tEmDee, C.c_int, C.c_int]),
    "EmDee_add_bond": (None, [tEmDee, C.c_int, C.c_int, C.c_void_p]),
    "EmDee_add_angle": (None, [tEmDee, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "EmDee_add_dihedral": (None, [tEmDee, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "EmDee_download": (None, [tEmDee, C.c_char_p, C.c_void_p]),
    "EmDee_upload": (None, [_mdp, C.c_char_p, C.c_void_p]),
    "EmDee_switch_model_layer": (None, [_mdp, C.c_int]),
    "EmDee_random_momenta": (None, [_mdp, C.c_double, C.c_bool, C.c_int]),
    "EmDee_boost": (None, [_mdp, C.c_double, C.c_double, C.c_double]),
    "EmDee_displace": (None, [_mdp, C.c_double, C.c_double, C.c_double]),
    "EmDee_verlet_step": (None, [_mdp, C.c_double]),
    "EmDee_compute_forces": (None, [_mdp]),
    "EmDee_rdf": (None, [tEmDee, C.c_int, C.c_double, C.c_int, _ip, _ip, _dp]),
    "EmDee_shifted": (C.c_void_p, [C.c_void_p]),
    "EmDee_shifted_force": (C.c_void_p, [C.c_void_p]),
    "EmDee_smoothed": (C.c_void_p, [C.c_void_p, C.c_double]),
    "EmDee_shifted_smoothed": (C.c_void_p, [C.c_void_p, C.c_double]),
    "EmDee_square_smoothed": (C.c_void_p, [C.c_void_p, C.c_double]),
    "EmDee_shifted_square_smoothed": (C.c_void_p, [C.c_void_p, C.c_double]),
    "EmDee_pair_none": (C.c_void_p, []),
    "EmDee_coul_none": (C.c_void_p, []),
    "EmDee_bond_none": (C.c_void_p, []),
    "EmDee_angle_none": (C.c_void_p, []),
    "EmDee_dihedral_none": (C.c_void_p, []),
    "EmDee_pair_lj_cut": (C.c_void_p, [C.c_double, C.c_double]),
    "EmDee_pair_softcore_cut": (C.c_void_p, [C.c_double, C.c_double, C.c_double]),
    "EmDee_coul_cut": (C.c_void_p, []),
    "EmDee_coul_sf": (C.c_void_p, []),
    "EmDee_coul_damped": (C.c_void_p, [C.c_double]),
    "EmDee_coul_long": (C.c_void_p, []),
    "EmDee_coul_damped_smoothed": (C.c_void_p, [C.c_double, C.c_double]),
    "EmDee_coul_damped_square_smoothed": (C.c_void_p, [C.c_double, C.c_double]),
    "EmDee_coul_square_smoothed": (C.c_void_p, [C.c_double]),
    "EmDee_coul_shifted_square_smoothed": (C.c_void_p, [C.c_double]),
    "EmDee_bond_harmonic": (C.c_void_p, [C.c_double, C.c_double]),
    "EmDee_angle_harmonic": (C.c_void_p, [C.c_double, C.c_double]),
    "EmDee_kspace_ewald": (C.c_void_p, [C.c_double]),
}

# include/emdee_ext.h (common part)
ABI_EXT = {
    "EmDeeX_pair_count": (C.c_longlong, [tEmDee]),
    "EmDeeX_download_pairs": (C.c_longlong, [tEmDee, _ip, C.c_longlong]),
    "EmDeeX_finalize": (None, [_mdp]),
    "EmDeeX_backend": (C.c_char_p, []),
}

# include/emdee_ext.h (product only)
ABI_EXT_PRODUCT = {
    "EmDeeX_stats": (None, [tEmDee, C.POINTER(tEmDeeXStats)]),
    "EmDeeX_set_kernel_timing": (None, [tEmDee, C.c_int]),
    "EmDeeX_synchronize": (None, [tEmDee]),
    "EmDeeX_comm_mode": (C.c_int, [tEmDee]),
    "EmDeeX_io_bytes": (None, [tEmDee, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "EmDeeX_kernel_times": (None, [tEmDee, _dp, C.POINTER(C.c_longlong)]),
    "EmDeeX_tune": (None, [tEmDee, C.c_char_p, C.c_int]),
    "EmDeeX_stream": (C.c_void_p, [tEmDee]),
    "EmDeeX_measure_fp64_tflops": (C.c_double, []),
    "EmDeeX_math_probe": (None, [C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "EmDeeX_comm_unique_id": (None, [C.c_char_p]),
    "EmDeeX_comm_init": (None, [tEmDee, C.c_int, C.c_int, C.c_char_p]),
    "EmDeeX_slab_range": (None, [C.c_int, C.c_int, C.c_int, _ip, _ip]),
}


class EmDeeLib:
    """A loaded library exporting the EmDee C ABI. Attribute access gives the raw C functions."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found -- build it first (python -c 'import __graft_entry__ as g; g.build()')")
        self.path = path
        self._dll = C.CDLL(path, mode=os.RTLD_NOW | os.RTLD_LOCAL)
        for table in (ABI, ABI_EXT):
            for name, (res, args) in table.items():
                fn = getattr(self._dll, name)  # raises AttributeError if the symbol is missing
                fn.restype = res
                fn.argtypes = args
                setattr(self, name, fn)
        self.backend = self.EmDeeX_backend().decode()
        if self.backend != "oracle-cpu":
            for name, (res, args) in ABI_EXT_PRODUCT.items():
                fn = getattr(self._dll, name)
                fn.restype = res
                fn.argtypes = args
                setattr(self, name, fn)

    # ---- conveniences used by tests and bench (thin; no logic of their own) ---------------------
    def system(self, threads: int, layers: int, rc: float, skin: float, N: int,
               types: Optional[np.ndarray] = None, masses: Optional[np.ndarray] = None,
               bodies: Optional[np.ndarray] = None) -> "System":
        return System(self, threads, layers, rc, skin, N, types, masses, bodies)


def _iptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_ip)


def _dptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_dp)


class System:
    """One EmDee system; methods are 1:1 with the reference entry points (same names minus prefix)."""

    def __init__(self, lib: EmDeeLib, threads, layers, rc, skin, N, types, masses, bodies):
        self.lib = lib
        self.N = int(N)
        self._types = None if types is None else np.ascontiguousarray(types, dtype=np.int32)
        self._masses = None if masses is None else np.ascontiguousarray(masses, dtype=np.float64)
        self._bodies = None if bodies is None else np.ascontiguousarray(bodies, dtype=np.int32)
        self.md = lib.EmDee_system(int(threads), int(layers), float(rc), float(skin), self.N,
                                   _iptr(self._types), _dptr(self._masses), _iptr(self._bodies))

    # setters --------------------------------------------------------------------------------
    def set_pair_model(self, itype: int, jtype: int, model, kCoul: float = 0.0):
        self.lib.EmDee_set_pair_model(self.md, itype, jtype, model, float(kCoul))

    def set_pair_multimodel(self, itype: int, jtype: int, models: Sequence, kCoul: Sequence[float]):
        arr = (C.c_void_p * len(models))(*models)
        k = np.ascontiguousarray(kCoul, dtype=np.float64)
        self.lib.EmDee_set_pair_multimodel(self.md, itype, jtype, arr, _dptr(k))

    def set_coul_model(self, model):
        self.lib.EmDee_set_coul_model(self.md, model)

    def set_coul_multimodel(self, models: Sequence):
        arr = (C.c_void_p * len(models))(*models)
        self.lib.EmDee_set_coul_multimodel(self.md, arr)

    def set_kspace_model(self, model):
        self.lib.EmDee_set_kspace_model(self.md, model)

    def layer_based_parameters(self, InternalRc: float, Apply: Sequence[int], Bonded: Sequence[int]):
        a = np.ascontiguousarray(Apply, dtype=np.int32)
        b = np.ascontiguousarray(Bonded, dtype=np.int32)
        self.lib.EmDee_layer_based_parameters(self.md, float(InternalRc), _iptr(a), _iptr(b))

    def ignore_pair(self, i: int, j: int):
        self.lib.EmDee_ignore_pair(self.md, int(i), int(j))

    def add_bond(self, i: int, j: int, model):
        self.lib.EmDee_add_bond(self.md, int(i), int(j), model)

    def add_angle(self, i: int, j: int, k: int, model):
        self.lib.EmDee_add_angle(self.md, int(i), int(j), int(k), model)

    def add_dihedral(self, i: int, j: int, k: int, l: int, model):
        self.lib.EmDee_add_dihedral(self.md, int(i), int(j), int(k), int(l), model)

    def switch_model_layer(self, layer: int):
        self.lib.EmDee_switch_model_layer(C.byref(self.md), int(layer))

    # transfers ------------------------------------------------------------------------------
    def upload(self, option: str, array):
        a = np.ascontiguousarray(array, dtype=np.float64)
        self.lib.EmDee_upload(C.byref(self.md), option.encode(), a.ctypes.data_as(C.c_void_p))

    def download(self, option: str, shape=None) -> np.ndarray:
        if option == "box":
            out = np.zeros(1)
        else:
            out = np.zeros(shape if shape is not None else (self.N, 3))
        self.lib.EmDee_download(self.md, option.encode(), out.ctypes.data_as(C.c_void_p))
        return out[0] if option == "box" else out

    # dynamics -------------------------------------------------------------------------------
    def rdf(self, bins: int, Rc: float, itype: Sequence[int], jtype: Sequence[int]) -> np.ndarray:
        """EmDee_rdf: g[pair, bin] from the current neighbor list (reference src/EmDeeCode.f90:1281-1395)."""
        it = np.ascontiguousarray(itype, dtype=np.int32)
        jt = np.ascontiguousarray(jtype, dtype=np.int32)
        g = np.zeros((len(it), bins), dtype=np.float64)
        self.lib.EmDee_rdf(self.md, bins, float(Rc), len(it), _iptr(it), _iptr(jt), _dptr(g))
        return g

    def random_momenta(self, kT: float, adjust: bool, seed: int):
        self.lib.EmDee_random_momenta(C.byref(self.md), float(kT), bool(adjust), int(seed))

    def boost(self, lam: float, alpha: float, dt: float):
        self.lib.EmDee_boost(C.byref(self.md), float(lam), float(alpha), float(dt))

    def displace(self, lam: float, alpha: float, dt: float):
        self.lib.EmDee_displace(C.byref(self.md), float(lam), float(alpha), float(dt))

    def compute_forces(self):
        self.lib.EmDee_compute_forces(C.byref(self.md))

    def verlet_step(self, dt: float):
        self.lib.EmDee_verlet_step(C.byref(self.md), float(dt))

    def memory_address(self, option: str, shape=None) -> np.ndarray:
        """EmDee_memory_address: a numpy VIEW of the library's own array (reference src/EmDeeCode.f90:212-235)."""
        ptr = self.lib.EmDee_memory_address(self.md, option.encode())
        shape = shape if shape is not None else (self.N, 3)
        n = int(np.prod(shape))
        return np.ctypeslib.as_array(C.cast(ptr, _dp), shape=(n,)).reshape(shape)

    def share_phase_space(self, other: "System"):
        """EmDee_share_phase_space(self, other): `other` gives up its own R, P, bodies and box for `self`'s."""
        self.lib.EmDee_share_phase_space(self.md, C.byref(other.md))

    # extensions -----------------------------------------------------------------------------
    def pair_count(self) -> int:
        """Number of neighbor pairs held by this process's list (each pair once)."""
        return int(self.lib.EmDeeX_pair_count(self.md))

    def pairs(self) -> np.ndarray:
        """Neighbor pairs as a lexicographically sorted (npairs, 2) int32 array (0-based)."""
        n = int(self.lib.EmDeeX_pair_count(self.md))
        buf = np.zeros((max(n, 1), 2), dtype=np.int32)
        got = int(self.lib.EmDeeX_download_pairs(self.md, buf.ctypes.data_as(_ip), n))
        buf = buf[:got]
        order = np.lexsort((buf[:, 1], buf[:, 0]))
        return buf[order]

    def stats(self) -> tEmDeeXStats:
        s = tEmDeeXStats()
        self.lib.EmDeeX_stats(self.md, C.byref(s))
        return s

    def set_kernel_timing(self, enabled: bool):
        self.lib.EmDeeX_set_kernel_timing(self.md, int(bool(enabled)))

    def synchronize(self):
        self.lib.EmDeeX_synchronize(self.md)

    def io_bytes(self):
        """(h2d, d2h) bytes moved so far by coordinate uploads / force downloads."""
        a, b = C.c_longlong(0), C.c_longlong(0)
        self.lib.EmDeeX_io_bytes(self.md, C.byref(a), C.byref(b))
        return a.value, b.value

    KERNEL_KINDS = ("force", "build", "boost", "displace", "refresh", "exchange", "binning", "other")

    def kernel_times(self) -> dict:
        """{kind: (accumulated ms, launches)} from the library's CUDA-event ring (set_kernel_timing(True) first)."""
        ms = (C.c_double * 8)()
        n = (C.c_longlong * 8)()
        self.lib.EmDeeX_kernel_times(self.md, ms, n)
        return {k: (ms[i], n[i]) for i, k in enumerate(self.KERNEL_KINDS)}

    def stream(self) -> int:
        """cudaStream_t (as an integer) the system's kernels run on."""
        return int(self.lib.EmDeeX_stream(self.md) or 0)

    def finalize(self):
        if self.md.Data:
            self.lib.EmDeeX_finalize(C.byref(self.md))


_product: Optional[EmDeeLib] = None


def load() -> EmDeeLib:
    """Bind the CUDA product library. There is no CPU fallback: a missing library is an error."""
    global _product
    if _product is None:
        _product = EmDeeLib(PRODUCT_LIB)
        if _product.backend != "b200-cuda":
            raise RuntimeError(f"{PRODUCT_LIB} is not the CUDA product library (backend={_product.backend})")
    return _product
