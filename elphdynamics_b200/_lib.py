"""ctypes binding of libelph_b200.so -- the C ABI declared in include/elph_b200.h.

This is the same boundary a Julia ``ccall`` shim binds.  There is
no Python or CPU implementation of any operator behind it: if the shared library
is missing or no B200 is present, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
HEADER = PKG.parent / "include" / "elph_b200.h"
LIB_PATH = PKG / "libelph_b200.so"

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)


class ElphError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libelph_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


class Config(C.Structure):
    """``elph_config`` (include/elph_b200.h)."""
    _fields_ = [
        ("model", C.c_int32), ("index_base", C.c_int32), ("device", C.c_int32), ("reserved0", C.c_int32),
        ("Ltau", C.c_int64), ("Nsites", C.c_int64), ("Nbonds", C.c_int64), ("Nph", C.c_int64),
        ("dtau", C.c_double),
        ("neighbor_table", c_int64_p),
        ("cosht", c_double_p), ("sinht", c_double_p), ("lambda_", c_double_p), ("lambda2", c_double_p),
        ("mu", c_double_p), ("omega", c_double_p), ("omega4", c_double_p),
        ("t", c_double_p), ("alpha", c_double_p), ("alpha2", c_double_p),
        ("checkerboard_perm", c_int64_p), ("inv_checkerboard_perm", c_int64_p),
        ("phonon_to_bond", c_int64_p), ("bond_to_phonon", c_int64_p), ("primary_field", c_int64_p),
        ("cg_tol", C.c_double), ("cg_maxiter", C.c_int64), ("cg_kappa_max", C.c_double),
        ("kpm_n", C.c_int64), ("kpm_buf", C.c_double), ("kpm_c1", C.c_double), ("kpm_c2", C.c_double),
        ("fa_Q", c_double_p), ("fa_M", c_double_p),
    ]


class SolveInfo(C.Structure):
    """``elph_solve_info``: ldiv! -> (iters, residual_error, flag), src/Models.jl:74-186."""
    _fields_ = [("iters", C.c_int64), ("residual", C.c_double), ("flag", C.c_int32),
                ("used_fallback", C.c_int32), ("pcg_iters", C.c_int64)]

    def astuple(self):
        return int(self.iters), float(self.residual), int(self.flag)


class KpmInfo(C.Structure):
    """``elph_kpm_info``: outcome of setup!(P), src/KPMPreconditioners.jl:269-321."""
    _fields_ = [("active", C.c_int32), ("recomputed", C.c_int32), ("e_min", C.c_double), ("e_max", C.c_double),
                ("lambda_lo", C.c_double), ("lambda_hi", C.c_double), ("total_order", C.c_int64),
                ("max_order", C.c_int64)]


def declared_symbols() -> list[str]:
    """Every function name declared in include/elph_b200.h (for the export test)."""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(elph_[a-zA-Z0-9_]+)\s*\(", text)))


_lib = None


def load() -> C.CDLL:
    """Load the in-tree shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ElphError(-1, f"{LIB_PATH} not found: build it with `python -m elphdynamics_b200.build` "
                            "(there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    H = C.c_void_p
    i32, i64, dbl = C.c_int32, C.c_int64, C.c_double
    dp, ip = c_double_p, c_int64_p
    sig = {
        "elph_version": (C.c_char_p, []),
        "elph_last_error": (C.c_char_p, [H]),
        "elph_create": (i32, [C.POINTER(Config), C.POINTER(H)]),
        "elph_destroy": (i32, [H]),
        "elph_set_stream": (i32, [H, C.c_void_p]),
        "elph_host_register": (i32, [H, C.c_void_p, i64]),
        "elph_host_unregister": (i32, [H, C.c_void_p]),
        "elph_synchronize": (i32, [H]),
        "elph_set_solver": (i32, [H, dbl, i64, dbl]),
        "elph_kpm_configure": (i32, [H, i64, dbl, dbl, dbl]),
        "elph_set_fourier_acceleration": (i32, [H, dp, dp]),
        "elph_set_x": (i32, [H, dp]),
        "elph_get_x": (i32, [H, dp]),
        "elph_set_mu": (i32, [H, dp]),
        "elph_update_model": (i32, [H]),
        "elph_get_expnV": (i32, [H, dp]),
        "elph_get_cosh_sinh": (i32, [H, dp, dp]),
        "elph_mulM": (i32, [H, dp, dp]),
        "elph_mulMT": (i32, [H, dp, dp]),
        "elph_mulMTM": (i32, [H, dp, dp]),
        "elph_mulMTM_batch": (i32, [H, i64, dp, dp]),
        "elph_muldMdx": (i32, [H, dp, dp, dp]),
        "elph_kpm_setup": (i32, [H, dp, C.POINTER(KpmInfo)]),
        "elph_kpm_apply": (i32, [H, dp, dp]),
        "elph_kpm_get_orders": (i32, [H, ip]),
        "elph_kpm_get_coeff": (i32, [H, i64, dp]),
        "elph_cg_solve": (i32, [H, dp, dp, i32, dbl, i64, ip, dp]),
        "elph_solve": (i32, [H, dp, dp, i32, dbl, C.POINTER(SolveInfo)]),
        "elph_solve_batch": (i32, [H, i64, dp, dp, i32, dbl, C.POINTER(SolveInfo)]),
        "elph_shard_p2p_export": (i32, [H, i32, i32, C.c_void_p]),
        "elph_shard_p2p_open": (i32, [H, C.c_void_p, C.c_void_p]),
        "elph_dev_shard_cg_p2p": (i32, [H, C.c_void_p, C.c_void_p, dbl, i64, ip, dp]),
        "elph_Minv_batch": (i32, [H, i64, dp, dp, i32, C.POINTER(SolveInfo)]),
        "elph_dev_solve_batch": (i32, [H, i64, C.c_void_p, C.c_void_p, i32, dbl, C.POINTER(SolveInfo)]),
        "elph_tau_to_omega": (i32, [H, dp, dp]),
        "elph_omega_to_tau": (i32, [H, dp, dp]),
        "elph_fourier_accelerate": (i32, [H, dp, dp, dbl, i32]),
        "elph_Sb": (i32, [H, i32, dp]),
        "elph_dSbdx": (i32, [H, i32, dp]),
        "elph_calc_dSdx": (i32, [H, dp, dp, i32, dp, dp, C.POINTER(SolveInfo)]),
        "elph_langevin_step": (i32, [H, i32, dbl, dp, dp, dp, dp, dp, i32, ip, C.POINTER(SolveInfo), C.POINTER(SolveInfo)]),
        "elph_hmc_set_v": (i32, [H, dp]),
        "elph_hmc_get": (i32, [H, i32, dp]),
        "elph_hmc_refresh_v": (i32, [H, dbl, dp]),
        "elph_hmc_refresh_phi": (i32, [H, dp, dp, dp]),
        "elph_hmc_calc_Oinv": (i32, [H, i32, dp, dbl, ip, C.POINTER(i32)]),
        "elph_greens_load": (i32, [H, i64, dp, dp]),
        "elph_greens_setup": (i32, [H, i64, i64, i64, i64, i64, i64, dp, dp, dp, dp]),
        "elph_hmc_calc_H": (i32, [H, dp, dp, dp]),
        "elph_hmc_special_update": (i32, [H, i32, i64, i64, dp, dp, dp, i32, dbl, C.POINTER(i32), dp, dp, ip, C.POINTER(i32)]),
        "elph_hmc_calc_dSdx": (i32, [H, i32, dp]),
        "elph_hmc_update": (i32, [H, dbl, i64, i64, dbl, dp, dp, dp, dp, i32, dbl, C.POINTER(i32), dp, dp, dp, C.POINTER(i32)]),
        "elph_dev_mulMTM": (i32, [H, C.c_void_p, C.c_void_p]),
        "elph_dev_mulM": (i32, [H, C.c_void_p, C.c_void_p]),
        "elph_dev_mulMT": (i32, [H, C.c_void_p, C.c_void_p]),
        "elph_dev_mulMTM_replicas": (i32, [H, i64, C.c_void_p, i64, C.c_void_p, C.c_void_p, i64]),
        "elph_dev_ssh_replica_tables": (i32, [H, i64, C.c_void_p, i64, C.c_void_p, i64]),
        "elph_dev_mulMTM_replicas_ssh": (i32, [H, i64, C.c_void_p, i64, C.c_void_p, C.c_void_p, i64]),
        "elph_set_shard": (i32, [H, i64, i64]),
        "elph_dev_shard_matvec": (i32, [H, i32, C.c_void_p, C.c_void_p]),
        "elph_dev_shard_halo": (i32, [H, C.c_void_p]),
        "elph_dev_shard_matvec_halo": (i32, [H, i32, C.c_void_p, C.c_void_p]),
        "elph_shard_cg_available": (i32, [H, C.POINTER(i32)]),
        "elph_dev_shard_muldMdx": (i32, [H, C.c_void_p, C.c_void_p, C.c_void_p, dbl]),
        "elph_dev_update_model": (i32, [H]),
        "elph_dev_shard_dSbdx": (i32, [H, C.c_void_p, C.c_void_p, i32]),
        "elph_dev_fourier_accelerate_cols": (i32, [H, C.c_void_p, C.c_void_p, i64, C.c_void_p, dbl]),
        "elph_dev_tau_to_omega_cols": (i32, [H, C.c_void_p, C.c_void_p, i64]),
        "elph_dev_omega_to_tau_cols": (i32, [H, C.c_void_p, C.c_void_p, i64]),
        "elph_dev_kpm_setup_bar": (i32, [H, C.c_void_p, dp, C.POINTER(KpmInfo)]),
        "elph_kpm_set_omega_subset": (i32, [H, i64, i64]),
        "elph_dev_kpm_chains": (i32, [H, C.c_void_p, C.c_void_p]),
        "elph_kpm_shard_export": (i32, [H, i32, i32, i64, i64, C.c_void_p]),
        "elph_kpm_shard_open": (i32, [H, C.c_void_p, C.c_void_p]),
        "elph_dev_kpm_shard_apply": (i32, [H, C.c_void_p, C.c_void_p]),
        "elph_kpm_shard_check": (i32, [H]),
        "elph_dev_lincomb": (i32, [H, C.c_void_p, dbl, C.c_void_p, dbl, C.c_void_p, dbl, C.c_void_p, i64]),
        "elph_dev_dot": (i32, [H, C.c_void_p, C.c_void_p, i64, C.c_void_p]),
        "elph_dev_to_engine_layout": (i32, [H, C.c_void_p, C.c_void_p, i64]),
        "elph_dev_from_engine_layout": (i32, [H, C.c_void_p, C.c_void_p, i64]),
        "elph_dev_ptr_x": (i32, [H, C.POINTER(C.c_void_p)]),
        "elph_dev_ptr_expnV": (i32, [H, C.POINTER(C.c_void_p)]),
        "elph_dev_ptr_cosh_sinh": (i32, [H, C.POINTER(C.c_void_p)]),
        "elph_dev_cg_solve": (i32, [H, C.c_void_p, C.c_void_p, i32, dbl, i64, ip, dp]),
        "elph_dev_kpm_apply": (i32, [H, C.c_void_p, C.c_void_p]),
        "elph_dev_fourier_accelerate": (i32, [H, C.c_void_p, C.c_void_p, dbl, i32]),
        "elph_launch_count": (i64, [H]),
        "elph_set_chunk": (i32, [H, i32]),
        "elph_set_tuning": (i32, [H, i32, i32]),
        "elph_get_tuning": (i32, [H, i32, C.POINTER(i32)]),
        "elph_get_kernel_info": (i32, [H, C.POINTER(i32), C.POINTER(i32)]),
        "elph_debug_hessenberg_eigvals": (i32, [i32, dp, dp, dp]),
    }
    for name, (res, args) in sig.items():
        if not hasattr(lib, name):
            continue  # reported by the export test; calling it will raise AttributeError
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(a: np.ndarray | None, dtype=np.float64):
    """Pointer to a C-contiguous numpy array of the given dtype (None -> NULL)."""
    if a is None:
        return None
    if not (isinstance(a, np.ndarray) and a.dtype == dtype and a.flags["C_CONTIGUOUS"]):
        raise TypeError(f"expected a C-contiguous numpy array of dtype {np.dtype(dtype).name}, got "
                        f"{type(a).__name__}" + (f" of dtype {a.dtype}" if isinstance(a, np.ndarray) else ""))
    return a.ctypes.data_as(C.POINTER(C.c_double if dtype == np.float64 else C.c_int64))


def out_ptr(a: np.ndarray, n: int, name: str = "output", dtype=np.float64):
    """Pointer to an OUTPUT array the library will write ``n`` entries into: the size is checked here because the C side
    cannot see it (a short buffer would be overrun silently)."""
    if not isinstance(a, np.ndarray):
        raise TypeError(f"{name}: expected a numpy array, got {type(a).__name__}")
    if a.size != n:
        raise ValueError(f"{name}: expected {n} entries, got {a.size}")
    if not a.flags["WRITEABLE"]:
        raise ValueError(f"{name}: array is read-only")
    return ptr(a, dtype)


def check(status: int, handle=None):
    if status != 0:
        lib = load()
        msg = lib.elph_last_error(handle)
        raise ElphError(status, msg.decode() if msg else "unknown")
