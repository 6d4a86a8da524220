"""ctypes binding of libsmoqyelph_b200.so (the C ABI in include/smoqyelph_b200.h).

This is the Python twin of the `ccall` shim in julia/SmoQyElPhB200.jl: thin, no arithmetic.  The
library is built in-tree by `build()` (nvcc, sm_100a only).  Loading fails loudly when the shared
object is missing -- there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.environ.get("SMOQYELPH_B200_LIB") or os.path.join(_PKG, "libsmoqyelph_b200.so")     # (same override as the Julia shim)
HEADER = os.path.join(_ROOT, "include", "smoqyelph_b200.h")

i64, f64, vp, i32 = C.c_int64, C.c_double, C.c_void_p, C.c_int
pp = C.POINTER(C.c_void_p)

# name -> argtypes (every function returns int status unless noted)
SIGNATURES = {
    "sq_device_count": [vp],
    "sq_fdm_create": [pp, i32, i64, i64, i64, vp, vp, i64, vp, vp, f64, i64, i32],
    "sq_fdm_destroy": [vp],
    "sq_fdm_update": [vp, vp, vp, f64],
    "sq_fdm_mul": [vp, i32, vp, vp],
    "sq_fdm_cg": [vp, vp, vp, i32, vp, i32, vp, f64, i64, vp, vp],
    "sq_fdm_cg_batch": [vp, vp, vp, i64, i32, vp, i32, vp, f64, i64, vp, vp],
    "sq_fdm_get_coefficients": [vp, vp, vp, vp],
    "sq_fdm_mul_dev": [vp, i32, vp, vp],
    "sq_fdm_cg_dev": [vp, vp, vp, i32, vp, f64, i64, vp, vp],
    "sq_fdm_set_tuning": [vp, i32, i32],
    "sq_fdm_get_tuning": [vp, vp, vp, vp],
    "sq_fdm_set_fast_path": [vp, i32],
    "sq_fdm_time_mul": [vp, i32, vp, vp, i32, vp, i64, vp],
    "sq_fdm_stream": [vp, pp],
    "sq_fdm_stats": [vp, vp, i32],
    "sq_nccl_unique_id": [vp],
    "sq_fdm_init_slab": [vp, i32, i32, vp],
    "sq_fdm_mailbox_create": [vp, vp],
    "sq_fdm_mailbox_open": [vp, vp],
    "sq_fdm_set_slab_range": [vp, i64, i64],
    "sq_fdm_set_sharded_solve": [vp, i32],
    "sq_fdm_get_slab": [vp, vp, vp, vp, vp],
    "sq_kpm_create": [pp, vp, f64, i64, f64, f64],
    "sq_kpm_destroy": [vp],
    "sq_kpm_set_seed": [vp, C.c_uint64],
    "sq_kpm_update": [vp, vp, vp, vp],
    "sq_kpm_set_bounds": [vp, f64, f64],
    "sq_kpm_get_orders": [vp, vp, vp],
    "sq_kpm_get_coefs": [vp, i64, vp],
    "sq_kpm_ldiv": [vp, vp, vp],
    "sq_kpm_ldiv_dev": [vp, vp, vp],
    "sq_kpm_fourier": [vp, vp, i32],
    "sq_elph_create": [pp, vp, f64, i64, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp],
    "sq_elph_set_bare": [vp, vp, vp],
    "sq_elph_set_dispersion": [vp, i64, vp, vp, vp],
    "sq_elph_potential_derivative": [vp, vp],
    "sq_elph_destroy": [vp],
    "sq_elph_set_x": [vp, vp],
    "sq_elph_get_x": [vp, vp],
    "sq_elph_scale_x": [vp, i64, i64, f64],
    "sq_elph_swap_x": [vp, i64, i64],
    "sq_elph_backup_x": [vp],
    "sq_elph_restore_x": [vp],
    "sq_elph_shift_mu": [vp, f64],
    "sq_elph_refresh_fdm": [vp],
    "sq_elph_get_Vt": [vp, vp, vp],
    "sq_elph_bosonic_action": [vp, vp],
    "sq_pff_create": [pp, vp],
    "sq_pff_destroy": [vp],
    "sq_pff_set_seed": [vp, C.c_uint64],
    "sq_pff_set_exact_holstein": [vp, i32],
    "sq_pff_sample": [vp, vp, vp],
    "sq_pff_action": [vp, vp, vp, f64, i64, vp, vp, vp],
    "sq_pff_force": [vp, vp, vp, vp, f64, i64, vp, vp, vp],
    "sq_pff_get_fields": [vp, vp, vp, vp],
    "sq_pff_set_Phi": [vp, vp],
    "sq_pff_lambda_op": [vp, i32, vp, vp],
    "sq_pff_dM_dx": [vp, vp, f64, vp, vp],
    "sq_pff_dLambda_dx": [vp, vp, f64, vp, vp],
    "sq_hmc_create": [pp, vp, i64, f64, f64, f64, C.c_uint64],
    "sq_hmc_destroy": [vp],
    "sq_hmc_set_seed": [vp, C.c_uint64],
    "sq_hmc_update": [vp, vp, f64, f64, i64, vp, i64, vp, vp],
    "sq_hmc_init_momentum": [vp, vp, vp, vp],
    "sq_hmc_kinetic": [vp, vp, vp],
    "sq_hmc_evolve": [vp, vp, vp, f64],
    "sq_greens_create": [pp, vp, i64, C.c_uint64],
    "sq_greens_destroy": [vp],
    "sq_greens_set_seed": [vp, C.c_uint64],
    "sq_greens_update": [vp, vp, vp, f64, i64, vp],
    "sq_greens_get": [vp, vp, vp],
    "sq_greens_set_GR": [vp, vp],
    "sq_greens_measure": [vp, vp, vp, vp],
    "sq_greens_measure_GD0": [vp, i32, i32, vp, i32, i32, vp],
    "sq_greens_measure_contraction": [vp, i32, i32, i32, vp, vp, vp, vp],
    "sq_greens_measure_contraction_weighted": [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp],
    "sq_greens_measure_n_orbital": [vp, i32, i32, vp],
    "sq_greens_measure_double_occ_orbital": [vp, i32, i32, vp],
    "sq_greens_weighted_density": [vp, vp, vp],
    "sq_greens_weighted_bonds": [vp, i64, vp, vp, vp],
}
SPECIAL = {"sq_last_error": (C.c_char_p, []), "sq_version": (i32, []), "sq_fdm_launch_count": (i64, [vp]),
           "sq_hmc_last_reject": (C.c_char_p, [vp])}


def header_symbols():
    """Every function name declared in include/smoqyelph_b200.h."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sq_[a-zA-Z0-9_]+)\s*\(", txt)))


def build(verbose=False):
    """Compile the CUDA sources for sm_100a into smoqyelphqmc.jl_b200/libsmoqyelph_b200.so."""
    cmd = ["make", "-C", os.path.join(_PKG, "csrc"), "-j8"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libsmoqyelph_b200.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return LIB_PATH


class SqError(RuntimeError):
    pass


class SqNumericalInstability(SqError):
    """ABI status 3: NaN / non-finite residual, Lanczos coefficient or action.  The reference's callers turn this -- and only
    this -- into a rejected update (src/EFAPFFHMCUpdater.jl:168-187, src/reflection_update.jl:111-127)."""


_lib = None
MISSING = []


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SqError(f"{LIB_PATH} is missing: run __graft_entry__.build() (nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    global MISSING
    MISSING = [n for n in list(SIGNATURES) + list(SPECIAL) if not hasattr(L, n)]
    for name, args in SIGNATURES.items():
        if name in MISSING:     # calling it raises AttributeError; tests/test_abi.py asserts MISSING == []
            continue
        fn = getattr(L, name)
        fn.restype = i32
        fn.argtypes = args
    for name, (res, args) in SPECIAL.items():
        if name in MISSING:
            continue
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(status):
    if status == 3:
        raise SqNumericalInstability(load().sq_last_error().decode())
    if status != 0:
        raise SqError(load().sq_last_error().decode())


def ptr(a):
    """Pointer of a numpy array, a raw integer address (device pointers) or None."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return a.ctypes.data_as(C.c_void_p)


def device_count() -> int:
    n = C.c_int(0)
    check(load().sq_device_count(C.byref(n)))
    return n.value
