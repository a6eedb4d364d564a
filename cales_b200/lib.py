"""ctypes binding of libcales_b200.so (the C ABI declared in include/cales_b200.h).

There is no fallback: if the shared library is missing or a call fails, an exception is raised."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# Two arithmetic variants of the same sources (cales_b200/csrc/Makefile):
#   "fma"    libcales_b200.so         the product: fp64 contraction on (what the reference's own GPU build does)
#   "strict" libcales_b200_strict.so  -fmad=false, bit-identical to a non-contracting CPU build (test build)
# CALES_B200_ARITH=strict|fma picks the default of load(); both can live in one process (linked -Bsymbolic, RTLD_LOCAL).
LIB_PATHS = {"fma": os.path.join(_HERE, "libcales_b200.so"), "strict": os.path.join(_HERE, "libcales_b200_strict.so")}
DEFAULT_ARITH = os.environ.get("CALES_B200_ARITH", "fma")
LIB_PATH = LIB_PATHS["fma"]

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)
vp = C.c_void_p


class Bound(C.Structure):
    """cales_bound: three device planes (src/typedef.f90:10-14)."""
    _fields_ = [("x", vp), ("y", vp), ("z", vp)]


class StepArgs(C.Structure):
    """cales_step_args (include/cales_b200.h): what the callees of one RK3 substep take, gathered once."""
    _fields_ = ([(nm, C.c_int * 3) for nm in ("n", "ng", "lo", "hi")] + [(nm, C.c_int * 6) for nm in ("nb", "is_bound", "lwm", "index_wm")] +
                [("is_forced", C.c_int * 3), ("plan", C.c_int)] +
                [(nm, C.c_double * 3) for nm in ("dl", "dli", "l", "velf", "bforce")] +
                [("visc", C.c_double), ("hwm", C.c_double), ("normfft", C.c_double),
                 ("cbcvel", C.c_char * 18), ("cbcpre", C.c_char * 6), ("cbcsgs", C.c_char * 6), ("sgstype", C.c_char * 8)] +
                [(nm, vp) for nm in ("zc", "zf", "dzc", "dzf", "dzci", "dzfi", "grid_vol_ratio_c", "grid_vol_ratio_f",
                                     "lambdaxy", "a", "b", "c", "rhsbx", "rhsby", "rhsbz")] +
                [(nm, Bound) for nm in ("bcu", "bcv", "bcw", "bcp", "bcs", "bcu_mag", "bcv_mag", "bcw_mag", "bcuf", "bcvf", "bcwf")] +
                [(nm, vp) for nm in ("u", "v", "w", "p", "pp", "visct")])


class CalesError(RuntimeError):
    pass


_libs = {}


def _ia(a):
    return np.ascontiguousarray(a, dtype=np.int32).ctypes.data_as(c_int_p)


def _da(a):
    return np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(c_dbl_p)


def _ca(a):
    """(0:1,3[,3]) character table in Fortran order -> bytes."""
    a = np.asarray(a)
    return b"".join(x.encode() for x in a.ravel(order="F"))


def _tab(a):
    """(0:1,3) integer/logical table in Fortran order."""
    return np.ascontiguousarray(np.asarray(a).astype(np.int32).ravel(order="F"))


SIGNATURES = {
    "cales_version": (C.c_char_p, []),
    "cales_last_error": (C.c_char_p, [vp]),
    "cales_get_unique_id": (C.c_int, [C.c_char_p]),
    "cales_init": (C.c_int, [C.POINTER(vp), c_int_p, c_int_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int, vp, C.c_int]),
    "cales_finalize": (C.c_int, [vp]),
    "cales_get_decomp": (C.c_int, [vp] + [c_int_p] * 10),
    "cales_distribute": (C.c_int, [C.c_int, C.c_int, c_int_p, c_int_p, c_int_p]),
    "cales_pencil": (C.c_int, [c_int_p, c_int_p, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p]),
    "cales_neighbours": (C.c_int, [c_int_p, C.c_int, C.c_char_p, C.c_int, c_int_p, c_int_p]),
    "cales_transpose_plan": (C.c_int, [c_int_p, c_int_p, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p]),
    "cales_stream_synchronize": (C.c_int, [vp]),
    "cales_launch_count": (C.c_long, [vp]),
    "cales_initsolver": (C.c_int, [vp, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p, C.c_char_p,
                                   C.c_char_p, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, c_int_p, c_dbl_p]),
    "cales_fftend": (C.c_int, [vp, C.c_int]),
    "cales_solver": (C.c_int, [vp, c_int_p, c_int_p, C.c_int, C.c_double, vp, vp, vp, vp, C.c_char_p, C.c_char_p, vp]),
    "cales_solver_exchange": (C.c_char_p, [vp]),
    "cales_solver_gaussel_z": (C.c_int, [vp, c_int_p, vp, vp, vp, C.c_char_p, C.c_char_p, vp]),
    "cales_rk": (C.c_int, [vp, c_dbl_p, c_int_p, c_dbl_p, vp, vp, vp, vp, C.c_double, C.c_double, vp, c_int_p, c_dbl_p, c_dbl_p,
                           vp, vp, vp, vp, c_dbl_p]),
    "cales_rk_fused": (C.c_int, [vp, c_dbl_p, c_int_p, c_dbl_p, vp, vp, vp, vp, C.c_double, C.c_double, vp, c_int_p, c_dbl_p, c_dbl_p,
                                 vp, vp, vp, vp, vp, vp, vp]),
    "cales_mom_xyz_ad": (C.c_int, [vp, c_int_p, C.c_double, C.c_double, vp, vp, C.c_double, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "cales_bulk_forcing": (C.c_int, [vp, c_int_p, c_int_p, c_dbl_p, vp, vp, vp]),
    "cales_bulk_mean": (C.c_int, [vp, c_int_p, vp, vp, c_dbl_p]),
    "cales_bounduvw": (C.c_int, [vp, C.c_char_p, c_int_p] + [C.POINTER(Bound)] * 6 + [c_int_p, c_int_p, c_int_p, c_dbl_p, c_dbl_p,
                                 vp, vp, vp, vp, C.c_double, C.c_double, c_int_p, C.c_int, C.c_int, vp, vp, vp]),
    "cales_boundp": (C.c_int, [vp, C.c_char_p, c_int_p, C.POINTER(Bound), c_int_p, c_int_p, c_dbl_p, vp, vp]),
    "cales_cmpt_rhs_b": (C.c_int, [vp, c_int_p, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p, C.c_char_p, C.POINTER(Bound), C.c_char_p, vp, vp, vp]),
    "cales_updt_rhs_b": (C.c_int, [vp, C.c_char_p, C.c_char_p, c_int_p, c_int_p, vp, vp, vp, vp]),
    "cales_scale": (C.c_int, [vp, C.c_long, C.c_double, vp, vp]),
    "cales_helmholtz_coeffs": (C.c_int, [vp, C.c_int, C.c_long, C.c_double, vp, vp, vp, vp, vp, vp, vp, vp]),
    "cales_fillps": (C.c_int, [vp, c_int_p, c_dbl_p, vp, C.c_double, vp, vp, vp, vp]),
    "cales_correc": (C.c_int, [vp, c_int_p, c_dbl_p, vp, C.c_double, vp, vp, vp, vp]),
    "cales_updatep": (C.c_int, [vp, c_int_p, c_dbl_p, vp, vp, C.c_double, vp, vp]),
    "cales_cmpt_sgs": (C.c_int, [vp, C.c_char_p, c_int_p, c_int_p, c_int_p, c_int_p, C.c_char_p, C.c_char_p, C.POINTER(Bound),
                                 c_int_p, c_int_p, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p, vp, vp, vp, vp, vp, vp, C.c_double,
                                 C.c_double, c_int_p, vp, vp, vp] + [C.POINTER(Bound)] * 6 + [vp]),
    "cales_set_sgs_options": (C.c_int, [vp, C.c_int, C.c_int]),
    "cales_strain_rate": (C.c_int, [vp, c_int_p, c_dbl_p, vp, vp, vp, vp, vp, vp, vp]),
    "cales_filter3d": (C.c_int, [vp, c_int_p, vp, vp]),
    "cales_chkdt": (C.c_int, [vp, c_int_p, c_dbl_p, vp, vp, C.c_double, vp, vp, vp, vp, c_dbl_p]),
    "cales_chkdiv": (C.c_int, [vp, c_int_p, c_int_p, c_dbl_p, vp, vp, vp, vp, c_dbl_p, c_dbl_p]),
    "cales_fft_lines": (C.c_int, [vp, c_int_p, C.c_int, C.c_char_p, C.c_char, C.c_int, vp]),
    "cales_gaussel": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp]),
    "cales_zdist_emulate": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp]),
    "cales_transpose": (C.c_int, [vp, C.c_int, vp, vp]),
    "cales_updthalo": (C.c_int, [vp, c_int_p, c_int_p, vp]),
    "cales_peer_alloc": (C.c_int, [vp, C.c_char_p, C.c_long, C.POINTER(vp)]),
    "cales_initflow": (C.c_int, [vp, C.c_char_p, c_dbl_p, c_int_p, c_int_p, c_int_p, c_dbl_p, c_dbl_p, vp, vp, vp, vp, C.c_double, c_int_p, c_dbl_p,
                                 c_dbl_p, C.c_int, vp, vp, vp, vp]),
    "cales_out1d_chan": (C.c_int, [vp, c_int_p, c_int_p, c_int_p, c_dbl_p, c_dbl_p, vp, vp, vp, vp, vp, vp, vp, c_dbl_p]),
    "cales_substep": (C.c_int, [vp, C.POINTER(StepArgs), C.c_int, C.c_double]),
    "cales_step": (C.c_int, [vp, C.POINTER(StepArgs), C.c_double, C.c_int]),
    "cales_step_args_layout": (C.c_int, [C.POINTER(C.c_long)]),
}


def load(arith=None):
    """Load one variant of the library ("fma" = libcales_b200.so, "strict" = libcales_b200_strict.so; None = the
    default, DEFAULT_ARITH) and attach the signatures.  Fails loudly when it is absent."""
    arith = arith or DEFAULT_ARITH
    if arith in _libs:
        return _libs[arith]
    if arith not in LIB_PATHS:
        raise CalesError("unknown arithmetic variant %r (fma | strict)" % (arith,))
    path = LIB_PATHS[arith]
    if not os.path.exists(path):
        raise CalesError("%s not found: build it with `make -C cales_b200/csrc` (there is no CPU fallback)" % path)
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    for name, (res, args) in SIGNATURES.items():
        f = getattr(lib, name)
        f.restype = res
        f.argtypes = args
    lib.arith = arith
    _libs[arith] = lib
    return lib


def check(ctx, rc, lib=None):
    if rc != 0:
        msg = (lib or load()).cales_last_error(ctx)
        raise CalesError("libcales_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))
