"""ctypes front end of oracle/c/libcales_cpu.so -- the C/OpenMP restatement of the explicit RK3 step with the static or the
dynamic Smagorinsky model for the tri-periodic deck and for the plane channel (z walls, forcing, van Driest damping, optional log-law wall model;
see the header of oracle/c/cales_cpu.c for the reference lines it follows).
TEST INFRASTRUCTURE ONLY: the checker in tests/ and the CPU arm of bench.py."""
import ctypes as C
import os

import numpy as np

_LIB = None
FIELDS = {"u": 0, "v": 1, "w": 2, "p": 3, "pp": 4, "visct": 5, "s0": 6}


def load():
    global _LIB
    if _LIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c", "libcales_cpu.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/c/libcales_cpu.so is missing: run `make -C oracle/c` (or __graft_entry__.build())")
        lib = C.CDLL(path)
        dp = C.POINTER(C.c_double)
        lib.cales_cpu_new.restype = C.c_void_p
        lib.cales_cpu_new.argtypes = [C.c_int, C.c_int, C.c_int, dp, C.c_double, dp, dp]
        lib.cales_cpu_field.restype = dp
        lib.cales_cpu_field.argtypes = [C.c_void_p, C.c_int]
        lib.cales_cpu_lambdaxy.restype = dp
        lib.cales_cpu_lambdaxy.argtypes = [C.c_void_p]
        for nm in ("start", "cmpt_sgs", "solver", "free"):
            getattr(lib, "cales_cpu_" + nm).argtypes = [C.c_void_p]
            getattr(lib, "cales_cpu_" + nm).restype = None
        lib.cales_cpu_boundp.argtypes = [C.c_void_p, C.c_int]; lib.cales_cpu_boundp.restype = None
        for nm in ("fillps", "correc", "step"):
            getattr(lib, "cales_cpu_" + nm).argtypes = [C.c_void_p, C.c_double]
            getattr(lib, "cales_cpu_" + nm).restype = None
        lib.cales_cpu_rk.argtypes = [C.c_void_p, C.c_int, C.c_double]; lib.cales_cpu_rk.restype = None
        lib.cales_cpu_divmax.argtypes = [C.c_void_p]; lib.cales_cpu_divmax.restype = C.c_double
        lib.cales_cpu_chkdt.argtypes = [C.c_void_p]; lib.cales_cpu_chkdt.restype = C.c_double
        lib.cales_cpu_threads.restype = C.c_int
        ip = C.POINTER(C.c_int)
        lib.cales_cpu_set_channel.restype = C.c_int
        lib.cales_cpu_set_channel.argtypes = [C.c_void_p, dp, dp, ip, C.c_double, ip, dp, dp]
        cp = C.c_char_p
        lib.cales_cpu_set_bc.restype = C.c_int
        lib.cales_cpu_set_bc.argtypes = [C.c_void_p, cp, dp, cp, cp, dp, dp, ip, dp, dp]
        lib.cales_cpu_set_wm.restype = C.c_int
        lib.cales_cpu_set_wm.argtypes = [C.c_void_p, ip, C.c_double]
        lib.cales_cpu_set_sgs.restype = None
        lib.cales_cpu_set_sgs.argtypes = [C.c_void_p, C.c_int]
        lib.cales_cpu_forcing.restype = C.c_double
        lib.cales_cpu_forcing.argtypes = [C.c_void_p, C.c_int]
        lib.cales_cpu_set_threads.argtypes = [C.c_int]; lib.cales_cpu_set_threads.restype = None
        _LIB = lib
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class CSim:
    """Mirror of oracle.main.Sim for the decks the C library covers (all-periodic and plane channel with either model; square
    duct and lid-driven cavity with the static model) on one rank."""

    @staticmethod
    def kind(deck):
        """'periodic', 'channel' or None (not covered by the C restatement)"""
        if deck.sgstype.strip() not in ("none", "smag", "dsmag") or deck.impdiff or tuple(deck.dims) != (1, 1):
            return None
        if (deck.cbcvel == "P").all() and (deck.cbcpre == "P").all() and not any(deck.is_forced) and not any(deck.bforce) \
                and not deck.lwm.any():
            return "periodic"
        ok = (deck.cbcvel[:, 0:2, :] == "P").all() and (deck.cbcvel[:, 2, :] == "D").all() and (deck.cbcpre[:, 0:2] == "P").all() and \
            (deck.cbcpre[:, 2] == "N").all() and (deck.cbcsgs[:, 0:2] == "P").all() and (deck.cbcsgs[:, 2] == "D").all() and \
            not deck.bcvel.any() and not deck.bcpre.any() and not deck.bcsgs.any() and not deck.lwm[:, 0:2].any() and \
            all(int(x) in (0, 1) for x in deck.lwm[:, 2])
        if ok:
            return "channel"
        pairs = [deck.cbcpre[0, q] + deck.cbcpre[1, q] for q in range(3)]
        wm_ok = not deck.lwm[:, 0].any() and all(int(x) in (0, 1) for x in deck.lwm.ravel()) and \
            all(pairs[q] == "NN" and (deck.cbcvel[:, q, :] == "D").all() for q in (1, 2) if deck.lwm[:, q].any())
        ok = deck.sgstype.strip() in ("none", "smag") and all(p in ("PP", "NN") for p in pairs) and wm_ok and \
            not deck.bcpre.any() and not deck.bcsgs.any() and \
            all(((deck.cbcvel[:, q, :] == "P").all() and (deck.cbcsgs[:, q] == "P").all()) if pairs[q] == "PP" else
                (np.isin(deck.cbcvel[:, q, :], ("D", "N")).all() and np.isin(deck.cbcsgs[:, q], ("D", "N")).all()) for q in range(3))
        return "walls" if ok else None                            # square duct, lid-driven cavity

    def __init__(self, deck, threads=None):
        """threads: OpenMP threads to use (None = the OpenMP default, i.e. OMP_NUM_THREADS or all cores)."""
        from .initflow import initflow
        from .initgrid import initgrid
        kind = self.kind(deck)
        assert kind is not None, "deck not covered by the C restatement (oracle/c)"
        self.lib = load()
        if threads:
            self.lib.cales_cpu_set_threads(int(threads))
        self.deck = deck
        n = self.n = tuple(int(x) for x in deck.ng)
        dzc, dzf, zc, zf = initgrid(deck.gtype, n[2], deck.gr, deck.l[2])
        l = np.array(deck.l, dtype=np.float64)
        self.h = C.c_void_p(self.lib.cales_cpu_new(n[0], n[1], n[2], _dp(l), float(deck.visc), _dp(np.ascontiguousarray(dzc)),
                                                   _dp(np.ascontiguousarray(dzf))))
        if kind == "channel":
            ia = lambda a: (C.c_int * len(a))(*[int(x) for x in a])
            da = lambda a: (C.c_double * len(a))(*[float(x) for x in a])
            rc = self.lib.cales_cpu_set_channel(self.h, _dp(np.ascontiguousarray(zc)), _dp(np.ascontiguousarray(zf)), ia(deck.lwm[:, 2]),
                                                float(deck.hwm), ia([bool(x) for x in deck.is_forced]), da(deck.velf), da(deck.bforce))
            assert rc == 0
        if kind == "walls":
            ia = lambda a: (C.c_int * len(a))(*[int(x) for x in a])
            da = lambda a: (C.c_double * len(a))(*[float(x) for x in a])
            fl = lambda a: "".join(np.asarray(a).ravel(order="F")).encode()
            rc = self.lib.cales_cpu_set_bc(self.h, fl(deck.cbcvel), da(np.asarray(deck.bcvel, dtype=float).ravel(order="F")), fl(deck.cbcpre),
                                           fl(deck.cbcsgs), _dp(np.ascontiguousarray(zc)), _dp(np.ascontiguousarray(zf)),
                                           ia([bool(x) for x in deck.is_forced]), da(deck.velf), da(deck.bforce))
            assert rc == 0
            if deck.lwm.any():
                assert self.lib.cales_cpu_set_wm(self.h, ia(np.asarray(deck.lwm).ravel(order="F")), float(deck.hwm)) == 0
        self.lib.cales_cpu_set_sgs(self.h, {"smag": 0, "dsmag": 1, "none": 2}[deck.sgstype.strip()])
        shp = (n[0] + 2, n[1] + 2, n[2] + 2)
        self.f = {nm: np.ctypeslib.as_array(self.lib.cales_cpu_field(self.h, i), shape=shp[::-1]).T for nm, i in FIELDS.items()}
        u, v, w, p = initflow(deck, (1, 1, 1), n, zc, zf, dzc, dzf)
        for nm, a in (("u", u), ("v", v), ("w", w), ("p", p)):
            self.f[nm][...] = a
        self.lib.cales_cpu_start(self.h)                                # main.f90:370-375
        self.dt_cfl = self.lib.cales_cpu_chkdt(self.h)                  # main.f90:395-398
        self.dt = deck.dt_f if deck.dt_f > 0. else min(deck.cfl * self.dt_cfl, deck.dtmax)
        self.istep = 0

    def step(self, icheck=0):
        self.istep += 1
        self.lib.cales_cpu_step(self.h, self.dt)
        if icheck > 0 and self.istep % icheck == 0:
            self.dt_cfl = self.lib.cales_cpu_chkdt(self.h)
            d = self.deck
            self.dt = d.dt_f if d.dt_f > 0. else min(d.cfl * self.dt_cfl, d.dtmax)
            return self.lib.cales_cpu_divmax(self.h)
        return None

    def forcing(self):
        """f(1:3) of the last substep (rk.f90:197-222)"""
        return [self.lib.cales_cpu_forcing(self.h, m) for m in range(3)]

    def threads(self):
        return int(self.lib.cales_cpu_threads())

    def close(self):
        if self.h:
            self.f = {}
            self.lib.cales_cpu_free(self.h)
            self.h = None
