"""Host driver: replays the call sequence of the reference's `program cans`
(src/main.f90:144-398 set-up, 405-544 time loop) through the C ABI of libcales_b200.so.

This is the executable stand-in for the Fortran host (no Fortran compiler exists in this image;
INTEGRATION.md holds the ISO_C_BINDING module a site build would use).  torch is used only for
device memory, the CUDA stream and the torchrun/NCCL-id plumbing."""
import ctypes as C
import os

import numpy as np
import torch

from . import checkpoint
from . import hostinit
from . import lib as L
from .deck import rkcoeff

_DIFF = {(False, False): 0, (True, False): 1, (True, True): 2}


def _dev(a, device):
    """numpy (any order) -> flat device tensor holding the Fortran-ordered bytes."""
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F"))).to(device)


class DevBound:
    """A `bound` on the device (three planes) + the ctypes struct pointing at them."""

    def __init__(self, planes, device):
        self.t = {ax: _dev(planes[ax], device) for ax in "xyz"}
        self.shape = {ax: planes[ax].shape for ax in "xyz"}
        self.c = L.Bound(self.t["x"].data_ptr(), self.t["y"].data_ptr(), self.t["z"].data_ptr())

    def host(self):
        return {ax: self.t[ax].cpu().numpy().reshape(self.shape[ax], order="F") for ax in "xyz"}


class Simulation:
    def __init__(self, deck, rank=0, nranks=1, uid=None, device=None, ave="channel", arith=None, fused=None, graph=None, filter_2d=False):
        """arith: "fma" (the product build, libcales_b200.so) | "strict" (the bit-exact -fmad=false build) | None = lib.DEFAULT_ARITH
        fused: drive the time loop through the fused entries cales_substep / cales_step (explicit diffusion; identical results)
               instead of the per-procedure sequence; None = CALES_B200_FUSED (default on)
        graph: replay each step from a CUDA graph (single rank, fused); None = CALES_B200_GRAPH (default on)"""
        if not torch.cuda.is_available():
            raise L.CalesError("no CUDA device: cales_b200 has no CPU fallback")
        self.lib = L.load(arith)
        self.deck = deck
        self.rank, self.nranks = rank, nranks
        dev_index = device if device is not None else rank % torch.cuda.device_count()
        self.device = torch.device("cuda", dev_index)
        torch.cuda.set_device(self.device)
        env = lambda k, dflt: os.environ.get(k, dflt) not in ("0", "")
        self.fused = (env("CALES_B200_FUSED", "1") if fused is None else bool(fused)) and not deck.impdiff
        self.graph = (env("CALES_B200_GRAPH", "1") if graph is None else bool(graph)) and self.fused and nranks == 1
        # the library's stream (the reference's OpenACC queue 1).  A CUDA graph cannot be captured on the legacy default
        # stream, so graph replay gets a stream of its own, made current for torch as well (copies, fills, event records).
        self._prev_stream = None
        if self.graph:
            self._prev_stream = torch.cuda.current_stream(self.device)
            self.stream = torch.cuda.Stream(self.device)
            self.stream.wait_stream(self._prev_stream)
            torch.cuda.set_stream(self.stream)
        else:
            self.stream = torch.cuda.current_stream(self.device)
        self.ctx = C.c_void_p()
        diff = _DIFF[(bool(deck.impdiff), bool(deck.impdiff_1d))]
        rc = self.lib.cales_init(C.byref(self.ctx), L._ia(deck.ng), L._ia(deck.dims), deck.ipencil, L._ca(deck.cbcpre), rank,
                                 nranks, uid, dev_index, C.c_void_p(self.stream.cuda_stream), diff)
        L.check(None, rc, self.lib)
        # dsmag averaging geometry / test filter: the reference's cpp switches _DIT/_CHANNEL/_DUCT/_CAVITY and _FILTER_2D (sgs.f90:8)
        self.chk(self.lib.cales_set_sgs_options(self.ctx, {"dit": 0, "channel": 1, "duct": 2, "cavity": 3}[ave], int(bool(filter_2d))))
        arrs = [np.zeros(3, dtype=np.int32) for _ in range(8)] + [np.zeros(6, dtype=np.int32) for _ in range(2)]
        self.chk(self.lib.cales_get_decomp(self.ctx, *[a.ctypes.data_as(L.c_int_p) for a in arrs]))
        (self.lo, self.hi, self.n, self.n_x_fft, self.n_y_fft, self.lo_z, self.hi_z, self.n_z, self.nb, self.is_bound_flat) = arrs
        self.is_bound = self.is_bound_flat.reshape((2, 3), order="F").astype(bool)
        n, ng = self.n, deck.ng
        self.shape = (n[0] + 2, n[1] + 2, n[2] + 2)
        self.ncell = int(np.prod(self.shape))
        dl, l = deck.dl, deck.l
        # grid, main.f90:246-283
        g = hostinit.initgrid(deck.gtype, ng[2], deck.gr, l[2])
        self.dzc_g, self.dzf_g, self.zc_g, self.zf_g = g
        self.dzci_g, self.dzfi_g = self.dzc_g ** (-1), self.dzf_g ** (-1)
        ksl = slice(self.lo[2] - 1, self.hi[2] + 2)
        self.h = {nm: a[ksl].copy() for nm, a in zip(("dzc", "dzf", "zc", "zf"), g)}
        self.h["dzci"] = self.h["dzc"] ** (-1)
        self.h["dzfi"] = self.h["dzf"] ** (-1)
        self.h["gvr_c"] = dl[0] * dl[1] * self.h["dzc"] / (l[0] * l[1] * l[2])
        self.h["gvr_f"] = dl[0] * dl[1] * self.h["dzf"] / (l[0] * l[1] * l[2])
        self.d = {k: _dev(v, self.device) for k, v in self.h.items()}
        # boundary conditions, main.f90:296
        cbcvel, bcu, bcv, bcw, bcp, bcs, index_wm = hostinit.initbc(deck, n, self.is_bound, self.h["zc"], self.h["dzc"])
        self.cbcvel, self.index_wm = cbcvel, index_wm
        mk = lambda pl: DevBound(pl, self.device)
        self.bcu, self.bcv, self.bcw, self.bcp, self.bcs = mk(bcu), mk(bcv), mk(bcw), mk(bcp), mk(bcs)
        self.bcu_mag, self.bcv_mag, self.bcw_mag = mk(bcu), mk(bcv), mk(bcw)
        self.bcuf, self.bcvf, self.bcwf = mk(bcu), mk(bcv), mk(bcw)
        # Poisson solver, main.f90:312-317
        self.poi = self._initsolver(deck.cbcpre, "ccc")
        self.rhsbp = self._new_rhsb()
        self.cmpt_rhs_b(deck.cbcpre, self.bcp, "ccc", self.rhsbp)
        if deck.impdiff:                                      # main.f90:318-346
            self.helm = [self._initsolver(self.cbcvel[:, :, c], cf) for c, cf in enumerate(("fcc", "cfc", "ccf"))]
            self.rhsb_tmp = self._new_rhsb()
            self.rhsb_vel = self._new_rhsb()
            z = lambda m: torch.zeros(m, dtype=torch.float64, device=self.device)
            self.aa, self.bb, self.cc = z(ng[2]), z(ng[2]), z(ng[2])
            self.lam_tmp = z(int(self.n_z[0]) * int(self.n_z[1]))
        # fields, main.f90:358-375
        self.istep, self.time, self.alpha = 0, 0.0, 0.0
        self.fields = {}
        for nm in ("u", "v", "w", "p", "pp", "visct"):
            self.fields[nm] = torch.zeros(self.ncell, dtype=torch.float64, device=self.device)
        self.f = np.zeros(3)
        self.want_f = False
        self.dt = self.dti = self.dt_cfl = 0.0
        self._step_args = None

    def step_args(self):
        """cales_step_args for this simulation (built once; every pointer in it is owned by self and stays alive)."""
        if self._step_args is None:
            d, D, a = self.deck, self.d, L.StepArgs()
            for nm, val in (("n", self.n), ("ng", d.ng), ("lo", self.lo), ("hi", self.hi), ("nb", self.nb), ("is_bound", self.is_bound_flat),
                            ("lwm", L._tab(d.lwm)), ("index_wm", L._tab(self.index_wm)), ("is_forced", [int(bool(x)) for x in d.is_forced])):
                setattr(a, nm, (C.c_int * len(val))(*[int(x) for x in val]))
            for nm, val in (("dl", d.dl), ("dli", d.dli), ("l", d.l), ("velf", d.velf), ("bforce", d.bforce)):
                setattr(a, nm, (C.c_double * 3)(*[float(x) for x in val]))
            a.plan, a.visc, a.hwm, a.normfft = self.poi["plan"], d.visc, d.hwm, self.poi["normfft"]
            a.cbcvel, a.cbcpre, a.cbcsgs, a.sgstype = L._ca(self.cbcvel), L._ca(d.cbcpre), L._ca(d.cbcsgs), d.sgstype.strip().encode()
            for nm, key in (("zc", "zc"), ("zf", "zf"), ("dzc", "dzc"), ("dzf", "dzf"), ("dzci", "dzci"), ("dzfi", "dzfi"),
                            ("grid_vol_ratio_c", "gvr_c"), ("grid_vol_ratio_f", "gvr_f")):
                setattr(a, nm, D[key].data_ptr())
            a.lambdaxy, a.a, a.b, a.c = (self.poi[k].data_ptr() for k in ("lam", "a", "b", "c"))
            a.rhsbx, a.rhsby, a.rhsbz = (self.rhsbp[k].data_ptr() for k in "xyz")
            for nm in ("bcu", "bcv", "bcw", "bcp", "bcs", "bcu_mag", "bcv_mag", "bcw_mag", "bcuf", "bcvf", "bcwf"):
                setattr(a, nm, getattr(self, nm).c)
            for nm in ("u", "v", "w", "p", "pp", "visct"):
                setattr(a, nm, self.fields[nm].data_ptr())
            self._step_args = a
        return self._step_args

    # ---- helpers -------------------------------------------------------------------------------------------
    def chk(self, rc):
        L.check(self.ctx, rc, self.lib)

    def ptr(self, nm):
        return C.c_void_p(self.fields[nm].data_ptr())

    def _new_rhsb(self):
        n = self.n
        z = lambda m: torch.zeros(m, dtype=torch.float64, device=self.device)
        return {"x": z(int(n[1]) * int(n[2]) * 2), "y": z(int(n[0]) * int(n[2]) * 2), "z": z(int(n[0]) * int(n[1]) * 2)}

    def _initsolver(self, cbc, cf):
        deck, ng = self.deck, self.deck.ng
        nzx, nzy = int(self.n_z[0]), int(self.n_z[1])
        lam = np.zeros(nzx * nzy); a = np.zeros(ng[2]); b = np.zeros(ng[2]); c = np.zeros(ng[2])
        plan = C.c_int(-1); normfft = C.c_double(0.)
        self.chk(self.lib.cales_initsolver(self.ctx, L._ia(ng), L._ia(self.n_x_fft), L._ia(self.n_y_fft), L._ia(self.lo_z),
                                           L._ia(self.hi_z), L._da(deck.dli), L._da(self.dzci_g), L._da(self.dzfi_g),
                                           L._ca(cbc), cf.encode(), lam.ctypes.data_as(L.c_dbl_p), a.ctypes.data_as(L.c_dbl_p),
                                           b.ctypes.data_as(L.c_dbl_p), c.ctypes.data_as(L.c_dbl_p), C.byref(plan), C.byref(normfft)))
        return dict(cbc=np.array(cbc), cf=cf, plan=plan.value, normfft=normfft.value, lam_h=lam.reshape((nzx, nzy), order="F"),
                    a_h=a, b_h=b, c_h=c, lam=_dev(lam, self.device), a=_dev(a, self.device), b=_dev(b, self.device), c=_dev(c, self.device))

    def set_fields(self, **kw):
        """Upload haloed Fortran-ordered host arrays."""
        with torch.cuda.stream(self.stream):
            for nm, a in kw.items():
                self.fields[nm].copy_(torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))))

    def get(self, nm):
        with torch.cuda.stream(self.stream):
            return self.fields[nm].cpu().numpy().reshape(self.shape, order="F")

    DEVICE_INIVEL = ("zer", "uni", "cou", "poi", "iop", "hcp", "pdc", "hdc", "tgv", "tgw", "ant", "duc")

    def init_flow(self, mean_allreduce=None, device=None):
        """initflow (main.f90:367).  The deterministic initial conditions are generated directly on the device
        (cales_initflow); the noisy ones ('log', 'hcl', 'tbl') and device=False go through the host (hostinit.initflow)."""
        d = self.deck
        if device is None:
            device = d.inivel.strip() in self.DEVICE_INIVEL
        if device:
            self.chk(self.lib.cales_initflow(self.ctx, d.inivel.strip().encode(), L._da(np.asarray(d.bcvel, dtype=np.float64).ravel(order="F")),
                                             L._ia(d.ng), L._ia(self.lo), L._ia(self.n), L._da(d.l), L._da(d.dl), self.d["zc"].data_ptr(),
                                             self.d["zf"].data_ptr(), self.d["dzc"].data_ptr(), self.d["dzf"].data_ptr(), d.visc,
                                             L._ia([int(bool(x)) for x in d.is_forced]), L._da(d.velf), L._da(d.bforce), int(bool(d.is_wallturb)),
                                             self.ptr("u"), self.ptr("v"), self.ptr("w"), self.ptr("p")))
            return
        u, v, w, p = hostinit.initflow(self.deck, self.lo, self.n, self.h["zc"], self.h["zf"], self.h["dzc"], self.h["dzf"], mean_allreduce)
        self.set_fields(u=u, v=v, w=w, p=p)

    def start(self):
        """main.f90:370-398: ghost cells, initial eddy viscosity, first dt."""
        d = self.deck
        self.bounduvw(True, False)
        self.boundp(d.cbcpre, self.bcp, "p")
        self.cmpt_sgs()
        self.boundp(d.cbcsgs, self.bcs, "visct")
        self.dt_cfl = self.chkdt()
        self.dt = d.dt_f if d.dt_f > 0. else min(d.cfl * self.dt_cfl, d.dtmax)
        self.dti = 1. / self.dt

    # ---- one-to-one wrappers of the ABI -----------------------------------------------------------------------------
    def bounduvw(self, is_updt_wm, is_correc, names=("u", "v", "w"), bcs=None):
        d, D = self.deck, self.d
        bcu, bcv, bcw = bcs or (self.bcu, self.bcv, self.bcw)
        self.chk(self.lib.cales_bounduvw(self.ctx, L._ca(self.cbcvel), L._ia(self.n), C.byref(bcu.c), C.byref(bcv.c), C.byref(bcw.c),
                                         C.byref(self.bcu_mag.c), C.byref(self.bcv_mag.c), C.byref(self.bcw_mag.c),
                                         L._ia(self.nb), L._ia(self.is_bound_flat), L._ia(L._tab(d.lwm)), L._da(d.l), L._da(d.dl),
                                         D["zc"].data_ptr(), D["zf"].data_ptr(), D["dzc"].data_ptr(), D["dzf"].data_ptr(),
                                         d.visc, d.hwm, L._ia(L._tab(self.index_wm)), int(is_updt_wm), int(is_correc),
                                         self.ptr(names[0]), self.ptr(names[1]), self.ptr(names[2])))

    def boundp(self, cbc, bc, name):
        self.chk(self.lib.cales_boundp(self.ctx, L._ca(cbc), L._ia(self.n), C.byref(bc.c), L._ia(self.nb), L._ia(self.is_bound_flat),
                                       L._da(self.deck.dl), self.d["dzc"].data_ptr(), self.ptr(name)))

    def cmpt_rhs_b(self, cbc, bc, cf, out):
        d = self.deck
        self.chk(self.lib.cales_cmpt_rhs_b(self.ctx, L._ia(d.ng), L._ia(self.n), L._da(d.dl), L._da(self.dzc_g), L._da(self.dzf_g),
                                           L._ca(cbc), C.byref(bc.c), cf.encode(), out["x"].data_ptr(), out["y"].data_ptr(),
                                           out["z"].data_ptr()))

    def updt_rhs_b(self, cf, cbc, rhsb, name, skip_xy=False):
        self.chk(self.lib.cales_updt_rhs_b(self.ctx, cf.encode(), L._ca(cbc), L._ia(self.n), L._ia(self.is_bound_flat),
                                           None if skip_xy else rhsb["x"].data_ptr(), None if skip_xy else rhsb["y"].data_ptr(),
                                           rhsb["z"].data_ptr(), self.ptr(name)))

    def rk(self, irk, want_f=True):
        """want_f=False leaves f(3) on the device (no host synchronisation); bulk_forcing(None) then consumes it there."""
        d, D = self.deck, self.d
        f = np.zeros(3)
        self.chk(self.lib.cales_rk(self.ctx, L._da(rkcoeff[irk]), L._ia(self.n), L._da(d.dli), D["dzci"].data_ptr(), D["dzfi"].data_ptr(),
                                   D["gvr_c"].data_ptr(), D["gvr_f"].data_ptr(), d.visc, self.dt, self.ptr("p"),
                                   L._ia(np.array(d.is_forced, dtype=np.int32)), L._da(d.velf), L._da(d.bforce), self.ptr("visct"),
                                   self.ptr("u"), self.ptr("v"), self.ptr("w"), f.ctypes.data_as(L.c_dbl_p) if want_f else None))
        return f if want_f else None

    def bulk_forcing(self, f):
        self.chk(self.lib.cales_bulk_forcing(self.ctx, L._ia(self.n), L._ia(np.array(self.deck.is_forced, dtype=np.int32)),
                                             L._da(f) if f is not None else None, self.ptr("u"), self.ptr("v"), self.ptr("w")))

    def fillps(self, dti):
        self.chk(self.lib.cales_fillps(self.ctx, L._ia(self.n), L._da(self.deck.dli), self.d["dzfi"].data_ptr(), dti,
                                       self.ptr("u"), self.ptr("v"), self.ptr("w"), self.ptr("pp")))

    def solver(self, s, name, lam=None, a=None, b=None, c=None):
        self.chk(self.lib.cales_solver(self.ctx, L._ia(self.n), L._ia(self.deck.ng), s["plan"], s["normfft"],
                                       (lam if lam is not None else s["lam"]).data_ptr(), (a if a is not None else s["a"]).data_ptr(),
                                       (b if b is not None else s["b"]).data_ptr(), (c if c is not None else s["c"]).data_ptr(),
                                       L._ca(s["cbc"]), s["cf"].encode(), self.ptr(name)))

    def correc(self, dt):
        self.chk(self.lib.cales_correc(self.ctx, L._ia(self.n), L._da(self.deck.dli), self.d["dzci"].data_ptr(), dt, self.ptr("pp"),
                                       self.ptr("u"), self.ptr("v"), self.ptr("w")))

    def updatep(self):
        self.chk(self.lib.cales_updatep(self.ctx, L._ia(self.n), L._da(self.deck.dli), self.d["dzci"].data_ptr(), self.d["dzfi"].data_ptr(),
                                        self.alpha, self.ptr("pp"), self.ptr("p")))

    def cmpt_sgs(self):
        d, D = self.deck, self.d
        self.chk(self.lib.cales_cmpt_sgs(self.ctx, d.sgstype.strip().encode(), L._ia(self.n), L._ia(d.ng), L._ia(self.lo), L._ia(self.hi),
                                         L._ca(self.cbcvel), L._ca(d.cbcsgs), C.byref(self.bcs.c), L._ia(self.nb), L._ia(self.is_bound_flat),
                                         L._ia(L._tab(d.lwm)), L._da(d.l), L._da(d.dl), L._da(d.dli), D["zc"].data_ptr(), D["zf"].data_ptr(),
                                         D["dzc"].data_ptr(), D["dzf"].data_ptr(), D["dzci"].data_ptr(), D["dzfi"].data_ptr(), d.visc,
                                         d.hwm, L._ia(L._tab(self.index_wm)), self.ptr("u"), self.ptr("v"), self.ptr("w"),
                                         C.byref(self.bcuf.c), C.byref(self.bcvf.c), C.byref(self.bcwf.c), C.byref(self.bcu_mag.c),
                                         C.byref(self.bcv_mag.c), C.byref(self.bcw_mag.c), self.ptr("visct")))

    def chkdt(self):
        d, D = self.deck, self.d
        out = C.c_double(0.)
        self.chk(self.lib.cales_chkdt(self.ctx, L._ia(self.n), L._da(d.dl), D["dzci"].data_ptr(), D["dzfi"].data_ptr(), d.visc,
                                      self.ptr("visct"), self.ptr("u"), self.ptr("v"), self.ptr("w"), C.byref(out)))
        return out.value

    def chkdiv(self):
        tot, mx = C.c_double(0.), C.c_double(0.)
        # chkdiv indexes its arrays from lo-1 (chkdiv.f90:24-26): the device pointers are the same local arrays
        self.chk(self.lib.cales_chkdiv(self.ctx, L._ia(self.lo), L._ia(self.hi), L._da(self.deck.dli), self.d["dzfi"].data_ptr(),
                                       self.ptr("u"), self.ptr("v"), self.ptr("w"), C.byref(tot), C.byref(mx)))
        return tot.value, mx.value

    def forcing_log(self):
        """what main.f90:554-572 writes to forcing.out: the pressure gradient that keeps the bulk velocity, dpdl = -sum_rk f / dt
        (main.f90:492, 508; -bforce when nothing is forced, main.f90:567), and the bulk means of the forced / driven components."""
        d = self.deck
        means = [0., 0., 0.]
        for c, (nm, g) in enumerate((("u", "gvr_f"), ("v", "gvr_f"), ("w", "gvr_c"))):
            if d.is_forced[c] or abs(d.bforce[c]) > 0.:
                out = C.c_double(0.)
                self.chk(self.lib.cales_bulk_mean(self.ctx, L._ia(self.n), self.d[g].data_ptr(), self.ptr(nm), C.byref(out)))
                means[c] = out.value
        if not any(d.is_forced):
            dpdl = [-float(b) for b in d.bforce]
        else:
            dpdl = list(getattr(self, "dpdl", [0., 0., 0.]))
        return dpdl, means

    def out1d_chan(self, fname=None):
        """on-the-fly channel statistics (out1d_single_point_chan, src/output.f90:509-691): the 27 profiles; see stats.py"""
        from . import stats
        return stats.out1d_chan(self, fname)

    def synchronize(self):
        self.chk(self.lib.cales_stream_synchronize(self.ctx))

    # ---- the time loop ---------------------------------------------------------------------------------------------------
    def substep(self, irk):
        """main.f90:418-506."""
        d = self.deck
        dtrk = (rkcoeff[irk][0] + rkcoeff[irk][1]) * self.dt
        dtrki = dtrk ** (-1)
        # f(3) stays on the device inside the time loop (the reference reads it back only for its log, main.f90:559-565)
        self.f = self.rk(irk, want_f=self.want_f)
        if self.want_f:                                        # dpdl(:) = dpdl(:) + f(:), main.f90:492
            self._dpdl_acc = [a + b for a, b in zip(getattr(self, "_dpdl_acc", [0., 0., 0.]) if irk else [0., 0., 0.], self.f)]
            if irk == 2:
                self.dpdl = [-a * self.dti for a in self._dpdl_acc]        # main.f90:508
        self.bulk_forcing(self.f)
        if d.impdiff:                                          # main.f90:423-491
            alpha = -.5 * d.visc * dtrk
            self.alpha = alpha
            for c, (nm, bc) in enumerate((("u", self.bcu), ("v", self.bcv), ("w", self.bcw))):
                hs = self.helm[c]
                self.cmpt_rhs_b(self.cbcvel[:, :, c], bc, hs["cf"], self.rhsb_vel)
                for ax in "xyz":
                    if d.impdiff_1d and ax != "z":
                        continue
                    t = self.rhsb_vel[ax]
                    self.chk(self.lib.cales_scale(self.ctx, t.numel(), alpha, t.data_ptr(), self.rhsb_tmp[ax].data_ptr()))
                self.updt_rhs_b(hs["cf"], self.cbcvel[:, :, c], self.rhsb_tmp, nm, skip_xy=d.impdiff_1d)
                self.chk(self.lib.cales_helmholtz_coeffs(self.ctx, d.ng[2], self.lam_tmp.numel(), alpha, hs["a"].data_ptr(),
                                                         hs["b"].data_ptr(), hs["c"].data_ptr(),
                                                         None if d.impdiff_1d else hs["lam"].data_ptr(), self.aa.data_ptr(),
                                                         self.bb.data_ptr(), self.cc.data_ptr(),
                                                         None if d.impdiff_1d else self.lam_tmp.data_ptr()))
                if not d.impdiff_1d:
                    self.solver(hs, nm, lam=self.lam_tmp, a=self.aa, b=self.bb, c=self.cc)
                else:
                    bcz = self.cbcvel[:, 2, c]
                    self.chk(self.lib.cales_solver_gaussel_z(self.ctx, L._ia(self.n), self.aa.data_ptr(), self.bb.data_ptr(),
                                                             self.cc.data_ptr(), (bcz[0] + bcz[1]).encode(), hs["cf"].encode(),
                                                             self.ptr(nm)))
        self.bounduvw(True, False)
        self.fillps(dtrki)
        self.updt_rhs_b("ccc", d.cbcpre, self.rhsbp, "pp")
        self.solver(self.poi, "pp")
        self.boundp(d.cbcpre, self.bcp, "pp")
        self.correc(dtrk)
        self.bounduvw(True, True)
        self.updatep()
        self.boundp(d.cbcpre, self.bcp, "p")
        self.cmpt_sgs()
        self.boundp(d.cbcsgs, self.bcs, "visct")

    def step(self, icheck=0):
        """main.f90:405-544 for one time step; returns (divtot, divmax) when checked."""
        self.istep += 1
        self.time += self.dt
        if self.fused:
            self.chk(self.lib.cales_step(self.ctx, C.byref(self.step_args()), self.dt, int(self.graph)))
        else:
            for irk in range(3):
                self.substep(irk)
        if icheck > 0 and self.istep % icheck == 0:
            d = self.deck
            self.dt_cfl = self.chkdt()
            self.dt = d.dt_f if d.dt_f > 0. else min(d.cfl * self.dt_cfl, d.dtmax)
            self.dti = 1. / self.dt
            return self.chkdiv()
        return None

    # ---- restart files (main.f90:358-369, 590-611) ---------------------------------------------------------------
    def save(self, filename, barrier=None):
        """`load_all('w', ...)`: download u,v,w,p (`!$acc update self`, main.f90:607) and write this rank's sub-box."""
        torch.cuda.synchronize(self.device)
        f = [self.get(nm) for nm in ("u", "v", "w", "p")]
        checkpoint.load_all("w", filename, self.deck.ng, self.lo, self.hi, *f, time=self.time, istep=self.istep, rank=self.rank,
                            barrier=barrier)

    def load(self, filename):
        """`load_all('r', ...)` instead of `initflow` (main.f90:365); call `start()` afterwards as after `init_flow()`."""
        f = [np.zeros(self.shape, order="F") for _ in range(4)]
        self.time, self.istep = checkpoint.load_all("r", filename, self.deck.ng, self.lo, self.hi, *f)
        self.set_fields(u=f[0], v=f[1], w=f[2], p=f[3])

    def close(self):
        if self.ctx:
            self.lib.cales_finalize(self.ctx)
            self.ctx = C.c_void_p()
            if self._prev_stream is not None:
                self._prev_stream.wait_stream(self.stream)
                if torch.cuda.current_stream(self.device) == self.stream:
                    torch.cuda.set_stream(self._prev_stream)
                self._prev_stream = None
