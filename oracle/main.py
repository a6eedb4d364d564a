"""Driver: the call sequence of `program cans`.  Follows src/main.f90:144-398 (setup) and 405-544
(time loop: 3 RK substeps, then chkdt/chkdiv every `icheck` steps).  All ranks are emulated
serially through decomp.World; a one-rank World is the single-process reference run."""
import numpy as np

from . import bound as bnd
from . import ops, rk as rkmod, sgs as sgsmod, solver as slv
from .decomp import World
from .initflow import initflow
from .initgrid import initgrid
from .param import rkcoeff, small


class RankState:
    pass


class Sim:
    def __init__(self, deck, ave="channel", filter_2d=False):
        """ave / filter_2d: the cpp switches _DIT/_CHANNEL/_DUCT/_CAVITY and _FILTER_2D of src/sgs.f90 (default: the reference's
        hard-wired `#define _CHANNEL`, sgs.f90:8, with the 3-D test filter)"""
        self.deck = deck
        self.ave = ave
        self.filter_2d = filter_2d
        ng = deck.ng
        self.world = World(ng, deck.dims, deck.cbcpre, deck.ipencil)
        w = self.world
        # main.f90:246 global grid, 261-283 local copies and inverses
        self.dzc_g, self.dzf_g, self.zc_g, self.zf_g = initgrid(deck.gtype, ng[2], deck.gr, deck.l[2])
        self.dzci_g = self.dzc_g ** (-1)
        self.dzfi_g = self.dzf_g ** (-1)
        self.st = []
        dl, l = deck.dl, deck.l
        for r in w.ranks:
            s = RankState()
            ksl = slice(r.lo[2] - 1, r.hi[2] + 2)
            s.zc, s.zf, s.dzc, s.dzf = self.zc_g[ksl].copy(), self.zf_g[ksl].copy(), self.dzc_g[ksl].copy(), self.dzf_g[ksl].copy()
            s.dzci = s.dzc ** (-1)
            s.dzfi = s.dzf ** (-1)
            s.grid_vol_ratio_c = dl[0] * dl[1] * s.dzc / (l[0] * l[1] * l[2])
            s.grid_vol_ratio_f = dl[0] * dl[1] * s.dzf / (l[0] * l[1] * l[2])
            s.l, s.dl, s.dli, s.visc, s.hwm, s.lwm = l, dl, deck.dli, deck.visc, deck.hwm, deck.lwm
            (cbcvel, s.bcu, s.bcv, s.bcw, s.bcp, s.bcs, s.bcu_mag, s.bcv_mag, s.bcw_mag,
             s.bcuf, s.bcvf, s.bcwf, s.index_wm) = bnd.initbc(deck, r.n, r.is_bound, s.zc, s.dzc)   # main.f90:296
            self.cbcvel = cbcvel
            s.rk = rkmod.RkState(r.n)
            s.sgs = sgsmod.SgsState()
            self.st.append(s)
        # Poisson solver, main.f90:312-317
        ccc = ("c", "c", "c")
        self.lambdaxyp, self.plan_p = [], None
        for r in w.ranks:
            lam, self.ap, self.bp, self.cp, self.plan_p = slv.initsolver(ng, r.lo_z, r.hi_z, deck.dli, self.dzci_g,
                                                                        self.dzfi_g, deck.cbcpre, ccc)
            self.lambdaxyp.append(lam)
        self.rhsbp = []
        for r, s in zip(w.ranks, self.st):
            self.rhsbp.append(bnd.cmpt_rhs_b(ng, dl, self.dzc_g, self.dzf_g, deck.cbcpre, s.bcp, ccc))
        if deck.impdiff:                                    # main.f90:318-346
            self.helm = []
            for c, cf in enumerate((("f", "c", "c"), ("c", "f", "c"), ("c", "c", "f"))):
                lams = []
                for r in w.ranks:
                    lam, a, b, c_, plan = slv.initsolver(ng, r.lo_z, r.hi_z, deck.dli, self.dzci_g, self.dzfi_g,
                                                         self.cbcvel[:, :, c], cf)
                    lams.append(lam)
                self.helm.append(dict(cf=cf, lam=lams, a=a, b=b, c=c_, plan=plan))
        # fields, main.f90:358-375
        self.istep = 0
        self.time = 0.
        U, V, W, P = [], [], [], []
        for r, s in zip(w.ranks, self.st):
            u, v, ww, p = initflow(deck, r.lo, r.n, s.zc, s.zf, s.dzc, s.dzf)
            U.append(u); V.append(v); W.append(ww); P.append(p)
        if deck.inivel in ("poi", "iop", "pdc", "duc") and w.nranks > 1:
            # set_mean uses a global MPI_ALLREDUCE (initflow.f90:317-333): redo with the global sum
            U, V, W, P = self._initflow_global_mean()
        self.U, self.V, self.W, self.P = U, V, W, P
        self.PP = w.zeros()
        self.VISCT = w.zeros()
        self.alpha = 0.
        self.bounduvw(True, False)
        bnd.boundp(w, deck.cbcpre, self.st, "bcp", self.P)
        self.cmpt_sgs()
        bnd.boundp(w, deck.cbcsgs, self.st, "bcs", self.VISCT)
        self.dt_cfl = self.chkdt()                          # main.f90:395-398
        self.dt = deck.dt_f if deck.dt_f > 0. else min(deck.cfl * self.dt_cfl, deck.dtmax)
        self.dti = 1. / self.dt
        self.f = [0., 0., 0.]

    def _initflow_global_mean(self):
        """Rank-count independent variant of the set_mean reduction: per-rank partial sums are
        combined in rank order like MPI_SUM would on one communicator."""
        w, deck = self.world, self.deck
        # partial sums
        parts = []
        for r, s in zip(w.ranks, self.st):
            cap = {}

            def grab(x, cap=cap):
                cap["v"] = x
                return float("nan")
            initflow(deck, r.lo, r.n, s.zc, s.zf, s.dzc, s.dzf, allreduce_sum=grab)
            parts.append(cap["v"])
        tot = w.allreduce_sum(parts)
        U, V, W, P = [], [], [], []
        for r, s in zip(w.ranks, self.st):
            u, v, ww, p = initflow(deck, r.lo, r.n, s.zc, s.zf, s.dzc, s.dzf, allreduce_sum=lambda x: tot)
            U.append(u); V.append(v); W.append(ww); P.append(p)
        return U, V, W, P

    # -- thin wrappers ---------------------------------------------------------------------------
    def bounduvw(self, is_updt_wm, is_correc):
        bnd.bounduvw(self.world, self.cbcvel, self.st, is_updt_wm, is_correc, self.U, self.V, self.W)

    def cmpt_sgs(self):
        sgsmod.cmpt_sgs(self.world, self.st, self.deck, self.cbcvel, self.U, self.V, self.W, self.VISCT, ave=self.ave, filter_2d=self.filter_2d)

    def chkdt(self):
        d = self.deck
        vals = [ops.chkdt_local(r.n, s.dl, s.dzci, s.dzfi, s.visc, self.VISCT[r.id], self.U[r.id], self.V[r.id],
                                self.W[r.id], d.impdiff, d.impdiff_1d) for r, s in zip(self.world.ranks, self.st)]
        return min(vals)

    def chkdiv(self):
        vals = [ops.chkdiv_local(r.n, s.dli, s.dzfi, self.U[r.id], self.V[r.id], self.W[r.id])
                for r, s in zip(self.world.ranks, self.st)]
        return self.world.allreduce_sum([v[0] for v in vals]), max(v[1] for v in vals)

    def solve_poisson(self, PP):
        slv.solver(self.world, self.plan_p, self.lambdaxyp, self.ap, self.bp, self.cp, self.deck.cbcpre,
                   ("c", "c", "c"), PP)

    # -- one RK substep, main.f90:417-507 -------------------------------------------------------------
    def substep(self, irk):
        d, w, st = self.deck, self.world, self.st
        R = w.ranks
        dt = self.dt
        dtrk = (rkcoeff[irk][0] + rkcoeff[irk][1]) * dt
        dtrki = dtrk ** (-1)
        f = rkmod.rk(w, st, rkcoeff[irk], dt, self.P, self.VISCT, self.U, self.V, self.W, d)
        self.f = f
        for r in R:
            from .mom import bulk_forcing
            bulk_forcing(r.n, d.is_forced, f, self.U[r.id], self.V[r.id], self.W[r.id])
        if d.impdiff:                                       # main.f90:423-491
            alpha = -.5 * d.visc * dtrk
            self.alpha = alpha
            for c, F in enumerate((self.U, self.V, self.W)):
                h = self.helm[c]
                for r, s in zip(R, st):
                    bc = (s.bcu, s.bcv, s.bcw)[c]
                    rx, ry, rz = bnd.cmpt_rhs_b(d.ng, d.dl, self.dzc_g, self.dzf_g, self.cbcvel[:, :, c], bc, h["cf"])
                    # the rhsb planes are for the GLOBAL face; restrict to this rank's patch
                    if d.impdiff_1d:
                        rx = ry = None
                    else:
                        rx, ry = rx * alpha, ry * alpha
                    rz = rz * alpha
                    bnd.updt_rhs_b(h["cf"], self.cbcvel[:, :, c], r.n, r.is_bound, rx, ry, rz, F[r.id])
                aa = h["a"] * alpha
                bb = h["b"] * alpha + 1.
                cc = h["c"] * alpha
                if not d.impdiff_1d:
                    lam = [x * alpha for x in h["lam"]]
                    slv.solver(w, h["plan"], lam, aa, bb, cc, self.cbcvel[:, :, c], h["cf"], F)
                else:
                    slv.solver_gaussel_z(w, aa, bb, cc, self.cbcvel[:, 2, c], h["cf"], F)
        self.bounduvw(True, False)                          # main.f90:493
        for r, s in zip(R, st):
            ops.fillps(r.n, s.dli, s.dzfi, dtrki, self.U[r.id], self.V[r.id], self.W[r.id], self.PP[r.id])
            rx, ry, rz = self.rhsbp[r.id]
            bnd.updt_rhs_b(("c", "c", "c"), d.cbcpre, r.n, r.is_bound, rx, ry, rz, self.PP[r.id])
        self.solve_poisson(self.PP)
        bnd.boundp(w, d.cbcpre, st, "bcp", self.PP)
        for r, s in zip(R, st):
            ops.correc(r.n, s.dli, s.dzci, dtrk, self.PP[r.id], self.U[r.id], self.V[r.id], self.W[r.id])
        self.bounduvw(True, True)                           # main.f90:500
        for r, s in zip(R, st):
            ops.updatep(r.n, s.dli, s.dzci, s.dzfi, self.alpha, self.PP[r.id], self.P[r.id], d.impdiff, d.impdiff_1d)
        bnd.boundp(w, d.cbcpre, st, "bcp", self.P)
        self.cmpt_sgs()
        bnd.boundp(w, d.cbcsgs, st, "bcs", self.VISCT)

    def step(self, icheck=0):
        """main.f90:405-544 for one time step.  Returns (divtot,divmax) when checked."""
        self.istep += 1
        self.time += self.dt
        for irk in range(3):
            self.substep(irk)
        if icheck > 0 and self.istep % icheck == 0:
            self.dt_cfl = self.chkdt()
            d = self.deck
            self.dt = d.dt_f if d.dt_f > 0. else min(d.cfl * self.dt_cfl, d.dtmax)
            self.dti = 1. / self.dt
            return self.chkdiv()
        return None
