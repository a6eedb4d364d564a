#!/usr/bin/env python
"""Kernel micro-benchmarks through the C ABI (CUDA events on the launch stream, working set > L2).

  python tools/kbench.py [--ng 256 256 256] [--only fft,gaussel,...] [--iters 10]

Prints one line per kernel: ms, algorithmic GB/s (SURVEY.md 8(d) bytes/cell) and the fraction of the measured HBM peak."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p)).get("hbm_gbs", 6650.0) if os.path.exists(p) else 6650.0


WARM = 3


def timeit(fn, iters):
    for _ in range(WARM):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ng", type=int, nargs=3, default=[256, 256, 256])
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--deck", default="tgv", choices=["tgv", "channel"])
    ap.add_argument("--warm", type=int, default=3)
    ap.add_argument("--sgs", default="smag", choices=["smag", "dsmag", "none"])
    ap.add_argument("--impdiff", default="", choices=["", "3d", "1d"], help="Crank-Nicolson diffusion (_IMPDIFF / _IMPDIFF_1D): times the per-procedure substep")
    ap.add_argument("--wall-model", action="store_true")
    args = ap.parse_args()
    global WARM
    WARM = args.warm
    from cales_b200 import lib as L
    from cales_b200 import deck as pd
    from cales_b200.driver import Simulation
    only = set(x for x in args.only.split(",") if x)
    ng = tuple(args.ng)
    if args.deck == "tgv":
        deck = pd.deck_tgv(ng=ng, sgstype=args.sgs)
    elif args.wall_model:
        deck = pd.deck_channel(ng=ng, sgstype=args.sgs, wall_model=True, gtype=6, gr=0., l=(12.8, 4.8, 2.), visci=43500.)
    else:
        deck = pd.deck_channel(ng=ng, sgstype=args.sgs)
    if args.impdiff:
        deck.impdiff = True; deck.impdiff_1d = args.impdiff == "1d"
    sim = Simulation(deck)
    print("# %s %s sgs=%s impdiff=%s arith=%s fused=%s graph=%s" % (args.deck, ng, args.sgs, args.impdiff or "no", sim.lib.arith, sim.fused, sim.graph))
    sim.init_flow(); sim.start()
    sim.step()
    n = sim.n; nn = L._ia(n); D = sim.d; d = deck
    ncell = float(np.prod(n))
    pk = peak()
    lib = sim.lib
    scr = [torch.zeros(int(ncell), dtype=torch.float64, device="cuda") for _ in range(3)]
    wk = torch.randn(int(ncell), dtype=torch.float64, device="cuda")
    rows = []

    def add(name, bpc, fn):
        if only and not any(name.startswith(o) for o in only):
            return
        ms = timeit(fn, args.iters)
        gbs = bpc * ncell / ms / 1e6
        rows.append((name, ms, bpc, gbs, gbs / pk))
        print("%-22s %8.4f ms  %4d B/cell  %8.1f GB/s  %5.1f%% of %.0f" % (name, ms, bpc, gbs, 100 * gbs / pk, pk), flush=True)

    add("mom_xyz_ad", 56, lambda: sim.chk(lib.cales_mom_xyz_ad(
        sim.ctx, nn, d.dli[0], d.dli[1], D["dzci"].data_ptr(), D["dzfi"].data_ptr(), d.visc, sim.ptr("u"), sim.ptr("v"), sim.ptr("w"),
        sim.ptr("visct"), scr[0].data_ptr(), scr[1].data_ptr(), scr[2].data_ptr(), None, None, None)))
    for bc in (b"PP", b"NN", b"DD"):
        for dir_ in (0, 1):
            for bw in (0, 1):
                add("fft_%s_%s_%s" % (bc.decode(), "xy"[dir_], "bwd" if bw else "fwd"), 16,
                    lambda: sim.chk(lib.cales_fft_lines(sim.ctx, nn, dir_, bc, b"c", bw, wk.data_ptr())))
    # cuFFT as the comparison point (north star; what src/fft.f90:86-97,247-272 calls): batched D2Z / Z2D over the same lines of
    # the same array through torch.fft (out of place: 8 B/cell read + ~8 B/cell written, like the hand-written passes)
    a3 = wk.view(int(n[2]), int(n[1]), int(n[0]))                 # C order (k, j, i) == Fortran (i, j, k)
    for dname, dim in (("x", 2), ("y", 1)):
        spec = torch.fft.rfft(a3, dim=dim)
        add("cufft_D2Z_%s" % dname, 16, lambda: torch.fft.rfft(a3, dim=dim))
        add("cufft_Z2D_%s" % dname, 16, lambda: torch.fft.irfft(spec, n=a3.shape[dim], dim=dim))
        del spec
    for per in (1, 0):
        wk.normal_()
        add("gaussel_%s" % ("periodic" if per else "nonper"), 16, lambda: sim.chk(lib.cales_gaussel(
            sim.ctx, int(n[0]), int(n[1]), int(n[2]), per, sim.poi["a"].data_ptr(), sim.poi["b"].data_ptr(), sim.poi["c"].data_ptr(),
            sim.poi["lam"].data_ptr(), wk.data_ptr())))
    add("solver", 80, lambda: sim.solver(sim.poi, "pp"))
    add("fillps", 32, lambda: sim.fillps(1.0))
    add("correc", 56, lambda: sim.correc(0.0))
    add("updatep", 24, lambda: sim.updatep())
    add("cmpt_sgs", 32, lambda: sim.cmpt_sgs())
    add("rk(mom+update+forcing)", 160, lambda: sim.rk(0))
    add("bounduvw", 0, lambda: sim.bounduvw(True, False))
    add("boundp", 0, lambda: sim.boundp(d.cbcpre, sim.bcp, "p"))
    add("chkdt", 32, lambda: sim.chkdt())
    if args.impdiff:
        add("substep(per-procedure)", 0, lambda: sim.substep(1))
    add("step/3", 0, lambda: sim.step())
    sim.close()


if __name__ == "__main__":
    main()
