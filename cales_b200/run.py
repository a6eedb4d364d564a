"""Run control of `program cans` (src/main.f90:358-369, 395-407, 512-622): restart or initial condition, the stopping
criteria, the periodic stability / divergence check that also updates dt, the 0-D logs and the checkpoint cadence -- the
caller side of the hot path, so that the reference's own `input.nml` decks drive this library unchanged:

    python -m cales_b200.run [input.nml] [--datadir data/]         (N ranks: python -m torch.distributed.run ... -m cales_b200.run)

The loop only talks to a small simulation interface (`step(icheck)`, `dt`, `dt_cfl`, `time`, `istep`, `save`, `load`,
`init_flow`, `start`, optional `forcing_log()`), which `cales_b200.driver.Simulation` provides; tests drive it with a
stand-in on the CPU.  1-D/2-D/3-D field output (`out1d.h90` ...) is out of scope."""
import math
import os
import sys
import time as _time

import numpy as np

from . import checkpoint
from . import sanity

SMALL = float(np.finfo(np.float64).eps) * 10 ** (15 // 2)      # epsilon(1._rp)*10**(precision(1._rp)/2), param.f90:24


def out0d(fname, var, rank=0):
    """`out0d`, src/output.f90:18-37: append one record in format (*(E16.7e3)) (rank 0 only)."""
    if rank != 0:
        return
    with open(fname, "a") as f:
        f.write("".join(_e16_7e3(float(v)) for v in var) + "\n")


def _e16_7e3(v):
    """Fortran E16.7E3: 0.dddddddE+xxx right-justified in 16 columns."""
    if v == 0.0 or not math.isfinite(v):
        body = "0.0000000E+000" if v == 0.0 else ("NaN" if math.isnan(v) else ("Infinity" if v > 0 else "-Infinity"))
        return body.rjust(16)
    m, e = ("%.6E" % abs(v)).split("E")              # d.dddddd E+xx  ->  0.ddddddd E+(xx+1)
    digits = m.replace(".", "")
    return (("-" if v < 0 else "") + "0." + digits + "E%+04d" % (int(e) + 1)).rjust(16)


class RunResult:
    def __init__(self):
        self.nsteps = 0; self.kill = False; self.saved = []; self.checks = []


def run(sim, deck, datadir="data/", rank=0, wtime=_time.time, log=print, barrier=None):
    """The time loop of main.f90:405-622 around `sim.step`.  Returns a RunResult."""
    os.makedirs(datadir, exist_ok=True) if rank == 0 else None
    say = (lambda *a: log(*a)) if rank == 0 else (lambda *a: None)
    res = RunResult()
    twi = wtime()                                                     # main.f90:143
    if not deck.restart:                                              # main.f90:358-369
        sim.init_flow()
        say("*** Initial condition succesfully set ***")
    else:
        sim.load(os.path.join(datadir, "fld.bin"))
        say("*** Checkpoint loaded at time = %r time step = %d. ***" % (sim.time, sim.istep))
    sim.start()                                                       # ghost cells, eddy viscosity, first dt (main.f90:370-398)
    say("dt_cfl = %r dt = %r" % (sim.dt_cfl, sim.dt))
    say("*** Calculation loop starts now ***")
    savecounter = 0
    is_done = False
    while not is_done:
        t12 = wtime()
        checked = deck.icheck > 0 and (sim.istep + 1) % max(deck.icheck, 1) == 0
        out = sim.step(icheck=deck.icheck if checked else 0)          # istep += 1, time += dt, three substeps (+ chkdt, chkdiv)
        res.nsteps += 1
        say("Time step #%d Time = %r" % (sim.istep, sim.time))
        if deck.stop_type[0] and sim.istep >= deck.nstep:             # main.f90:512-521
            is_done = True
        if deck.stop_type[1] and sim.time >= deck.time_max:
            is_done = True
        if deck.stop_type[2] and (wtime() - twi) / 3600. >= deck.tw_max:
            is_done = True
        if checked:                                                   # main.f90:523-545
            say("Checking stability and divergence...")
            say("dt_cfl = %r dt = %r" % (sim.dt_cfl, sim.dt))
            if sim.dt_cfl < SMALL:
                say("ERROR: time step is too small."); say("Aborting...")
                is_done = True; res.kill = True
            divtot, divmax = out
            res.checks.append((sim.istep, sim.dt, divtot, divmax))
            say("Total divergence = %r | Maximum divergence = %r" % (divtot, divmax))
            if divmax > SMALL or math.isnan(divtot):
                say("ERROR: maximum divergence is too large."); say("Aborting...")
                is_done = True; res.kill = True
        if deck.iout0d > 0 and sim.istep % max(deck.iout0d, 1) == 0:  # main.f90:549-576
            out0d(os.path.join(datadir, "time.out"), [1. * sim.istep, sim.dt, sim.time], rank)
            flog = getattr(sim, "forcing_log", None)
            if flog is not None and (any(deck.is_forced) or any(abs(b) > 0. for b in deck.bforce)):
                dpdl, means = flog()
                out0d(os.path.join(datadir, "forcing.out"), [sim.time] + list(dpdl) + list(means), rank)
        if (deck.isave > 0 and sim.istep % max(deck.isave, 1) == 0) or (is_done and not res.kill):      # main.f90:590-611
            if deck.is_overwrite_save:
                filename = "fld.bin"
            else:
                filename = "fld_%07d.bin" % sim.istep
                if deck.nsaves_max > 0:
                    if savecounter >= deck.nsaves_max:
                        savecounter = 0
                    savecounter += 1
                    filename = "fld_%04d.bin" % savecounter
                    out0d(os.path.join(datadir, "log_checkpoints.out"), [1. * sim.istep, sim.time, 1. * savecounter], rank)
                if rank == 0:
                    checkpoint.gen_alias(datadir, filename, "fld.bin")
            sim.save(os.path.join(datadir, filename), barrier=barrier)
            res.saved.append(filename)
            say("*** Checkpoint saved at time = %r time step = %d. ***" % (sim.time, sim.istep))
        say("Elapsed time of the step: %r" % (wtime() - t12))
    if not res.kill:
        say("*** Fim ***")                                            # main.f90:633
    return res


def calc_dims(cbcvel, sgstype, ipencil, nproc):
    """dims for a deck that leaves them to the code (dims = 0,0).  The reference lets cuDecomp autotune, except that for
    'smag' `calc_dims` (src/initmpi.f90:230-259) first makes sure that at most two subdomains lie between two opposite walls
    (the van Driest damping needs a wall on every rank): the first decomposed direction with no-slip walls gets 2 ranks (1
    when nproc is odd), the other one the rest.  Without walls in a decomposed direction: z slabs (1 x nproc)."""
    t = {1: (1, 2), 2: (0, 2), 3: (0, 1)}[int(ipencil)]            # ipencil_t: the two decomposed directions (0-based)
    if sgstype.strip() == "smag" and nproc >= 2:
        for i, idir in enumerate(t):
            if cbcvel[0, idir, idir] + cbcvel[1, idir, idir] == "DD":
                d = [0, 0]
                d[i], d[1 - i] = (1, nproc) if nproc % 2 == 1 else (2, nproc // 2)
                return tuple(d)
    return (1, nproc)


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="run a CaLES input deck on libcales_b200 (one process per GPU)")
    ap.add_argument("deck", nargs="?", default="input.nml")
    ap.add_argument("--datadir", default="data/")
    args = ap.parse_args(argv)
    import ctypes as C
    import torch
    from . import lib as L
    from .deck import read_input
    from .driver import Simulation
    deck = read_input(args.deck)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if deck.dims[0] * deck.dims[1] == 0:                 # dims = 0,0: the reference autotunes (initmpi.f90:47-54); here z slabs,
        deck.dims = calc_dims(deck.cbcvel, deck.sgstype, deck.ipencil, world)   # after calc_dims for 'smag' (initmpi.f90:60-62)
    if deck.dims[0] * deck.dims[1] != world:
        raise SystemExit("dims=%s needs %d ranks, launched with %d" % (deck.dims, deck.dims[0] * deck.dims[1], world))
    torch.cuda.set_device(local)
    uid, barrier, mean_allreduce = None, None, None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = C.create_string_buffer(128)
            L.check(None, L.load().cales_get_unique_id(raw))
            buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        uid = bytes(buf.cpu().numpy().tobytes())
        barrier = dist.barrier

        def mean_allreduce(x):
            t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
            return float(t.item())
    sim = Simulation(deck, rank=rank, nranks=world, uid=uid, device=local)
    sanity.test_sanity_input(deck, sim.n, sim.is_bound, sim.h["zc"])         # main.f90:284-286
    if not deck.restart and mean_allreduce is not None:
        init = sim.init_flow
        sim.init_flow = lambda: init(mean_allreduce)
    res = run(sim, deck, args.datadir, rank=rank, barrier=barrier)
    sim.close()
    return 1 if res.kill else 0


if __name__ == "__main__":
    sys.exit(main())
