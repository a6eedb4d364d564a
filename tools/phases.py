#!/usr/bin/env python
"""Per-phase timing of one RK3 substep (any number of ranks; launch N>1 with torch.distributed.run).

Each phase of main.f90:418-506 is timed separately with CUDA events on the library's stream (device synchronised and
ranks barriered around every phase, max over ranks), so the sum is larger than a pipelined substep; the table shows where
the multi-GPU overhead sits (halo exchanges, transposes inside `solver`).

  python tools/phases.py [--nloc 256 256 256] [--iters 10]"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nloc", type=int, nargs=3, default=[256, 256, 256])
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--dims", type=int, nargs=2, default=None)
    args = ap.parse_args()
    from cales_b200 import lib as L
    from cales_b200.deck import deck_tgv, rkcoeff
    from cales_b200.driver import Simulation
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        lib = L.load()
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = C.create_string_buffer(128)
            L.check(None, lib.cales_get_unique_id(raw))
            buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        uid = bytes(buf.cpu().numpy().tobytes())
    dims = tuple(args.dims) if args.dims else (1, world)
    nl = args.nloc
    ng = (nl[0], nl[1] * dims[0], nl[2] * dims[1])
    deck = deck_tgv(ng=ng, dims=dims)
    sim = Simulation(deck, rank=rank, nranks=world, uid=uid, device=local)
    sim.init_flow(); sim.start()
    for _ in range(3):
        sim.step()
    d = deck
    dtrk = (rkcoeff[1][0] + rkcoeff[1][1]) * sim.dt

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timeit(fn):
        fn(); barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.
        for _ in range(args.iters):
            barrier()
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        t = torch.tensor([tot / args.iters], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    phases = [
        ("rk (mom+update)", lambda: sim.rk(1, want_f=False)),
        ("bulk_forcing", lambda: sim.bulk_forcing(None)),
        ("bounduvw", lambda: sim.bounduvw(True, False)),
        ("fillps", lambda: sim.fillps(1. / dtrk)),
        ("updt_rhs_b", lambda: sim.updt_rhs_b("ccc", d.cbcpre, sim.rhsbp, "pp")),
        ("solver", lambda: sim.solver(sim.poi, "pp")),
        ("boundp(pp)", lambda: sim.boundp(d.cbcpre, sim.bcp, "pp")),
        ("correc", lambda: sim.correc(0.0)),
        ("bounduvw(correc)", lambda: sim.bounduvw(True, True)),
        ("updatep", lambda: sim.updatep()),
        ("boundp(p)", lambda: sim.boundp(d.cbcpre, sim.bcp, "p")),
        ("cmpt_sgs", lambda: sim.cmpt_sgs()),
        ("boundp(visct)", lambda: sim.boundp(d.cbcsgs, sim.bcs, "visct")),
        ("substep (pipelined)", lambda: sim.substep(1)),
        ("step (pipelined)", lambda: sim.step()),
    ]
    rows = [(nm, timeit(fn)) for nm, fn in phases]
    if rank == 0:
        print("ranks %d dims %s grid %s" % (world, dims, ng))
        s = 0.
        for nm, ms in rows:
            print("%-22s %8.4f ms" % (nm, ms))
            if not nm.startswith(("substep", "step")):
                s += ms
        print("%-22s %8.4f ms" % ("sum of phases", s))
    sim.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
