"""Worker for the multi-GPU parity tests (launched by torch.distributed.run, one rank per GPU).
Runs the product on dims(1) x dims(2) ranks, gathers the fields on rank 0 and compares them with the
oracle's serial emulation of the same decomposition.  Usage: mgpu_worker.py CASE P_ROW P_COL NSTEPS"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    case, prow, pcol, nsteps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from cales_b200 import lib as L
    import cales_b200.deck as pd
    from cales_b200.driver import Simulation
    from test_gpu_step import CASES
    lib = L.load()
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = C.create_string_buffer(128)
        L.check(None, lib.cales_get_unique_id(raw))
        buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    uid = bytes(buf.cpu().numpy().tobytes())
    name, kw = CASES[case]
    kw = dict(kw); kw["dims"] = (prow, pcol)
    deck = getattr(pd, name)(**kw)
    sim = Simulation(deck, rank=rank, nranks=world, uid=uid, device=local)

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return t.item()
    sim.init_flow(mean_allreduce=allsum)
    sim.start()
    res = None
    for _ in range(nsteps):
        res = sim.step(icheck=1)
    # gather interiors on rank 0
    out = {}
    for nm in ("u", "v", "w", "p", "visct"):
        loc = sim.get(nm)[1:-1, 1:-1, 1:-1]
        objs = [None] * world if rank == 0 else None
        dist.gather_object((list(map(int, sim.lo)), list(map(int, sim.hi)), loc), objs, dst=0)
        if rank == 0:
            g = np.zeros(deck.ng, order="F")
            for lo, hi, a in objs:
                g[lo[0] - 1:hi[0], lo[1] - 1:hi[1], lo[2] - 1:hi[2]] = a
            out[nm] = g
    ok = True
    if rank == 0:
        import oracle.param as op
        from oracle.main import Sim
        o = Sim(getattr(op, name)(**kw))
        ro = None
        for _ in range(nsteps):
            ro = o.step(icheck=1)
        errs = {}
        for nm, on in (("u", "U"), ("v", "V"), ("w", "W"), ("p", "P"), ("visct", "VISCT")):
            a, b = out[nm], o.world.gather(getattr(o, on))
            if nm == "p":
                a = a - a.mean(); b = b - b.mean()
            errs[nm] = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
        ok = all(v <= 1e-10 for v in errs.values()) and abs(res[1] - ro[1]) < 1e-11 and abs(sim.dt - o.dt) <= 1e-10 * o.dt
        print(json.dumps({"case": case, "dims": [prow, pcol], "steps": nsteps, "errs": errs, "divmax": [res[1], ro[1]], "ok": ok}))
    sim.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
