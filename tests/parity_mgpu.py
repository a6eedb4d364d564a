"""Multi-GPU parity checks shared by tests/mgpu_worker.py and bench.py (which runs them on the live process group before
timing whenever WORLD_SIZE > 1, so every scaling number carries an oracle comparison on the same decomposition).

  case_vs_oracle   the product on dims(1) x dims(2) ranks, fields gathered on rank 0 and compared with the oracle's serial
                   emulation of the same decomposition (TEST INFRASTRUCTURE: the oracle is the checker, never the thing timed)
  transpose_round_trip   the four device transposes x->y->z->y->x on a global-linear-index payload, every stage compared
                   exactly with the owner's sub-box -- the convention of dependencies/cuDecomp/tests/cc/transpose_test.cc:116-160
                   (and 2decomp test2d.f90:110-139); both the NCCL pack/send/unpack path and the peer-memory path."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "tests")):
    if p_ not in sys.path:
        sys.path.insert(0, p_)

CASES = {
    "channel_dsmag": ("deck_channel", dict(ng=(32, 24, 32), sgstype="dsmag")),
    "channel_smag": ("deck_channel", dict(ng=(32, 24, 32), sgstype="smag")),
    "channel_wm_smag": ("deck_channel", dict(ng=(32, 16, 24), sgstype="smag", wall_model=True, gtype=6, gr=0., l=(12.8, 4.8, 2.), visci=43500.)),
    "channel_wm_dsmag": ("deck_channel", dict(ng=(32, 16, 24), sgstype="dsmag", wall_model=True, gtype=6, gr=0., l=(12.8, 4.8, 2.), visci=43500.)),
    "tgv_smag": ("deck_tgv", dict(ng=(32, 32, 32))),
    "duct_smag": ("deck_duct", dict(ng=(16, 24, 24))),
    "duct_wm_smag": ("deck_duct", dict(ng=(16, 24, 24), wall_model=True)),
    "cavity_smag": ("deck_cavity", dict(ng=(24, 24, 24))),
}


def nccl_uid(lib, L, rank):
    """rank 0 draws the NCCL unique id through the C ABI, the host broadcasts it (cuDecomp cudecomp.cc:66-80)."""
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = C.create_string_buffer(128)
        L.check(None, lib.cales_get_unique_id(raw), lib)
        buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())


def allsum(x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t)
    return t.item()


def gather_field(sim, nm, rank, world):
    loc = sim.get(nm)[1:-1, 1:-1, 1:-1]
    objs = [None] * world if rank == 0 else None
    dist.gather_object((list(map(int, sim.lo)), list(map(int, sim.hi)), loc), objs, dst=0)
    if rank != 0:
        return None
    g = np.zeros(sim.deck.ng, order="F")
    for lo, hi, a in objs:
        g[lo[0] - 1:hi[0], lo[1] - 1:hi[1], lo[2] - 1:hi[2]] = a
    return g


def case_vs_oracle(name, kw, dims, nsteps, rank, world, local, uid, arith=None, tol=1e-10, impdiff=None):
    """Returns {"case", "dims", "steps", "errs", "divmax", "ok"} on rank 0, {"ok": ...} elsewhere (ok is broadcast)."""
    import cales_b200.deck as pd
    from cales_b200.driver import Simulation
    kw = dict(kw); kw["dims"] = tuple(dims)
    deck = getattr(pd, name)(**kw)
    if impdiff:
        deck.impdiff = True; deck.impdiff_1d = impdiff == "1d"
    sim = Simulation(deck, rank=rank, nranks=world, uid=uid, device=local, arith=arith)
    sim.init_flow(mean_allreduce=allsum)
    sim.start()
    res = None
    for _ in range(nsteps):
        res = sim.step(icheck=1)
    out = {nm: gather_field(sim, nm, rank, world) for nm in ("u", "v", "w", "p", "visct")}
    exch = sim.lib.cales_solver_exchange(sim.ctx).decode()
    rec = {"ok": True}
    if rank == 0:
        import oracle.param as op
        from oracle.main import Sim
        od = getattr(op, name)(**kw)
        if impdiff:
            od.impdiff = True; od.impdiff_1d = impdiff == "1d"
        o = Sim(od)
        ro = None
        for _ in range(nsteps):
            ro = o.step(icheck=1)
        errs = {}
        ref = {nm: o.world.gather(getattr(o, on)) for nm, on in (("u", "U"), ("v", "V"), ("w", "W"), ("p", "P"), ("visct", "VISCT"))}
        vscale = max(float(np.abs(ref[k]).max()) for k in "uvw")        # velocities: relative to the largest component
        for nm in ("u", "v", "w", "p", "visct"):
            a, b = out[nm], ref[nm]
            if nm == "p":
                a = a - a.mean(); b = b - b.mean()
            errs[nm] = float(np.abs(a - b).max() / max(vscale if nm in "uvw" else max(np.abs(b).max(), vscale ** 2) if nm == "p" else np.abs(b).max(), 1e-300))
        ok = all(v <= tol for v in errs.values()) and abs(res[1] - ro[1]) < 1e-11 and abs(sim.dt - o.dt) <= 1e-10 * o.dt
        rec = {"case": name + ":" + kw.get("sgstype", "smag") + (":impdiff_" + impdiff if impdiff else ""), "ng": list(deck.ng), "dims": list(dims),
               "steps": nsteps, "solver_exchange": exch, "errs": errs, "divmax": [res[1], ro[1]], "tol": tol, "ok": bool(ok)}
    sim.close()
    flag = torch.tensor([1 if rec["ok"] else 0], device="cuda")
    dist.broadcast(flag, 0)
    rec["ok"] = bool(flag.item())
    return rec


def transpose_round_trip(ng, dims, rank, world, local, uid, arith=None):
    """x->y->z->y->x on the device with payload = global linear index; exact comparison at every stage, for the NCCL path
    (plain device destination) and the peer-memory path (destination from cales_peer_alloc)."""
    from cales_b200 import lib as L
    lib = L.load(arith)
    ctx = C.c_void_p()
    stream = torch.cuda.current_stream()
    L.check(None, lib.cales_init(C.byref(ctx), L._ia(ng), L._ia(dims), 1, b"PPPPPP", rank, world, uid, local, C.c_void_p(stream.cuda_stream), 0), lib)
    gidx = np.arange(np.prod(ng), dtype=np.float64).reshape(ng, order="F")

    def pencil(ax):
        lo = np.zeros(3, dtype=np.int32); hi = np.zeros(3, dtype=np.int32); sz = np.zeros(3, dtype=np.int32)
        assert lib.cales_pencil(L._ia(ng), L._ia(dims), rank, ax, lo.ctypes.data_as(L.c_int_p), hi.ctypes.data_as(L.c_int_p), sz.ctypes.data_as(L.c_int_p)) == 0
        return np.asfortranarray(gidx[lo[0] - 1:hi[0], lo[1] - 1:hi[1], lo[2] - 1:hi[2]])
    nmax = 0
    for r in range(world):
        for ax in (1, 2, 3):
            sz = np.zeros(3, dtype=np.int32); lo = np.zeros(3, dtype=np.int32); hi = np.zeros(3, dtype=np.int32)
            lib.cales_pencil(L._ia(ng), L._ia(dims), r, ax, lo.ctypes.data_as(L.c_int_p), hi.ctypes.data_as(L.c_int_p), sz.ctypes.data_as(L.c_int_p))
            nmax = max(nmax, int(np.prod(sz)))
    bad = []
    for path in ("nccl", "peer"):
        if path == "peer":
            bufs = []
            for q in range(2):
                ptr = C.c_void_p()
                L.check(ctx, lib.cales_peer_alloc(ctx, b"trt%d" % q, nmax * 8, C.byref(ptr)), lib)
                bufs.append(ptr.value)
            hold = None
        else:
            hold = [torch.zeros(nmax, dtype=torch.float64, device="cuda") for _ in range(2)]
            bufs = [t.data_ptr() for t in hold]
        cur = pencil(1)
        src = torch.from_numpy(np.ascontiguousarray(cur.ravel(order="F"))).cuda()
        sp = src.data_ptr()
        for q, (which, ax) in enumerate(((0, 2), (1, 3), (2, 2), (3, 1))):
            dp = bufs[q % 2]
            L.check(ctx, lib.cales_transpose(ctx, which, C.c_void_p(sp), C.c_void_p(dp)), lib)
            L.check(ctx, lib.cales_stream_synchronize(ctx), lib)
            want = pencil(ax)
            got = _as_tensor(dp, want.size).cpu().numpy().reshape(want.shape, order="F")   # view of the raw device pointer
            if not np.array_equal(got, want):
                bad.append((path, which))
            sp = dp
        del hold
    L.check(ctx, lib.cales_finalize(ctx), lib)
    flag = torch.tensor([len(bad)], device="cuda")
    dist.all_reduce(flag)
    return {"ng": list(ng), "dims": list(dims), "paths": ["nccl", "peer"], "payload": "global linear index, exact", "ok": flag.item() == 0, "bad": bad}


def _as_tensor(ptr, count):
    """float64 CUDA tensor viewing `count` doubles at device address `ptr` (no copy)."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(h, device="cuda")
