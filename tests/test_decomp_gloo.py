"""world_size-2 (and 4) CPU test of the N>1 host logic over torch.distributed/gloo: each process takes its
decomposition, neighbour table and transpose exchange plan from the C-ABI library (the same integer code the
NCCL path runs), moves real data with gloo send/recv, and checks
  - the four pencil transposes x->y->z->y->x on a global-linear-index payload (cuDecomp transpose_test oracle),
  - a halo exchange against the oracle's serial emulation (bound.f90:619-696 semantics)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _plan(lib, L, ng, dims, rank, which):
    npeers = C.c_int()
    peers = np.zeros(16, dtype=np.int32); sb = np.zeros(96, dtype=np.int32); rb = np.zeros(96, dtype=np.int32)
    A = np.zeros(3, dtype=np.int32); B = np.zeros(3, dtype=np.int32)
    rc = lib.cales_transpose_plan(L._ia(ng), L._ia(dims), rank, which, C.byref(npeers), peers.ctypes.data_as(L.c_int_p),
                                  sb.ctypes.data_as(L.c_int_p), rb.ctypes.data_as(L.c_int_p), A.ctypes.data_as(L.c_int_p), B.ctypes.data_as(L.c_int_p))
    assert rc == 0
    P = npeers.value
    return P, peers[:P], sb[:6 * P].reshape(P, 6), rb[:6 * P].reshape(P, 6), tuple(int(x) for x in A), tuple(int(x) for x in B)


def _exchange(pairs):
    """pairs: list of (peer, send ndarray, recv ndarray).  Self copies are local; others isend/irecv."""
    reqs, bufs = [], []
    for peer, s, r in pairs:
        if peer == dist.get_rank():
            r[...] = s
            continue
        st = torch.from_numpy(np.ascontiguousarray(s.ravel(order="F")))
        rt = torch.empty(r.size, dtype=torch.float64)
        reqs.append(dist.isend(st, int(peer)))
        reqs.append(dist.irecv(rt, int(peer)))
        bufs.append((rt, r))
    for q in reqs:
        q.wait()
    for rt, r in bufs:
        r[...] = rt.numpy().reshape(r.shape, order="F")


def _worker(rank, world, dims, ng, port, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from cales_b200 import lib as L
        from oracle.decomp import World
        lib = L.load()
        cbc = "PPPPNN"
        cbcpre = np.array([[cbc[0], cbc[2], cbc[4]], [cbc[1], cbc[3], cbc[5]]])
        w = World(ng, dims, cbcpre, 1)
        me = w.ranks[rank]
        gidx = np.arange(np.prod(ng), dtype=np.float64).reshape(ng, order="F")
        # --- transposes -------------------------------------------------------------------------------
        cur = w.to_pencils(gidx, "x")[rank]
        for which, dst in ((0, "y"), (1, "z"), (2, "y"), (3, "x")):
            P, peers, sb, rb, A, B = _plan(lib, L, ng, dims, rank, which)
            assert cur.shape == A
            new = np.full(B, -1.0, order="F")
            pairs = []
            for qq in range(P):
                s = tuple(slice(sb[qq][d], sb[qq][d] + sb[qq][3 + d]) for d in range(3))
                t = tuple(slice(rb[qq][d], rb[qq][d] + rb[qq][3 + d]) for d in range(3))
                pairs.append((peers[qq], cur[s], new[t]))
            _exchange(pairs)
            assert np.array_equal(new, w.to_pencils(gidx, dst)[rank]), ("transpose", which)
            cur = new
        # --- halo exchange ------------------------------------------------------------------------------
        nb = np.zeros(6, dtype=np.int32); ib = np.zeros(6, dtype=np.int32)
        assert lib.cales_neighbours(L._ia(dims), 1, cbc.encode(), rank, nb.ctypes.data_as(L.c_int_p), ib.ctypes.data_as(L.c_int_p)) == 0
        nb = nb.reshape((2, 3), order="F")
        rng = np.random.default_rng(3)
        glob = rng.standard_normal(ng)
        fields = w.scatter(glob)
        for r_, a in zip(w.ranks, fields):        # distinct ghost values per rank, so stale ghosts would show
            a[0, :, :] = a[-1, :, :] = a[:, 0, :] = a[:, -1, :] = a[:, :, 0] = a[:, :, -1] = 100. + r_.id
        mine = fields[rank].copy(order="F")
        w.updthalo_all(fields)                    # the oracle's serial emulation
        n = me.n
        for idir in (1, 2):                       # X-pencils: y then z, full-extent faces
            sl = [slice(None)] * 3
            def pl(i):
                s = list(sl); s[idir] = i; return tuple(s)
            nb0, nb1 = int(nb[0, idir]), int(nb[1, idir])
            pairs = []
            if nb0 == rank and nb1 == rank:       # periodic self-neighbour: p(0)=p(n), p(n+1)=p(1)
                mine[pl(0)] = mine[pl(n[idir])]
                mine[pl(n[idir] + 1)] = mine[pl(1)]
                continue
            if nb1 >= 0:
                pairs.append((nb1, mine[pl(n[idir])].copy(), mine[pl(n[idir] + 1)]))
            if nb0 >= 0:
                pairs.append((nb0, mine[pl(1)].copy(), mine[pl(0)]))
            # order sends like the NCCL path: plane 1 -> nb0, plane n -> nb1; receives: upper ghost first
            if nb0 == nb1 and nb0 >= 0 and nb0 != rank:
                st0 = torch.from_numpy(np.ascontiguousarray(mine[pl(1)].ravel(order="F")))
                st1 = torch.from_numpy(np.ascontiguousarray(mine[pl(n[idir])].ravel(order="F")))
                r_hi = torch.empty(st0.numel(), dtype=torch.float64); r_lo = torch.empty(st0.numel(), dtype=torch.float64)
                reqs = [dist.isend(st0, nb0, tag=1), dist.isend(st1, nb1, tag=2), dist.irecv(r_hi, nb1, tag=1), dist.irecv(r_lo, nb0, tag=2)]
                for q_ in reqs:
                    q_.wait()
                shp = mine[pl(0)].shape
                mine[pl(n[idir] + 1)] = r_hi.numpy().reshape(shp, order="F")
                mine[pl(0)] = r_lo.numpy().reshape(shp, order="F")
            else:
                _exchange(pairs)
        assert np.array_equal(mine, fields[rank]), "halo"
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))


@pytest.mark.parametrize("dims,ng,port", [((1, 2), (16, 12, 10), 29601), ((2, 1), (16, 13, 10), 29602), ((2, 2), (18, 12, 14), 29603)])
def test_gloo_transposes_and_halos(dims, ng, port):
    world = dims[0] * dims[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, dims, ng, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
