"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/cales_b200.h declares, fails
loudly without a device, and its pure-integer decomposition maps are BIT-EXACT against the restatement of
2decomp `distribute`/`partition` + initmpi (oracle/decomp.py) -- exhaustively over uneven splits."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "cales_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(cales_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 35
    for nm in sorted(names):
        assert hasattr(lib, nm), "symbol %s declared in the header is not exported" % nm


def test_python_signatures_cover_header(lib):
    from cales_b200 import lib as L
    hdr = open(os.path.join(ROOT, "include", "cales_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(cales_[a-z0-9_]+)\s*\(", hdr))
    assert names == set(L.SIGNATURES), names ^ set(L.SIGNATURES)


def test_no_device_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from cales_b200 import lib as L
    ctx = C.c_void_p()
    rc = lib.cales_init(C.byref(ctx), L._ia([8, 8, 8]), L._ia([1, 1]), 1, b"PPPPPP", 0, 1, None, 0, None, 0)
    assert rc != 0 and b"no CPU fallback" in lib.cales_last_error(None)
    from cales_b200.deck import deck_tgv
    from cales_b200.driver import Simulation
    with pytest.raises(L.CalesError):
        Simulation(deck_tgv(ng=(8, 8, 8)))


def test_product_never_imports_oracle():
    for dp, _, fs in os.walk(os.path.join(ROOT, "cales_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


@pytest.mark.parametrize("n", [1, 7, 64, 97, 192, 1000, 1024])
def test_distribute_bitexact(lib, n):
    from cales_b200 import lib as L
    from oracle.decomp import distribute
    for proc in range(1, min(n, 17) + 1):
        st, en, sz = (np.zeros(proc, dtype=np.int32) for _ in range(3))
        assert lib.cales_distribute(n, proc, st.ctypes.data_as(L.c_int_p), en.ctypes.data_as(L.c_int_p), sz.ctypes.data_as(L.c_int_p)) == 0
        rst, ren, rsz = distribute(n, proc)
        assert list(st) == rst and list(en) == ren and list(sz) == rsz
        assert sum(sz) == n and st[0] == 1 and en[-1] == n


@pytest.mark.parametrize("ng", [(64, 64, 64), (30, 17, 23), (512, 256, 192), (1024, 512, 512)])
@pytest.mark.parametrize("dims", [(1, 1), (1, 2), (2, 1), (2, 2), (1, 8), (8, 1), (2, 4), (4, 2), (3, 5)])
@pytest.mark.parametrize("cbc", ["PPPPNN", "PPPPPP", "NNNNNN", "PPNNNN"])
def test_pencils_and_neighbours_bitexact(lib, ng, dims, cbc):
    from cales_b200 import lib as L
    from oracle.decomp import World
    cbcpre = np.array([[cbc[0], cbc[2], cbc[4]], [cbc[1], cbc[3], cbc[5]]])
    for ipencil in (1, 2, 3):
        w = World(ng, dims, cbcpre, ipencil)
        for r in w.ranks:
            for axis, (st, en, sz) in enumerate(((r.xstart, r.xend, r.xsize), (r.ystart, r.yend, r.ysize), (r.zstart, r.zend, r.zsize)), 1):
                lo, hi, s = (np.zeros(3, dtype=np.int32) for _ in range(3))
                assert lib.cales_pencil(L._ia(ng), L._ia(dims), r.id, axis, lo.ctypes.data_as(L.c_int_p), hi.ctypes.data_as(L.c_int_p),
                                        s.ctypes.data_as(L.c_int_p)) == 0
                assert list(lo) == st and list(hi) == en and list(s) == sz
            nb, ib = np.zeros(6, dtype=np.int32), np.zeros(6, dtype=np.int32)
            assert lib.cales_neighbours(L._ia(dims), ipencil, cbc.encode(), r.id, nb.ctypes.data_as(L.c_int_p), ib.ctypes.data_as(L.c_int_p)) == 0
            assert np.array_equal(nb.reshape((2, 3), order="F"), r.nb)
            assert np.array_equal(ib.reshape((2, 3), order="F").astype(bool), r.is_bound)


def plan(lib, ng, dims, rank, which):
    from cales_b200 import lib as L
    npeers = C.c_int()
    peers = np.zeros(16, dtype=np.int32); sb = np.zeros(96, dtype=np.int32); rb = np.zeros(96, dtype=np.int32)
    A = np.zeros(3, dtype=np.int32); B = np.zeros(3, dtype=np.int32)
    assert lib.cales_transpose_plan(L._ia(ng), L._ia(dims), rank, which, C.byref(npeers), peers.ctypes.data_as(L.c_int_p),
                                    sb.ctypes.data_as(L.c_int_p), rb.ctypes.data_as(L.c_int_p), A.ctypes.data_as(L.c_int_p),
                                    B.ctypes.data_as(L.c_int_p)) == 0
    P = npeers.value
    return P, peers[:P].copy(), sb[:6 * P].reshape(P, 6).copy(), rb[:6 * P].reshape(P, 6).copy(), tuple(A), tuple(B)


@pytest.mark.parametrize("ng", [(16, 12, 10), (30, 17, 23), (64, 48, 40)])
@pytest.mark.parametrize("dims", [(1, 2), (2, 1), (2, 2), (2, 4), (4, 2), (3, 5), (1, 8)])
def test_transpose_plan_moves_global_index_exactly(lib, ng, dims):
    """The cuDecomp transpose_test oracle (tests/cc/transpose_test.cc:116-160): every element carries its global
    linear index; after x->y->z->y->x every pencil must hold exactly the indices of its own sub-box."""
    from oracle.decomp import World
    cbcpre = np.full((2, 3), "P")
    w = World(ng, dims, cbcpre, 1)
    gidx = np.arange(np.prod(ng), dtype=np.float64).reshape(ng, order="F")
    cur = w.to_pencils(gidx, "x")
    for which, dstname in ((0, "y"), (1, "z"), (2, "y"), (3, "x")):
        plans = [plan(lib, ng, dims, r.id, which) for r in w.ranks]
        new = [np.full(pl[5], -1.0, order="F") for pl in plans]
        for r in w.ranks:
            P, peers, sb, rb, A, B = plans[r.id]
            assert cur[r.id].shape == A
            for q in range(P):
                # what I send to peers[q] is what peers[q] receives from me: its recv box for the slot of my row/col index
                peer = int(peers[q])
                Pq, peers_q, sbq, rbq, Aq, Bq = plans[peer]
                slot = list(peers_q).index(r.id)
                s = tuple(slice(sb[q][d], sb[q][d] + sb[q][3 + d]) for d in range(3))
                t = tuple(slice(rbq[slot][d], rbq[slot][d] + rbq[slot][3 + d]) for d in range(3))
                assert cur[r.id][s].shape == new[peer][t].shape
                new[peer][t] = cur[r.id][s]
        ref = w.to_pencils(gidx, dstname)
        for a, b in zip(new, ref):
            assert np.array_equal(a, b)
        cur = new


def test_step_args_layout_matches_the_library(lib):
    """the ctypes mirror of cales_step_args has the layout the library was compiled with (both variants)"""
    import ctypes as C
    from cales_b200 import lib as L
    for arith in ("fma", "strict"):
        out = (C.c_long * 4)()
        assert L.load(arith).cales_step_args_layout(out) == 0
        S = L.StepArgs
        assert list(out) == [C.sizeof(S), S.cbcvel.offset, S.zc.offset, S.visct.offset]
