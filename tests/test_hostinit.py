"""CPU tests: the product's host-side input synthesis (cales_b200/hostinit.py, deck.py) reproduces the oracle's
restatement of initgrid / initflow / initbc bit for bit, and the deck reader parses the reference decks."""
import glob
import os

import numpy as np
import pytest

import cales_b200.deck as pd
import oracle.param as op
from cales_b200 import hostinit
from oracle import bound as ob
from oracle.initflow import initflow as o_initflow
from oracle.initgrid import initgrid as o_initgrid


@pytest.mark.parametrize("gtype,gr", [(1, 0.), (1, 5.), (2, 2.), (3, 2.), (4, 1.5), (5, 0.), (6, 0.)])
@pytest.mark.parametrize("n", [32, 48, 192])
def test_initgrid(gtype, gr, n):
    a = hostinit.initgrid(gtype, n, gr, 2.0)
    b = o_initgrid(gtype, n, gr, 2.0)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert abs(a[3][n] - 2.0) < 1e-12 and np.all(a[1][1:-1] > 0)


CASES = [("deck_channel", dict(ng=(16, 12, 14), sgstype="dsmag")),
         ("deck_channel", dict(ng=(16, 12, 14), sgstype="smag", wall_model=True, gtype=6, gr=0., l=(12.8, 4.8, 2.), visci=43500.)),
         ("deck_tgv", dict(ng=(12, 10, 8))), ("deck_duct", dict(ng=(8, 12, 10), wall_model=True)), ("deck_cavity", dict(ng=(8, 8, 8)))]


@pytest.mark.parametrize("name,kw", CASES)
def test_initflow_and_initbc(name, kw):
    dd, od = getattr(pd, name)(**kw), getattr(op, name)(**kw)
    n = list(dd.ng); lo = [1, 1, 1]
    g = hostinit.initgrid(dd.gtype, n[2], dd.gr, dd.l[2])
    dzc, dzf, zc, zf = g
    a = hostinit.initflow(dd, lo, n, zc, zf, dzc, dzf)
    b = o_initflow(od, lo, n, zc, zf, dzc, dzf)
    for x, y in zip(a, b):
        if dd.inivel == "duc":      # 100-term cosh/cos series: vector vs scalar libm calls may differ in the last bit
            assert np.abs(x - y).max() <= 1e-14 * max(np.abs(y).max(), 1e-30)
        else:
            assert np.array_equal(x, y)
    is_bound = np.ones((2, 3), dtype=bool)
    cbc, bcu, bcv, bcw, bcp, bcs, iwm = hostinit.initbc(dd, n, is_bound, zc, dzc)
    r = ob.initbc(od, n, is_bound, zc, dzc)
    assert np.array_equal(cbc, r[0]) and np.array_equal(iwm, r[12])
    for x, y in zip((bcu, bcv, bcw, bcp, bcs), r[1:6]):
        for ax in "xyz":
            assert np.array_equal(x[ax], y[ax])


@pytest.mark.parametrize("inivel", ["cou", "iop", "zer", "uni", "pdc", "hdc", "hcp", "ant", "tgw", "log", "hcl", "tbl"])
@pytest.mark.parametrize("turb", [False, True])
def test_every_initial_condition_of_the_reference(inivel, turb):
    """All `inivel` values of initflow.f90:51-209 on the product side against the oracle's separate restatement, on a
    2 x 2 decomposition; the noisy ones ('log', 'hcl', 'tbl') use the documented stand-in for the compiler's random_number
    stream and must not depend on the decomposition."""
    from oracle import decomp as odec
    kw = dict(ng=(10, 8, 12))
    dd, od = pd.deck_channel(**kw), op.deck_channel(**kw)
    for d in (dd, od):
        d.inivel = inivel; d.is_wallturb = turb
        d.bforce = (0.3, 0., 0.)
        d.bcvel = d.bcvel.copy(); d.bcvel[0, 2, 0] = 1.0; d.bcvel[1, 2, 0] = -0.5
    ng = list(dd.ng)
    dzc, dzf, zc, zf = hostinit.initgrid(dd.gtype, ng[2], dd.gr, dd.l[2])
    glob = [np.zeros(ng, order="F") for _ in range(4)]
    dims = (2, 2)
    parts = []
    for r in range(4):
        lo, hi, n = odec.partition(ng, (1, 2, 3), dims, (r // 2, r % 2))
        ksl = slice(lo[2] - 1, hi[2] + 2)
        parts.append((lo, hi, n, ksl))
    def allsum_factory(fn, deck):
        # set_mean needs the global sum: first pass collects the partial sums, second pass uses their total
        tot = []
        for lo, hi, n, ksl in parts:
            fn(deck, lo, n, zc[ksl], zf[ksl], dzc[ksl], dzf[ksl], lambda x: tot.append(x) or x)
        return float(np.sum(tot))
    tp, to = allsum_factory(hostinit.initflow, dd), allsum_factory(o_initflow, od)
    for lo, hi, n, ksl in parts:
        a = hostinit.initflow(dd, lo, n, zc[ksl], zf[ksl], dzc[ksl], dzf[ksl], lambda x: tp)
        b = o_initflow(od, lo, n, zc[ksl], zf[ksl], dzc[ksl], dzf[ksl], lambda x: to)
        for q, (x, y) in enumerate(zip(a, b)):
            assert np.abs(x - y).max() <= 4e-15 * max(np.abs(y).max(), 1e-30), (inivel, q)
            glob[q][lo[0] - 1:hi[0], lo[1] - 1:hi[1], lo[2] - 1:hi[2]] = x[1:-1, 1:-1, 1:-1]
    # decomposition independence: one rank gives the same field (up to the rounding of the set_mean sum)
    one = hostinit.initflow(dd, [1, 1, 1], ng, zc, zf, dzc, dzf)
    for q in range(4):
        assert np.abs(one[q][1:-1, 1:-1, 1:-1] - glob[q]).max() <= 1e-13 * max(np.abs(glob[q]).max(), 1e-30), (inivel, q)
    if inivel in ("log", "hcl", "tbl"):
        noise = one[1][1:-1, 1:-1, 1:-1]                                   # without the vortex pair v carries only the noise
        if not turb:
            assert 0.03 < np.abs(noise).max() <= 0.05 and abs(noise.mean()) < 0.01          # 2(rn-.5)*0.05
    with pytest.raises(ValueError):
        hostinit.initflow(dd.copy(inivel="xyz"), [1, 1, 1], ng, zc, zf, dzc, dzf)


def test_deck_reader_on_reference_style_deck(tmp_path):
    txt = """&dns
ng(1:3) = 192, 72, 48
l(1:3) = 12.8, 4.8, 2.
gtype = 6, gr = 0.
cfl = 0.95, dtmax = 1.e5, dt_f = -1.
visci = 125000.
inivel = 'poi'
is_wallturb = T
nstep = 10000, time_max = 800., tw_max = 12
stop_type(1:3) = F, T, F
restart = F, is_overwrite_save = T, nsaves_max = 0
icheck = 10, iout0d = 10, iout1d = 100, iout2d = 1000, iout3d = 1000000, isave = 1000
cbcvel(0:1,1:3,1) = 'P','P',  'P','P',  'D','D'
cbcvel(0:1,1:3,2) = 'P','P',  'P','P',  'D','D'
cbcvel(0:1,1:3,3) = 'P','P',  'P','P',  'D','D'
cbcpre(0:1,1:3)   = 'P','P',  'P','P',  'N','N'
cbcsgs(0:1,1:3)   = 'P','P',  'P','P',  'D','D'
bcvel(0:1,1:3,1) =  0.,0.,   0.,0.,   0.,0.
bcvel(0:1,1:3,2) =  0.,0.,   0.,0.,   0.,0.
bcvel(0:1,1:3,3) =  0.,0.,   0.,0.,   0.,0.
bcpre(0:1,1:3)   =  0.,0.,   0.,0.,   0.,0.
bcsgs(0:1,1:3)   =  0.,0.,   0.,0.,   0.,0.
bforce(1:3) = 0., 0., 0.
is_forced(1:3) = T, F, F
velf(1:3) = 1., 0., 0.
dims(1:2) = 0, 0
/

&les
sgstype = 'smag'
lwm(0:1,1:3) = 0,0, 0,0, 1,1
hwm = 0.1
/
"""
    p = tmp_path / "input.nml"
    p.write_text(txt)
    for mod in (pd, op):
        d = mod.read_input(str(p))
        assert d.ng == (192, 72, 48) and d.l == (12.8, 4.8, 2.0) and d.gtype == 6 and d.visci == 125000.
        assert d.inivel == "poi" and d.is_wallturb is True and d.sgstype == "smag" and d.hwm == 0.1
        assert "".join(d.cbcvel[:, 2, 0]) == "DD" and "".join(d.cbcpre[:, 2]) == "NN" and "".join(d.cbcsgs[:, 0]) == "PP"
        assert list(d.lwm[:, 2]) == [1, 1] and list(d.lwm[:, 0]) == [0, 0]
        assert d.is_forced == (True, False, False) and d.velf == (1.0, 0.0, 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("inivel", ["zer", "uni", "cou", "poi", "iop", "hcp", "pdc", "hdc", "tgv", "tgw", "ant", "duc"])
@pytest.mark.parametrize("wallturb", [False, True])
def test_device_initflow_matches_the_host_path(inivel, wallturb, arith):
    """the deterministic initial conditions generated directly on the device (cales_initflow) against the host
    path (hostinit.initflow, itself held against the oracle above): sin/cos/exp/cosh and the set_mean reduction differ by
    round-off only (1e-13 relative to the largest value of the field)."""
    from conftest import need_gpu
    need_gpu()
    import cales_b200.deck as pd
    from cales_b200 import hostinit
    from cales_b200.driver import Simulation
    d = pd.deck_channel(ng=(20, 12, 16), sgstype="smag", gtype=1, gr=2.0, l=(3.0, 2.0, 1.5))
    d.inivel = inivel; d.is_wallturb = wallturb
    d.is_forced = (True, False, False); d.velf = (1.2, 0.0, 0.0); d.bforce = (0.5, 0.0, 0.0)
    d.bcvel[0, 2, 0] = 1.0; d.bcvel[1, 2, 0] = -0.5
    g = Simulation(d, graph=False)
    g.init_flow(device=True)
    ref = hostinit.initflow(d, g.lo, g.n, g.h["zc"], g.h["zf"], g.h["dzc"], g.h["dzf"])
    I = (slice(1, -1),) * 3
    for nm, r in zip(("u", "v", "w", "p"), ref):
        got = g.get(nm)
        scale = max(float(np.abs(r[I]).max()), 1e-3)
        assert np.abs(got[I] - r[I]).max() <= 1e-13 * scale, (nm, np.abs(got[I] - r[I]).max(), scale)
    g.close()


@pytest.mark.gpu
def test_device_initflow_refuses_the_noisy_cases():
    from conftest import need_gpu
    need_gpu()
    import cales_b200.deck as pd
    from cales_b200 import lib as L
    from cales_b200.driver import Simulation
    d = pd.deck_channel(ng=(16, 12, 16), sgstype="smag")
    d.inivel = "log"
    g = Simulation(d, graph=False)
    with pytest.raises(L.CalesError, match="host"):
        g.init_flow(device=True)
    g.init_flow()                                   # default: falls back to the host path for 'log'
    assert float(np.abs(g.get("u")).max()) > 0.
    g.close()
