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
