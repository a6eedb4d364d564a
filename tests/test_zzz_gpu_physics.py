"""The CUDA path against closed-form physics, without the oracle in between (the oracle is unpinned by the reference, DESIGN.md
section 2; tests/test_oracle_physics.py holds the oracle to the same answers on CPU): laminar channel -> dense solution of the
discrete Poiseuille problem, laminar duct -> series solution at second order, Smagorinsky / van Driest closed forms on a
linear shear, the dynamic model switching off in a laminar shear.  Both library variants."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
from conftest import need_gpu  # noqa: E402
from test_oracle_physics import _duct_series, discrete_poiseuille  # noqa: E402


@pytest.fixture(autouse=True)
def _gpu(arith):
    need_gpu()


def _steady(g, t_end):
    while g.time < t_end:
        g.step()


@pytest.mark.parametrize("gr", [0., 2.])
def test_laminar_channel_reaches_the_discrete_poiseuille_solution(gr):
    import cales_b200.deck as pd
    from cales_b200.driver import Simulation
    from oracle.initgrid import initgrid
    n3 = 16
    d = pd.deck_channel(ng=(32, 24, n3), sgstype="none", visci=2., gr=gr)
    d.is_wallturb = False
    g = Simulation(d)
    g.init_flow(); g.start()
    _steady(g, 3.0)
    dzc, dzf, zc, zf = initgrid(d.gtype, n3, d.gr, d.l[2])
    u_ref, G = discrete_poiseuille(dzc, dzf, d.visc, d.l[2])
    u = g.get("u")[1:-1, 1:-1, 1:-1]
    assert np.abs(u - u_ref[None, None, :]).max() < 1e-8, gr
    assert np.abs(g.get("v")).max() < 1e-11 and np.abs(g.get("w")).max() < 1e-11
    g.close()


def test_laminar_duct_converges_to_the_series_solution_at_second_order():
    import cales_b200.deck as pd
    from cales_b200.driver import Simulation
    errs = []
    for n in (16, 32):
        d = pd.deck_duct(ng=(16, n, n), sgstype="none", visci=1.)
        g = Simulation(d)
        g.init_flow(); g.start()
        _steady(g, 1.5)
        yc = (np.arange(1, n + 1) - .5) * d.dl[1] - 1.
        zc = (np.arange(1, n + 1) - .5) * d.dl[2] - 1.
        ref = _duct_series(yc, zc)
        ref = ref / ref.mean()
        u = g.get("u")[1:-1, 1:-1, 1:-1]
        assert np.abs(u - u[:1]).max() < 1e-11               # x-invariant
        errs.append(np.abs(u[0] - ref).max())
        g.close()
    assert errs[1] < 5e-3 and 3. < errs[0] / errs[1] < 5.5, errs


def _shear_fields(g, S, zc, wall_ghosts):
    shp = g.shape
    u = np.asfortranarray(np.broadcast_to((S * zc)[None, None, :], shp).copy())
    if wall_ghosts:
        u[:, :, 0] = -u[:, :, 1]
        u[:, :, -1] = -u[:, :, -2]
    z = np.zeros(shp, order="F")
    g.set_fields(u=u, v=z, w=z)


def test_smagorinsky_and_van_driest_closed_forms_on_a_linear_shear():
    import cales_b200.deck as pd
    from cales_b200.driver import Simulation
    from oracle.initgrid import initgrid
    from oracle.param import c_smag
    # no walls: nu_t = (c_s Delta)^2 |S|
    S = 3.7
    d = pd.deck_tgv(ng=(32, 24, 20), sgstype="smag")
    g = Simulation(d)
    g.init_flow(); g.start()
    dzc, dzf, zc, zf = initgrid(d.gtype, d.ng[2], d.gr, d.l[2])
    _shear_fields(g, S, zc, False)
    g.cmpt_sgs()
    expect = (c_smag * (d.dl[0] * d.dl[1] * dzf[1:-1]) ** (1. / 3.)) ** 2 * abs(S)
    v = g.get("visct")[1:-1, 1:-1, 1:-1]
    assert np.abs(v - expect[None, None, :]).max() < 1e-13 * expect.max()
    g.close()
    # channel walls: van Driest damping with the wall shear nu S of the lower wall
    S, n3 = 40., 16
    d = pd.deck_channel(ng=(32, 24, n3), sgstype="smag", visci=500., gr=0.)
    d.is_wallturb = False
    g = Simulation(d)
    g.init_flow(); g.start()
    dzc, dzf, zc, zf = initgrid(d.gtype, n3, d.gr, d.l[2])
    _shear_fields(g, S, zc, True)
    g.cmpt_sgs()
    delta = (d.dl[0] * d.dl[1] * dzf[1:-1]) ** (1. / 3.)
    dplus = zc[1:-1] * np.sqrt(d.visc * S) / d.visc
    expect = (c_smag * delta * (1. - np.exp(-dplus / 25.))) ** 2 * S
    v = g.get("visct")[1:-1, 1:-1, 1:-1]
    K = slice(1, n3 // 2)
    assert np.abs(v[:, :, K] - expect[None, None, K]).max() < 1e-12 * expect[K].max()
    g.close()


def test_dynamic_smagorinsky_switches_off_in_a_laminar_shear():
    import cales_b200.deck as pd
    from cales_b200.driver import Simulation
    from oracle.initgrid import initgrid
    from oracle.param import c_smag
    S, n3 = 5., 16
    d = pd.deck_channel(ng=(32, 24, n3), sgstype="dsmag", visci=100., gr=0.)
    d.is_wallturb = False
    g = Simulation(d)
    g.init_flow(); g.start()
    dzc, dzf, zc, zf = initgrid(d.gtype, n3, d.gr, d.l[2])
    _shear_fields(g, S, zc, True)
    g.cmpt_sgs()
    v = g.get("visct")[1:-1, 1:-1, 1:-1]
    smag = (c_smag * (d.dl[0] * d.dl[1] * d.dl[2]) ** (1. / 3.)) ** 2 * S
    assert np.abs(v[:, :, 3:n3 - 3]).max() < 1e-10 * smag
    g.close()
