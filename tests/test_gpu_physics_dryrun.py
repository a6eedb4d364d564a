"""Dry run of tests/test_zzz_gpu_physics.py on the CPU: the same test bodies, with cales_b200.driver.Simulation replaced by a
stand-in that drives the oracle through the same calls (init_flow/start/step/get/set_fields/cmpt_sgs).  Checks the HARNESS of
the GPU physics tests -- indexing, normalisations, run lengths, thresholds -- where no GPU is available; on the GPU box the
real library takes the stand-in's place."""
import dataclasses

import numpy as np
import pytest

import oracle.param as op
from oracle.main import Sim


class OracleBackedSimulation:
    def __init__(self, deck, **kw):
        names = {f.name for f in dataclasses.fields(op.Deck)}
        od = op.Deck()
        for f in dataclasses.fields(type(deck)):
            if f.name in names:
                v = getattr(deck, f.name)
                setattr(od, f.name, v.copy() if isinstance(v, np.ndarray) else v)
        if od.sgstype.strip() == "none":        # the long laminar runs: the solution does not depend on the periodic directions
            od.ng = (2, 4 if od.cbcpre[0, 1] == "P" else od.ng[1], od.ng[2])
        self.deck = deck
        self.s = Sim(od)
        self.shape = tuple(x + 2 for x in od.ng)

    time = property(lambda self: self.s.time)

    def init_flow(self):
        pass

    def start(self):
        pass

    def step(self, icheck=0):
        return self.s.step(icheck)

    def get(self, nm):
        return getattr(self.s, nm.upper())[0]

    def set_fields(self, **kw):
        for nm, a in kw.items():
            getattr(self.s, nm.upper())[0][...] = a

    def cmpt_sgs(self):
        self.s.cmpt_sgs()

    def close(self):
        pass


@pytest.fixture
def physics(monkeypatch):
    import cales_b200.driver as drv
    import test_zzz_gpu_physics as mod
    monkeypatch.setattr(drv, "Simulation", OracleBackedSimulation)
    return mod


@pytest.mark.parametrize("gr", [0., 2.])
def test_dryrun_channel(physics, gr):
    physics.test_laminar_channel_reaches_the_discrete_poiseuille_solution(gr)


def test_dryrun_duct(physics):
    physics.test_laminar_duct_converges_to_the_series_solution_at_second_order()


def test_dryrun_shear(physics):
    physics.test_smagorinsky_and_van_driest_closed_forms_on_a_linear_shear()
    physics.test_dynamic_smagorinsky_switches_off_in_a_laminar_shear()
