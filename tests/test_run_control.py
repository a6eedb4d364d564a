"""Run control (cales_b200/run.py vs src/main.f90:358-369, 512-622) driven by a stand-in simulation on the CPU: stopping
criteria, check cadence and abort conditions, 0-D logs in the reference's record format, checkpoint naming."""
import os

import numpy as np
import pytest

from cales_b200 import run as R
from cales_b200.deck import Deck, read_input


class FakeSim:
    """Implements the interface `run` uses; divergence / dt_cfl scripted per step."""

    def __init__(self, dt=0.5, div=None, dtcfl=None):
        self.istep, self.time, self.dt, self.dt_cfl = 0, 0.0, dt, dt / 0.95
        self.div = div or {}
        self.dtcfl = dtcfl or {}
        self.calls = []

    def init_flow(self):
        self.calls.append("init")

    def load(self, fn):
        self.calls.append(("load", os.path.basename(fn)))
        self.istep, self.time = 40, 20.0

    def start(self):
        self.calls.append("start")

    def step(self, icheck=0):
        self.istep += 1
        self.time += self.dt
        if icheck > 0 and self.istep % icheck == 0:
            self.dt_cfl = self.dtcfl.get(self.istep, self.dt_cfl)
            return self.div.get(self.istep, (1e-18, 1e-14))
        return None

    def save(self, fn, barrier=None):
        self.calls.append(("save", os.path.basename(fn), self.istep))
        open(fn, "wb").write(b"x")

    def forcing_log(self):
        return (0.1, 0., 0.), (1., 0., 0.)


def deck(**kw):
    d = Deck()
    d.nstep, d.time_max, d.tw_max = 10, 1e9, 1e9
    d.stop_type = (True, False, False)
    d.icheck, d.iout0d, d.isave = 4, 5, 0
    d.is_forced = (True, False, False)
    for k, v in kw.items():
        setattr(d, k, v)
    return d


def test_stop_on_nstep_checks_logs_and_final_save(tmp_path):
    sim = FakeSim()
    lines = []
    res = R.run(sim, deck(), str(tmp_path) + "/", log=lines.append)
    assert res.nsteps == 10 and not res.kill and sim.calls[:2] == ["init", "start"]
    assert [c[0] for c in res.checks] == [4, 8]                       # mod(istep,icheck)==0
    assert res.saved == ["fld.bin"] and sim.calls[-1] == ("save", "fld.bin", 10)       # is_done and not kill
    assert lines[-1] == "*** Fim ***"
    rec = open(tmp_path / "time.out").read().splitlines()
    assert len(rec) == 2 and rec[0] == "  0.5000000E+001  0.5000000E+000  0.2500000E+001"     # istep=5, dt, time; (*(E16.7e3))
    f = open(tmp_path / "forcing.out").read().splitlines()
    assert len(f) == 2 and len(f[0]) == 7 * 16


def test_stop_on_time_and_wallclock(tmp_path):
    res = R.run(FakeSim(dt=0.5), deck(stop_type=(False, True, False), time_max=2.2), str(tmp_path) + "/", log=lambda *a: None)
    assert res.nsteps == 5                                            # time >= time_max after 5 steps of 0.5
    clock = iter(np.arange(0., 1e6, 1800.))                           # every wtime() call is half an hour later
    res = R.run(FakeSim(), deck(stop_type=(False, False, True), tw_max=2.0, nstep=10**9), str(tmp_path) + "/",
                wtime=lambda: float(next(clock)), log=lambda *a: None)
    assert 1 <= res.nsteps <= 3


def test_abort_on_divergence_or_small_dt_does_not_save(tmp_path):
    sim = FakeSim(div={4: (1e-3, 1e-3)})
    lines = []
    res = R.run(sim, deck(), str(tmp_path) + "/", log=lines.append)
    assert res.kill and res.nsteps == 4 and res.saved == [] and "ERROR: maximum divergence is too large." in lines
    assert "*** Fim ***" not in lines
    res = R.run(FakeSim(div={4: (float("nan"), 0.)}), deck(), str(tmp_path) + "/", log=lambda *a: None)
    assert res.kill and res.nsteps == 4                               # is_nan(divtot)
    res = R.run(FakeSim(dtcfl={8: 1e-12}), deck(), str(tmp_path) + "/", log=lambda *a: None)
    assert res.kill and res.nsteps == 8                               # dt_cfl < small
    assert R.SMALL == np.finfo(np.float64).eps * 10 ** 7              # param.f90:24


def test_checkpoint_cadence_and_names(tmp_path):
    d = deck(isave=3, is_overwrite_save=False, nsaves_max=2, nstep=9)
    sim = FakeSim()
    res = R.run(sim, d, str(tmp_path) + "/", log=lambda *a: None)
    assert res.saved == ["fld_0001.bin", "fld_0002.bin", "fld_0001.bin"]        # round robin over nsaves_max files
    assert os.path.realpath(tmp_path / "fld.bin") == os.path.realpath(tmp_path / "fld_0001.bin")     # gen_alias
    assert len(open(tmp_path / "log_checkpoints.out").read().splitlines()) == 3
    res = R.run(FakeSim(), deck(isave=4, is_overwrite_save=False, nsaves_max=0, nstep=8), str(tmp_path) + "/", log=lambda *a: None)
    assert res.saved == ["fld_0000004.bin", "fld_0000008.bin"]                  # fldnum = i7.7


def test_restart_reads_fld_bin(tmp_path):
    sim = FakeSim()
    res = R.run(sim, deck(restart=True, nstep=42), str(tmp_path) + "/", log=lambda *a: None)
    assert sim.calls[0] == ("load", "fld.bin") and sim.calls[1] == "start" and res.nsteps == 2 and sim.istep == 42


def test_deck_reader_run_control(tmp_path):
    p = tmp_path / "input.nml"
    p.write_text("""&dns
ng(1:3) = 16, 16, 16
l(1:3) = 1., 1., 1.
visci = 100.
cbcvel(0:1,1:3,1) = 6*'P'
cbcvel(0:1,1:3,2) = 6*'P'
cbcvel(0:1,1:3,3) = 6*'P'
cbcpre(0:1,1:3) = 6*'P'
nstep = 250, time_max = 12.5, tw_max = 0.2
stop_type(1:3) = F, T, T
restart = T, is_overwrite_save = F, nsaves_max = 3
icheck = 7, iout0d = 11, iout1d = 13, iout2d = 17, iout3d = 19, isave = 23
dims(1:2) = 2, 2
/
&les
sgstype = 'dsmag'
/
""")
    d = read_input(str(p))
    assert (d.nstep, d.time_max, d.tw_max) == (250, 12.5, 0.2) and d.stop_type == (False, True, True)
    assert d.restart and not d.is_overwrite_save and d.nsaves_max == 3
    assert (d.icheck, d.iout0d, d.iout1d, d.iout2d, d.iout3d, d.isave) == (7, 11, 13, 17, 19, 23) and d.dims == (2, 2)


@pytest.mark.parametrize("v,s", [(1.0, "  0.1000000E+001"), (-0.5, " -0.5000000E+000"), (123456.789, "  0.1234568E+006"), (0.0, "  0.0000000E+000"),
                                 (9.99999996e-5, "  0.1000000E-003")])
def test_e16_7e3(v, s):
    assert R._e16_7e3(v) == s
