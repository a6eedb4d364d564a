"""The oracle evaluates its per-cell stencils, transforms and column solves slab by slab on a thread pool when the grid is
BASELINE-sized (oracle/mom.py: run_slabs; the 512x256x192 parity case would otherwise cost minutes per step).  Cutting the
level range of per-cell work, or the batch of independent lines / columns, must not change a single bit: the same short
runs with the slab paths forced on (tiny slabs, every grid size) and off."""
import sys

import numpy as np
import pytest

import oracle.param as op
from oracle.main import Sim

CASES = {"channel_wm_smag": ("deck_channel", dict(ng=(40, 24, 30), sgstype="smag", wall_model=True, gtype=6, gr=0., l=(12.8, 4.8, 2.), visci=43500.), None),
         "channel_dsmag_odd": ("deck_channel", dict(ng=(30, 18, 21), sgstype="dsmag"), None),
         "tgv_smag": ("deck_tgv", dict(ng=(24, 24, 24)), None),
         "cavity_smag": ("deck_cavity", dict(ng=(24, 20, 16), sgstype="smag"), None),
         "channel_impdiff": ("deck_channel", dict(ng=(24, 16, 20), sgstype="smag"), "3d"),
         "channel_impdiff_1d": ("deck_channel", dict(ng=(24, 16, 20), sgstype="smag"), "1d")}


def run(case):
    name, kw, imp = CASES[case]
    d = getattr(op, name)(**kw)
    if imp:
        d.impdiff = True
        d.impdiff_1d = imp == "1d"
    o = Sim(d)
    for _ in range(2):
        o.step(icheck=1)
    return {k: getattr(o, k)[0].copy() for k in ("U", "V", "W", "P", "VISCT")}, o.dt


@pytest.mark.parametrize("case", list(CASES))
def test_slab_paths_give_identical_bits(case, monkeypatch):
    ref, dt = run(case)
    for m in list(sys.modules.values()):
        if getattr(m, "__name__", "").startswith("oracle."):
            if hasattr(m, "SLAB_MIN_CELLS"):
                monkeypatch.setattr(m, "SLAB_MIN_CELLS", 0)
            if hasattr(m, "SLAB_CELLS"):
                monkeypatch.setattr(m, "SLAB_CELLS", 2000)
    got, dt2 = run(case)
    assert dt2 == dt
    for k in ref:
        assert np.array_equal(got[k], ref[k], equal_nan=True), k


def test_slab_partition_covers_every_level_once():
    from oracle.mom import run_slabs
    for n3, nplane in ((1, 10), (7, 3000), (64, 4096), (192, 131072), (5, 10 ** 7)):
        seen = np.zeros(n3, dtype=int)

        def job(k0, nb):
            seen[k0:k0 + nb] += 1
        run_slabs(n3, nplane, job)
        assert (seen == 1).all()
