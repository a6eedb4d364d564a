"""Generates tests/golden/*.npz from the CPU oracle (the Fortran reference cannot be built or run in this image,
so the oracle -- pinned by tests/test_oracle_identities.py -- is the only executable statement of the algorithm).
Run from the repo root:  python tests/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle.param as op  # noqa: E402
from oracle.main import Sim  # noqa: E402

GOLDEN = {
    "channel_dsmag_16": ("deck_channel", dict(ng=(16, 12, 16), sgstype="dsmag"), 3),
    "channel_wm_smag_16": ("deck_channel", dict(ng=(16, 8, 12), sgstype="smag", wall_model=True, gtype=6, gr=0., l=(12.8, 4.8, 2.), visci=43500.), 3),
    "tgv_smag_16": ("deck_tgv", dict(ng=(16, 16, 16)), 3),
    "cavity_smag_12": ("deck_cavity", dict(ng=(12, 12, 12)), 3),
    "duct_wm_smag_12": ("deck_duct", dict(ng=(8, 12, 12), wall_model=True), 3),
}


def run(name):
    deck, kw, nsteps = GOLDEN[name]
    s = Sim(getattr(op, deck)(**kw))
    res = None
    for _ in range(nsteps):
        res = s.step(icheck=1)
    out = {nm.lower(): getattr(s, nm)[0][1:-1, 1:-1, 1:-1] for nm in ("U", "V", "W", "P", "VISCT")}
    out["p"] = out["p"] - out["p"].mean()
    out["dt"] = np.array(s.dt); out["divmax"] = np.array(res[1]); out["time"] = np.array(s.time)
    return out


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name in GOLDEN:
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **run(name))
        print("wrote", name)
