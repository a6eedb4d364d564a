"""The product-side namelist reader (cales_b200/deck.py read_input = src/param.f90:88-152), written independently of the
oracle's reader: legal Fortran namelist forms (inline '/', repeat counts, values over several lines, comments), loud
failure where the Fortran runtime would stop, agreement with the oracle's reader on every example deck of the reference
(when /root/reference is mounted), and the derived constants against hand-computed values (param.f90:154-158)."""
import glob
import os

import numpy as np
import pytest

import cales_b200.deck as pd

GOOD = """! a comment line
&dns
ng(1:3) = 32, 16,
          8                      ! continued on the next line
l = 1. 2. 3., visci = 100. , gtype = 1, gr = 0.
cbcvel(:,:,1) = 6*'P', cbcvel(0:1,1:3,2)= 'P','P', 'P','P', 'D','D', cbcvel(:,:,3)=6*'P'
cbcpre = 4*'P' 2*'N', bforce = 3*0., is_forced = T, F, .false., velf = 1.d0, 0., 0.
inivel = 'poi' dims = 2 1 /
&les sgstype='smag', lwm = 4*0, 1, 1, hwm = 1.e-1 /
"""


def write(tmp_path, txt):
    p = tmp_path / "input.nml"
    p.write_text(txt)
    return str(p)


def test_legal_namelist_forms(tmp_path):
    d = pd.read_input(write(tmp_path, GOOD))
    assert d.ng == (32, 16, 8) and d.l == (1.0, 2.0, 3.0) and d.visci == 100.0 and d.dims == (2, 1)
    assert list(d.cbcvel[:, 2, 1]) == ["D", "D"] and list(d.cbcvel[:, 2, 0]) == ["P", "P"]
    assert list(d.cbcpre[:, 2]) == ["N", "N"] and list(d.cbcpre[:, 0]) == ["P", "P"]
    assert d.bforce == (0.0, 0.0, 0.0) and d.is_forced == (True, False, False) and d.velf == (1.0, 0.0, 0.0)
    assert d.sgstype == "smag" and d.hwm == 0.1 and list(d.lwm[:, 2]) == [1, 1] and d.lwm[:, :2].sum() == 0
    # derived constants, param.f90:154-158, hand-computed
    assert np.array_equal(d.dl, np.array([1. / 32., 2. / 16., 3. / 8.]))
    assert np.array_equal(d.dli, np.array([1. / 32., 2. / 16., 3. / 8.]) ** (-1)) and d.visc == 100.0 ** (-1)


def test_backslash_terminator(tmp_path):
    d = pd.read_input(write(tmp_path, GOOD.replace("hwm = 1.e-1 /", "hwm = 1.e-1\n\\\n")))
    assert d.sgstype == "smag" and d.hwm == 0.1


@pytest.mark.parametrize("bad,what", [
    (GOOD.replace("dims = 2 1 /", "dims = 2 1"), "not terminated"),
    (GOOD.replace("gr = 0.", "gr = 0., foo = 3"), "unknown entry"),
    (GOOD.split("&les")[1].join(["&les", ""]), "no &dns"),
    (GOOD.replace("ng(1:3) = 32, 16,\n          8 ", "ng = 32, 16 "), "expected 3"),
    (GOOD.replace("visci = 100. ,", ""), "required entries missing"),
    (GOOD.replace("visci = 100.", "visci = 'x'"), "not of type"),
    (GOOD.replace("4*'P' 2*'N'", "5*'P'"), "expected 6"),
])
def test_rejected_decks(tmp_path, bad, what):
    with pytest.raises(pd.DeckError, match=what):
        pd.read_input(write(tmp_path, bad))


@pytest.mark.skipif(not os.path.isdir("/root/reference/examples"), reason="reference tree not mounted")
def test_agrees_with_the_oracle_reader_on_every_reference_deck():
    import oracle.param as op
    decks = sorted(glob.glob("/root/reference/examples/**/input.nml", recursive=True))
    assert len(decks) >= 20
    for f in decks:
        a, b = pd.read_input(f), op.read_input(f)
        for k in ("ng", "l", "gtype", "gr", "cfl", "visci", "inivel", "is_wallturb", "bforce", "is_forced", "velf", "dims", "sgstype", "hwm", "dtmax", "dt_f"):
            assert getattr(a, k) == getattr(b, k), (f, k)
        for k in ("cbcvel", "cbcpre", "cbcsgs", "bcvel", "bcpre", "bcsgs", "lwm"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), (f, k)


# ---- writer: a Deck as an input.nml in the reference's layout, read back by both readers -------------------------------------
def _same(a, b, names):
    import numpy as np
    for nm in names:
        x, y = getattr(a, nm), getattr(b, nm)
        if isinstance(x, np.ndarray):
            assert np.array_equal(x, y), nm
        else:
            assert x == y and type(x) is type(y), (nm, x, y)


@pytest.mark.parametrize("name", ["config1", "config2", "config3", "config4_duct", "config4_cavity", "config5"])
def test_write_input_round_trip(name, tmp_path):
    import dataclasses
    import cales_b200.deck as pd
    import oracle.param as op
    d = pd.BASELINE_DECKS[name]()
    d.nstep, d.isave, d.dims, d.gr, d.bforce = 100, 100, (2, 4), 1.25e-5, (0., -9.81, 1e-7)
    p = tmp_path / "input.nml"
    txt = pd.write_input(d, str(p))
    assert txt.startswith("&dns\n") and "\n&les\n" in txt and txt.count("\n/\n") == 2
    back = pd.read_input(str(p))
    _same(d, back, [f.name for f in dataclasses.fields(pd.Deck) if f.name not in ("impdiff", "impdiff_1d", "ipencil")])
    ob = op.read_input(str(p))                               # the oracle's independent reader takes the same file
    _same(d, ob, ["ng", "l", "gtype", "gr", "cfl", "dtmax", "dt_f", "visci", "inivel", "is_wallturb", "cbcvel", "cbcpre", "cbcsgs",
                  "bcvel", "bcpre", "bcsgs", "bforce", "is_forced", "velf", "dims", "sgstype", "lwm", "hwm"])


def test_write_input_names_the_build_options(tmp_path):
    import cales_b200.deck as pd
    d = pd.deck_channel(ng=(16, 16, 16))
    d.impdiff, d.impdiff_1d, d.ipencil = True, True, 3
    last = pd.write_input(d).strip().splitlines()[-1]
    assert last.startswith("!") and "-D_IMPDIFF" in last and "-D_IMPDIFF_1D" in last and "-D_DECOMP_Z" in last
