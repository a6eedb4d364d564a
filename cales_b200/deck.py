"""Input deck and constants of the host side.  Mirrors src/param.f90:18-33 (constants) and
src/param.f90:88-165 (namelists &dns, &les; derived dl, dli, visc).  Build-time cpp switches of the
reference (_IMPDIFF, _IMPDIFF_1D, _DECOMP_X/_Y/_Z) are run-time fields here."""
from __future__ import annotations

import dataclasses
import re
from dataclasses import dataclass, field

import numpy as np

pi = float(np.arccos(-1.0))                      # param.f90:18
eps = float(np.finfo(np.float64).eps)            # param.f90:20  epsilon(1._rp)
small = eps * 10 ** (15 // 2)                    # param.f90:24  epsilon*10**(precision/2), precision(dp)=15
big = float(np.finfo(np.float64).max)            # param.f90:25
rkcoeff = np.array([[32.0 / 60.0, 0.0],
                    [25.0 / 60.0, -17.0 / 60.0],
                    [45.0 / 60.0, -25.0 / 60.0]])  # param.f90:27-29, rkcoeff[irk] = rkcoeff(:,irk+1)
kap_log = 0.41                                   # param.f90:31
b_log = 5.20                                     # param.f90:32
c_smag = 0.11                                    # param.f90:33


def _c3(a, b, c):
    return np.array([[a[0], b[0], c[0]], [a[1], b[1], c[1]]], dtype="U1")


@dataclass
class Deck:
    """The &dns and &les namelists (param.f90:95-120).  cbcvel[ib,idir,ivel]."""
    ng: tuple = (64, 64, 64)
    l: tuple = (6.0, 3.0, 1.0)
    gtype: int = 1
    gr: float = 0.0
    cfl: float = 0.95
    dtmax: float = 1.0e5
    dt_f: float = -1.0
    visci: float = 5640.0
    inivel: str = "poi"
    is_wallturb: bool = False
    cbcvel: np.ndarray = field(default_factory=lambda: np.full((2, 3, 3), "P", dtype="U1"))
    cbcpre: np.ndarray = field(default_factory=lambda: np.full((2, 3), "P", dtype="U1"))
    cbcsgs: np.ndarray = field(default_factory=lambda: np.full((2, 3), "P", dtype="U1"))
    bcvel: np.ndarray = field(default_factory=lambda: np.zeros((2, 3, 3)))
    bcpre: np.ndarray = field(default_factory=lambda: np.zeros((2, 3)))
    bcsgs: np.ndarray = field(default_factory=lambda: np.zeros((2, 3)))
    bforce: tuple = (0.0, 0.0, 0.0)
    is_forced: tuple = (False, False, False)
    velf: tuple = (0.0, 0.0, 0.0)
    dims: tuple = (1, 1)
    # run control (param.f90:104-108; consumed by cales_b200/run.py, main.f90:358-369, 512-611)
    nstep: int = 100
    time_max: float = 100.0
    tw_max: float = 0.1
    stop_type: tuple = (True, False, False)
    restart: bool = False
    is_overwrite_save: bool = True
    nsaves_max: int = 0
    icheck: int = 10
    iout0d: int = 10
    iout1d: int = 0
    iout2d: int = 0
    iout3d: int = 0
    isave: int = 0
    sgstype: str = "none"
    lwm: np.ndarray = field(default_factory=lambda: np.zeros((2, 3), dtype=np.int32))
    hwm: float = 0.1
    # build-time switches of the reference (cpp macros), selected at run time here
    impdiff: bool = False       # _IMPDIFF
    impdiff_1d: bool = False    # _IMPDIFF_1D
    ipencil: int = 1            # 1,2,3 <- _DECOMP_X/_Y/_Z (initmpi.f90:56-62)

    def copy(self, **kw):
        d = dataclasses.replace(self)
        for name in ("cbcvel", "cbcpre", "cbcsgs", "bcvel", "bcpre", "bcsgs", "lwm"):
            setattr(d, name, getattr(self, name).copy())
        for k, v in kw.items():
            setattr(d, k, v)
        return d

    # derived quantities, param.f90:154-158
    @property
    def dl(self):
        return np.array([self.l[i] / (1.0 * self.ng[i]) for i in range(3)])

    @property
    def dli(self):
        return self.dl ** (-1)

    @property
    def visc(self):
        return self.visci ** (-1)


# ---------------------------------------------------------------------------
# Decks for the BASELINE.json configurations.  Grid sizes are
# overridable so that the same physics can be run at oracle-friendly sizes.
# ---------------------------------------------------------------------------
def deck_channel(ng=(64, 64, 64), sgstype="dsmag", wall_model=False, visci=5640.0, gtype=1, gr=5.0,
                 l=(6.0, 3.0, 1.0), dims=(1, 1)):
    """Configs 1/3/5: periodic channel, walls in z (examples/dns/_manuscript_turbulent_channel
    and examples/les/_manuscript_turbulent_channel[_wall_model]/input.nml)."""
    d = Deck(ng=tuple(ng), l=tuple(l), gtype=gtype, gr=gr, visci=visci, inivel="poi", is_wallturb=True,
             is_forced=(True, False, False), velf=(1.0, 0.0, 0.0), sgstype=sgstype, dims=tuple(dims))
    for ivel in range(3):
        d.cbcvel[:, 2, ivel] = "D"
    d.cbcpre[:, 2] = "N"
    d.cbcsgs[:, 2] = "D"
    if wall_model:
        d.lwm[:, 2] = 1
        d.hwm = 0.1
    return d


def deck_tgv(ng=(256, 256, 256), sgstype="smag", visci=1600.0, dims=(1, 1)):
    """Config 2: tri-periodic decaying Taylor-Green vortex
    (examples/dns/_manuscript_taylor_green_vortex/input.nml)."""
    two_pi = 2.0 * pi
    return Deck(ng=tuple(ng), l=(two_pi, two_pi, two_pi), gtype=1, gr=0.0, visci=visci, inivel="tgv",
                sgstype=sgstype, dims=tuple(dims))


def deck_duct(ng=(512, 256, 256), sgstype="smag", visci=4410.0, wall_model=False, dims=(1, 1)):
    """Config 4a: square duct, periodic in x, walls in y and z (examples/dns/periodic_duct)."""
    d = Deck(ng=tuple(ng), l=(10.0, 2.0, 2.0), gtype=1, gr=0.0, visci=visci, inivel="duc",
             is_forced=(True, False, False), velf=(1.0, 0.0, 0.0), sgstype=sgstype, dims=tuple(dims))
    for idir in (1, 2):
        for ivel in range(3):
            d.cbcvel[:, idir, ivel] = "D"
        d.cbcpre[:, idir] = "N"
        d.cbcsgs[:, idir] = "D"
    if wall_model:
        d.lwm[:, 1] = 1
        d.lwm[:, 2] = 1
    return d


def deck_cavity(ng=(512, 256, 256), sgstype="smag", visci=1000.0, dims=(1, 1)):
    """Config 4b: lid-driven cavity, all walls, lid moving in x at the top z face
    (examples/dns/lid_driven_cavity/input.nml: bcvel(1,3,1)=1)."""
    d = Deck(ng=tuple(ng), l=(1.0, 1.0, 1.0), gtype=1, gr=0.0, visci=visci, inivel="zer", sgstype=sgstype,
             dims=tuple(dims))
    d.cbcvel[:] = "D"
    d.cbcpre[:] = "N"
    d.cbcsgs[:] = "D"
    d.bcvel[1, 2, 0] = 1.0
    return d


# ---------------------------------------------------------------------------
# namelist reader for the reference's input.nml decks (read_input, src/param.f90:88-152)
# ---------------------------------------------------------------------------
class DeckError(ValueError):
    """an input deck the reference would not have accepted (or that this reader cannot interpret unambiguously)"""


_TOKEN = re.compile(r"""\s*(?:
      (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")      # quoted string
    | (?P<amp>&\w+)                               # group start
    | (?P<end>/|\\)                              # group terminator: '/' (or a backslash, as in some DNS decks) anywhere outside quotes
    | (?P<eq>=)
    | (?P<comma>,)
    | (?P<sub>\([^)]*\))                          # array section, e.g. (0:1,1:3,2)
    | (?P<word>[^\s,=/\\()'"!]+)                  # name, number, logical, repeat count n*value
    | (?P<comment>![^\n]*)
    )""", re.X)


def _tokens(txt):
    pos, out = 0, []
    while pos < len(txt):
        m = _TOKEN.match(txt, pos)
        if m is None or m.end() == pos:
            if txt[pos:].strip() == "":
                break
            raise DeckError("cannot tokenise the deck at: %r" % txt[pos:pos + 30])
        pos = m.end()
        kind = m.lastgroup
        if kind != "comment":
            out.append((kind, m.group(kind)))
    return out


def _value(word):
    u = word.upper()
    if u in ("T", ".T.", ".TRUE.", "TRUE"):
        return True
    if u in ("F", ".F.", ".FALSE.", "FALSE"):
        return False
    try:
        return int(word)
    except ValueError:
        pass
    try:
        return float(u.replace("D", "E"))
    except ValueError:
        raise DeckError("cannot interpret the value %r" % word)


def _groups(txt):
    """{group: [(name, section or None, [values])]} of a namelist file: values may be separated by commas and/or blanks, run
    over several lines, carry repeat counts (3*0.), and the group ends at the first '/' outside quotes."""
    toks = _tokens(txt)
    groups, i = {}, 0
    while i < len(toks):
        kind, tok = toks[i]
        if kind != "amp":
            i += 1
            continue
        gname = tok[1:].lower()
        entries, i = [], i + 1
        cur = None
        closed = False
        rep = 1
        while i < len(toks):
            kind, tok = toks[i]
            if kind == "end":
                closed = True
                i += 1
                break
            if kind == "amp":
                raise DeckError("group &%s is not terminated by '/'" % gname)
            if kind == "word" and i + 1 < len(toks) and (toks[i + 1][0] == "eq" or (toks[i + 1][0] == "sub" and i + 2 < len(toks) and toks[i + 2][0] == "eq")):
                sub = toks[i + 1][1] if toks[i + 1][0] == "sub" else None
                cur = (tok.lower(), sub, [])
                entries.append(cur)
                i += 3 if sub else 2
                continue
            if cur is None:
                raise DeckError("value %r before any name in group &%s" % (tok, gname))
            if kind == "str":
                q = tok[0]
                cur[2].extend([tok[1:-1].replace(q + q, q)] * rep)
                rep = 1
            elif kind == "word":
                if "*" in tok:                                  # repeat count r*value (the value may be a quoted string: next token)
                    r, _, v = tok.partition("*")
                    if not r.isdigit() or rep != 1:
                        raise DeckError("bad repeat count in %r" % tok)
                    if v == "":
                        rep = int(r)
                    else:
                        cur[2].extend([_value(v)] * int(r))
                else:
                    cur[2].extend([_value(tok)] * rep)
                    rep = 1
            elif kind != "comma":
                raise DeckError("unexpected %r in group &%s" % (tok, gname))
            i += 1
        if not closed:
            raise DeckError("group &%s is not terminated by '/'" % gname)
        groups.setdefault(gname, []).extend(entries)
    return groups


# name -> (type, count) of the scalar / vector entries of &dns and &les (param.f90:95-120); tables are handled separately
_DNS = {"ng": (int, 3), "l": (float, 3), "gtype": (int, 1), "gr": (float, 1), "cfl": (float, 1), "dtmax": (float, 1), "dt_f": (float, 1),
        "visci": (float, 1), "inivel": (str, 1), "is_wallturb": (bool, 1), "nstep": (int, 1), "time_max": (float, 1), "tw_max": (float, 1),
        "stop_type": (bool, 3), "restart": (bool, 1), "is_overwrite_save": (bool, 1), "nsaves_max": (int, 1), "icheck": (int, 1),
        "iout0d": (int, 1), "iout1d": (int, 1), "iout2d": (int, 1), "iout3d": (int, 1), "isave": (int, 1), "bforce": (float, 3),
        "is_forced": (bool, 3), "velf": (float, 3), "dims": (int, 2)}
_LES = {"sgstype": (str, 1), "hwm": (float, 1)}
_DNS_TABLES = ("cbcvel", "cbcpre", "cbcsgs", "bcvel", "bcpre", "bcsgs")    # namelist /dns/, param.f90:95-111
_LES_TABLES = ("lwm",)                                                      # namelist /les/, param.f90:112-115
_REQUIRED = ("ng", "l", "visci", "cbcvel", "cbcpre")


def _conv(name, typ, v):
    if typ is float and isinstance(v, (int, float)) and not isinstance(v, bool):
        return float(v)
    if typ is int and isinstance(v, int) and not isinstance(v, bool):
        return v
    if typ is bool and isinstance(v, bool):
        return v
    if typ is str and isinstance(v, str):
        return v
    raise DeckError("entry %s: value %r is not of type %s" % (name, v, typ.__name__))


def read_input(path):
    """Read &dns and &les of an input.nml into a Deck (read_input, src/param.f90:88-152).  Raises DeckError when there is no
    &dns group, a required entry (ng, l, visci, cbcvel, cbcpre) is missing, a name is unknown to the group, a group is
    not terminated, or a value has the wrong type or count -- where the Fortran runtime would stop with an I/O error."""
    groups = _groups(open(path).read())
    if "dns" not in groups:
        raise DeckError("%s: no &dns group found" % path)
    d = Deck()
    seen = set()
    for gname, scalars, tables in (("dns", _DNS, _DNS_TABLES), ("les", _LES, _LES_TABLES)):
        for name, sub, vals in groups.get(gname, []):
            seen.add(name)
            if name in tables:
                arr = getattr(d, name)
                typ = str if arr.dtype.kind == "U" else (int if arr.dtype.kind == "i" else float)
                if arr.ndim == 3:                                   # (0:1,1:3,1:3): Fortran order, optionally one velocity component
                    spec = [x.strip() for x in sub.strip("()").split(",")] if sub else []
                    comps = [int(spec[2]) - 1] if len(spec) == 3 and ":" not in spec[2] else [0, 1, 2]
                    if len(vals) != 6 * len(comps):
                        raise DeckError("entry %s%s: %d values, expected %d" % (name, sub or "", len(vals), 6 * len(comps)))
                    it = iter(vals)
                    for kk in comps:
                        for idir in range(3):
                            for ib in range(2):
                                arr[ib, idir, kk] = _conv(name, typ, next(it))
                else:
                    if len(vals) != 6:
                        raise DeckError("entry %s: %d values, expected 6" % (name, len(vals)))
                    it = iter(vals)
                    for idir in range(3):
                        for ib in range(2):
                            arr[ib, idir] = _conv(name, typ, next(it))
            elif name in scalars:
                typ, cnt = scalars[name]
                if len(vals) != cnt:
                    raise DeckError("entry %s: %d values, expected %d" % (name, len(vals), cnt))
                conv = [_conv(name, typ, v) for v in vals]
                setattr(d, name, conv[0] if cnt == 1 else tuple(conv))
            else:
                raise DeckError("unknown entry %r in group &%s" % (name, gname))
    missing = [k for k in _REQUIRED if k not in seen]
    if missing:
        raise DeckError("%s: required entries missing from &dns: %s" % (path, ", ".join(missing)))
    return d


# ---------------------------------------------------------------------------
# writer: a Deck as an input.nml the REFERENCE reads (namelists /dns/ and /les/, param.f90:95-120; layout of
# examples/les/_manuscript_turbulent_channel_wall_model/input.nml), so that a site-built CaLES and this library can be run
# on the same file ("Pinning the oracle off this box")
# ---------------------------------------------------------------------------
def _f(x):
    s = repr(float(x))
    return s.replace("e", "d") if "e" in s else s            # 1e-05 -> 1d-05: a double-precision literal either way


def _l(x):
    return "T" if x else "F"


def write_input(d, path=None):
    """Returns (and, with `path`, writes) the text of input.nml for deck `d`.  The cpp switches that are run-time options here
    (impdiff, impdiff_1d, ipencil) are build options of the reference and appear as a comment."""
    def pairs(arr, fmt):
        return ",  ".join(",".join(fmt(arr[ib, idir]) for ib in range(2)) for idir in range(3))
    q = lambda s: "'%s'" % s
    L = ["&dns",
         "ng(1:3) = %d, %d, %d" % tuple(d.ng),
         "l(1:3) = %s, %s, %s" % tuple(_f(x) for x in d.l),
         "gtype = %d, gr = %s" % (d.gtype, _f(d.gr)),
         "cfl = %s, dtmax = %s, dt_f = %s" % (_f(d.cfl), _f(d.dtmax), _f(d.dt_f)),
         "visci = %s" % _f(d.visci),
         "inivel = '%s'" % d.inivel,
         "is_wallturb = %s" % _l(d.is_wallturb),
         "nstep = %d, time_max = %s, tw_max = %s" % (d.nstep, _f(d.time_max), _f(d.tw_max)),
         "stop_type(1:3) = %s, %s, %s" % tuple(_l(x) for x in d.stop_type),
         "restart = %s, is_overwrite_save = %s, nsaves_max = %d" % (_l(d.restart), _l(d.is_overwrite_save), d.nsaves_max),
         "icheck = %d, iout0d = %d, iout1d = %d, iout2d = %d, iout3d = %d, isave = %d" % (d.icheck, d.iout0d, d.iout1d, d.iout2d, d.iout3d, d.isave)]
    for c in range(3):
        L.append("cbcvel(0:1,1:3,%d) = %s" % (c + 1, pairs(d.cbcvel[:, :, c], q)))
    L.append("cbcpre(0:1,1:3)   = %s" % pairs(d.cbcpre, q))
    L.append("cbcsgs(0:1,1:3)   = %s" % pairs(d.cbcsgs, q))
    for c in range(3):
        L.append("bcvel(0:1,1:3,%d) = %s" % (c + 1, pairs(d.bcvel[:, :, c], _f)))
    L.append("bcpre(0:1,1:3)   = %s" % pairs(d.bcpre, _f))
    L.append("bcsgs(0:1,1:3)   = %s" % pairs(d.bcsgs, _f))
    L += ["bforce(1:3) = %s, %s, %s" % tuple(_f(x) for x in d.bforce),
          "is_forced(1:3) = %s, %s, %s" % tuple(_l(x) for x in d.is_forced),
          "velf(1:3) = %s, %s, %s" % tuple(_f(x) for x in d.velf),
          "dims(1:2) = %d, %d" % tuple(d.dims),
          "/", "",
          "&les",
          "sgstype = '%s'" % d.sgstype,
          "lwm(0:1,1:3) = %s" % pairs(d.lwm, lambda x: "%d" % x),
          "hwm = %s" % _f(d.hwm),
          "/", ""]
    build = []
    if d.impdiff:
        build.append("-D_IMPDIFF")
    if d.impdiff_1d:
        build.append("-D_IMPDIFF_1D")
    build.append("-D_DECOMP_%s" % "XYZ"[d.ipencil - 1])
    L.append("! reference build options matching this deck: " + " ".join(build))
    txt = "\n".join(L) + "\n"
    if path:
        with open(path, "w") as f:
            f.write(txt)
    return txt


BASELINE_DECKS = {   # BASELINE.json configs; nstep / isave as the pinning recipe wants them
    "config1": lambda: deck_channel(ng=(64, 64, 64), sgstype="dsmag"),
    "config2": lambda: deck_tgv(ng=(256, 256, 256), sgstype="smag"),
    "config3": lambda: deck_channel(ng=(512, 256, 192), sgstype="smag", wall_model=True, gtype=6, gr=0., l=(12.8, 4.8, 2.), visci=43500.),
    "config4_duct": lambda: deck_duct(ng=(512, 256, 256), sgstype="smag"),
    "config4_cavity": lambda: deck_cavity(ng=(512, 256, 256), sgstype="smag"),
    "config5": lambda: deck_channel(ng=(1024, 512, 512), sgstype="smag", wall_model=True, gtype=6, gr=0., l=(12.8, 4.8, 2.), visci=43500.),
}
