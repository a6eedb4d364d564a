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
# minimal namelist reader for the reference's input.nml decks (param.f90:88-152)
# ---------------------------------------------------------------------------
def _parse_values(s):
    out = []
    for tok in re.split(r"[,\s]+", s.strip()):
        if not tok:
            continue
        t = tok.strip()
        if t.startswith(("'", '"')):
            out.append(t.strip("'\""))
        elif t.upper() in ("T", ".TRUE.", "TRUE"):
            out.append(True)
        elif t.upper() in ("F", ".FALSE.", "FALSE"):
            out.append(False)
        else:
            try:
                out.append(int(t))
            except ValueError:
                out.append(float(t.replace("d", "e").replace("D", "e")))
    return out


def read_input(path):
    """Read &dns and &les of an input.nml into a Deck (param.f90:88-152).  Groups may be terminated by '/' or,
    as in some DNS decks, by a backslash."""
    txt = "\n".join(line.split("!")[0] for line in open(path).read().splitlines())
    d = Deck()
    arrays = {"cbcvel": d.cbcvel, "bcvel": d.bcvel, "cbcpre": d.cbcpre, "cbcsgs": d.cbcsgs, "bcpre": d.bcpre, "bcsgs": d.bcsgs,
              "lwm": d.lwm}
    scal = {}
    for gm in re.finditer(r"&(\w+)(.*?)^\s*[/\\]\s*$", txt, re.S | re.M):
        if gm.group(1).lower() not in ("dns", "les"):
            continue
        body = gm.group(2)
        ms = list(re.finditer(r"([A-Za-z_]\w*)\s*(\([^)]*\))?\s*=", body))
        for i, m in enumerate(ms):
            name = m.group(1).lower()
            vals = _parse_values(body[m.end():ms[i + 1].start() if i + 1 < len(ms) else len(body)])
            if name in arrays:
                arr = arrays[name]
                spec = (m.group(2) or "").strip("()").split(",")
                if arr.ndim == 3:
                    k = int(spec[2]) - 1 if len(spec) == 3 and ":" not in spec[2] else None
                    ks = [k] if k is not None else range(3)
                    it = iter(vals)
                    for kk in ks:
                        for idir in range(3):
                            for ib in range(2):
                                arr[ib, idir, kk] = next(it)
                else:
                    it = iter(vals)
                    for idir in range(3):
                        for ib in range(2):
                            arr[ib, idir] = next(it)
            else:
                scal[name] = vals
    def get(name, default):
        return scal.get(name, default)
    d.ng = tuple(int(x) for x in get("ng", d.ng)); d.l = tuple(float(x) for x in get("l", d.l))
    d.gtype = int(get("gtype", [d.gtype])[0]); d.gr = float(get("gr", [d.gr])[0])
    d.cfl = float(get("cfl", [d.cfl])[0]); d.dtmax = float(get("dtmax", [d.dtmax])[0])
    d.dt_f = float(get("dt_f", [d.dt_f])[0]); d.visci = float(get("visci", [d.visci])[0])
    d.inivel = get("inivel", [d.inivel])[0]; d.is_wallturb = bool(get("is_wallturb", [d.is_wallturb])[0])
    d.bforce = tuple(float(x) for x in get("bforce", d.bforce))
    d.is_forced = tuple(bool(x) for x in get("is_forced", d.is_forced)); d.velf = tuple(float(x) for x in get("velf", d.velf))
    d.dims = tuple(int(x) for x in get("dims", d.dims))
    d.nstep = int(get("nstep", [d.nstep])[0]); d.time_max = float(get("time_max", [d.time_max])[0])
    d.tw_max = float(get("tw_max", [d.tw_max])[0]); d.stop_type = tuple(bool(x) for x in get("stop_type", d.stop_type))
    d.restart = bool(get("restart", [d.restart])[0]); d.is_overwrite_save = bool(get("is_overwrite_save", [d.is_overwrite_save])[0])
    d.nsaves_max = int(get("nsaves_max", [d.nsaves_max])[0])
    for nm in ("icheck", "iout0d", "iout1d", "iout2d", "iout3d", "isave"):
        setattr(d, nm, int(get(nm, [getattr(d, nm)])[0]))
    d.sgstype = get("sgstype", [d.sgstype])[0]; d.hwm = float(get("hwm", [d.hwm])[0])
    return d
