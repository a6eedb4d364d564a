"""A-priori checks of an input deck, `test_sanity_input` of src/sanity.f90:33-293, so that an invalid `input.nml` stops
before the first kernel with the reference's own messages instead of producing a wrong run.

Differences from the reference, all on the permissive side of what THIS library implements:
  * the reference's GPU build rejects pressure BCs 'ND'/'DN' along x or y (sanity.f90:263-272) because its GPU transforms
    lack them (fft.f90:567-569); this library implements all ten transform kinds, so the check is not applied;
  * `_IMPDIFF_1D` without `_IMPDIFF` cannot be expressed here (deck.impdiff_1d implies deck.impdiff in the driver), the
    corresponding error (sanity.f90:60-62) is reported for an inconsistent deck object."""
import numpy as np

VALID = ("PP", "ND", "DN", "NN", "DD")


class SanityError(ValueError):
    """*** Simulation aborted due to errors in the input file *** (sanity.f90:295-303)"""


def _pair(c, idir, ivel=None):
    return (c[0, idir, ivel] + c[1, idir, ivel]) if ivel is not None else (c[0, idir] + c[1, idir])


def chk_stop_type(stop_type):                                         # sanity.f90:69-78
    return [] if any(stop_type) else ["ERROR: stopping criterion not chosen."]


def chk_dims(ng, dims, cbcvel, sgstype, ipencil=1):                   # sanity.f90:80-113
    errs = []
    ii = [d for d in (0, 1, 2) if d != ipencil - 1]
    if not (all(dims[q] <= ng[ii[q]] for q in range(2)) and all(d >= 1 for d in dims)):
        errs.append("ERROR: 1 <= dims(:) <= [itot,jtot], or [itot,ktot], or [jtot ktot] depending on the decomposition.")
    if sgstype.strip() == "smag":
        ok = True
        for q in range(2):
            idir = ivel = ii[q]
            if _pair(cbcvel, idir, ivel) == "DD":
                ok = ok and dims[q] <= 2
        if not ok:
            errs.append("ERROR: more than two subdomains between two opposite walls.")
    return errs


def chk_bc(deck, n, is_bound, zc):                                    # sanity.f90:115-274
    cbcvel, cbcpre, cbcsgs, bcpre, bcvel, lwm = deck.cbcvel, deck.cbcpre, deck.cbcsgs, deck.bcpre, deck.bcvel, deck.lwm
    l, dl, h = deck.l, deck.dl, deck.hwm
    errs = []
    if not all(_pair(cbcvel, idir, ivel) in VALID for ivel in range(3) for idir in range(3)):
        errs.append("ERROR: velocity BCs not valid.")
    if not all(_pair(cbcpre, idir) in VALID for idir in range(3)):
        errs.append("ERROR: pressure BCs not valid.")
    compat_p = {"PP": "PP", "ND": "DN", "DN": "ND", "DD": "NN", "NN": "DD"}
    if not all(compat_p.get(_pair(cbcvel, d, d)) == _pair(cbcpre, d) for d in range(3)):
        errs.append("ERROR: velocity and pressure BCs not compatible.")
    if not all(_pair(cbcsgs, idir) in VALID for idir in range(3)):
        errs.append("ERROR: sgs BCs not valid.")
    compat_s = {"PP": "PP", "ND": "DD", "DN": "DD", "DD": "DD", "NN": "DD"}
    if not all(compat_s.get(_pair(cbcvel, d, d)) == _pair(cbcsgs, d) for d in range(3)):
        errs.append("ERROR: velocity and sgs BCs not compatible.")
    if not all(bcpre[0, d] == 0. and bcpre[1, d] == 0. for d in range(2)):
        errs.append("ERROR: pressure BCs in directions x and y must be homogeneous (value = 0.).")
    if not all(cbcvel[i, d, v] == "D" for d in range(3) for i in range(2) if lwm[i, d] != 0 for v in range(3)):
        errs.append("ERROR: wall model BCs must be Dirichlet.")
    ok = True
    for d in range(2):                                                # sanity.f90:224-227
        for i in range(2):
            if is_bound[i, d] and lwm[i, d] != 0:
                ok = ok and (h > 0.5 * dl[d] and h < (n[d] - 0.5) * dl[d])
    if is_bound[0, 2] and lwm[0, 2] != 0:
        ok = ok and (h > zc[1] and h < zc[n[2]])
    if is_bound[1, 2] and lwm[1, 2] != 0:
        ok = ok and (h > l[2] - zc[n[2]] and h < l[2] - zc[1])
    if not ok:
        errs.append("ERROR: invalid wall model height.")
    if deck.impdiff and not deck.impdiff_1d:                          # sanity.f90:233-261
        if any(_pair(cbcvel, d, v) == "NN" for v in range(3) for d in range(2)):
            errs.append("ERROR: Neumann-Neumann velocity BCs with implicit diffusion currently not supported in x and y; only in z.")
        if not all(bcvel[0, d, v] == 0. and bcvel[1, d, v] == 0. for v in range(3) for d in range(2)):
            errs.append("ERROR: velocity BCs with implicit diffusion in directions x and y must be homogeneous (value = 0.).")
        if not (lwm[0, 0] == 0 and lwm[1, 0] == 0 and lwm[0, 1] == 0 and lwm[1, 1] == 0):
            errs.append("ERROR: wall model BCs cannot be used in x and y directions when 3D implicit diffusion is applied.")
    return errs


def chk_forcing(cbcpre, is_forced):                                   # sanity.f90:276-293
    if all(_pair(cbcpre, d) == "PP" for d in range(3) if is_forced[d]):
        return []
    return ["ERROR: Flow cannot be forced in a non-periodic direction; check the BCs and is_forced in `input.nml`."]


def test_sanity_input(deck, n, is_bound, zc):
    """All checks in the reference's order; raises SanityError listing every failed one (the reference aborts at the first
    failing group, sanity.f90:56-59)."""
    errs = (chk_dims(deck.ng, deck.dims, deck.cbcvel, deck.sgstype, deck.ipencil) + chk_stop_type(deck.stop_type) +
            chk_bc(deck, n, np.asarray(is_bound), zc) + chk_forcing(deck.cbcpre, deck.is_forced))
    if deck.impdiff_1d and not deck.impdiff:
        errs.append("ERROR: `_IMPDIFF_1D` cpp macro requires building with `_IMPDIFF` too.")
    if errs:
        raise SanityError("\n".join(errs) + "\n*** Simulation aborted due to errors in the input file ***\n    check INFO_INPUT.md")
    return True


test_sanity_input.__test__ = False      # not a pytest test
