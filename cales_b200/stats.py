"""Host side of the on-the-fly channel statistics: the reductions run on the device (cales_out1d_chan, csrc/stats.cu);
this module writes the reference's files -- out1d_single_point_chan, src/output.f90:684-691: `fname.out` with
zc_g, zf_g, the 27 profiles, dzc_g, dzf_g per line in (*(es24.16e3,1x)), and `fname.bin` with buf(1:27,1:ng3) as a stream."""
import ctypes as C

import numpy as np

from . import lib as L

NVARS = 27


def _es24(v):
    """Fortran ES24.16E3: one digit before the point, 16 after, three exponent digits, width 24."""
    if v != v or v in (float("inf"), float("-inf")):
        s = "NaN" if v != v else ("Infinity" if v > 0 else "-Infinity")
    elif v == 0.0:
        s = "0.0000000000000000E+000"
    else:
        m, e = ("%.16E" % v).split("E")
        s = "%sE%+04d" % (m, int(e))
    return s.rjust(24)


def out1d_chan(sim, fname=None):
    """Simulation -> buf(27, ng3) (rank-summed); rank 0 writes fname.out / fname.bin when fname is given."""
    d = sim.deck
    buf = np.zeros((NVARS, int(d.ng[2])), order="F")
    sim.chk(sim.lib.cales_out1d_chan(sim.ctx, L._ia(d.ng), L._ia(sim.lo), L._ia(sim.hi), L._da(d.l), L._da(d.dl), sim.d["dzc"].data_ptr(),
                                     sim.d["dzf"].data_ptr(), sim.ptr("u"), sim.ptr("v"), sim.ptr("w"), sim.ptr("p"), sim.ptr("visct"),
                                     buf.ctypes.data_as(L.c_dbl_p)))
    if fname is not None and sim.rank == 0:
        with open(fname + ".out", "w") as f:
            for k in range(1, int(d.ng[2]) + 1):
                row = [sim.zc_g[k], sim.zf_g[k]] + list(buf[:, k - 1]) + [sim.dzc_g[k], sim.dzf_g[k]]
                f.write("".join(_es24(float(x)) + " " for x in row).rstrip(" ") + "\n")
        buf.ravel(order="F").tofile(fname + ".bin")
    return buf
