"""Wall-normal grid.  Follows src/initgrid.f90:15-196 statement by statement, including the
default-real (single precision) literals the Fortran source carries: `z0 = (k-0.)/(1.*n)` is a
real(4) division promoted to real(8) (initgrid.f90:53)."""
import numpy as np

from .param import pi

f32 = np.float32


def _two_end(kg, nzg, alpha, z0):            # initgrid.f90:87-101
    if alpha != 0.0:
        return 0.5 * (1.0 + np.tanh((z0 - 0.5) * alpha) / np.tanh(alpha / 2.0))
    return z0


def _one_end(kg, nzg, alpha, z0):            # initgrid.f90:102-116
    if alpha != 0.0:
        return 1.0 * (1.0 + np.tanh((z0 - 1.0) * alpha) / np.tanh(alpha / 1.0))
    return z0


def _one_end_r(kg, nzg, alpha, r0):          # initgrid.f90:117-131
    if alpha != 0.0:
        return 1.0 - 1.0 * (1.0 + np.tanh((1.0 - r0 - 1.0) * alpha) / np.tanh(alpha / 1.0))
    return r0


def _middle(kg, nzg, alpha, z0):             # initgrid.f90:132-152
    if alpha != 0.0:
        if z0 <= 0.5:
            return 0.5 * (1.0 - 1.0 + np.tanh(2.0 * alpha * (z0 - 0.0)) / np.tanh(alpha))
        return 0.5 * (1.0 + 1.0 + np.tanh(2.0 * alpha * (z0 - 1.0)) / np.tanh(alpha))
    return z0


def _wall_model(kg, nzg, alpha, z0):         # initgrid.f90:153-166
    # dzc = 0.1*32./nzg is evaluated in default real (single) and stored in real(rp)
    dzc = float(f32(f32(0.1) * f32(32.0)) / f32(nzg))
    return z0 - (dzc * nzg / 2.0 - 1.0) / (2.0 * pi) * np.sin(2.0 * pi * z0)


def _natural(kg, nzg, dummy, z0):            # initgrid.f90:167-195
    kb, alpha, c_eta, dyp = 32.0, pi / 1.5, 0.8, 0.05
    n = nzg / 2.0
    retau = 1.0 / (1.0 + (n / kb) ** 2) * (dyp * n + (3.0 / 4.0 * alpha * c_eta * n) ** (4.0 / 3.0) * (n / kb) ** 2)
    k = 1.0 * min(kg, (nzg - kg))
    z = 1.0 / (1.0 + (k / kb) ** 2) * (dyp * k + (3.0 / 4.0 * alpha * c_eta * k) ** (4.0 / 3.0) * (k / kb) ** 2) / (2.0 * retau)
    if kg > nzg - kg:
        z = 1.0 - z
    return z


_GRIDPOINT = {1: _two_end, 2: _one_end, 3: _one_end_r, 4: _middle, 5: _natural, 6: _wall_model}


def initgrid(gtype, n, gr, lz):
    """Returns dzc, dzf, zc, zf, each of extent 0:n+1 (initgrid.f90:15-81)."""
    gridpoint = _GRIDPOINT.get(gtype, _two_end)
    dzc = np.zeros(n + 2); dzf = np.zeros(n + 2); zc = np.zeros(n + 2); zf = np.zeros(n + 2)
    zf[0] = 0.0
    for k in range(1, n + 1):
        z0 = float(f32(k) / f32(n))                       # (k-0.)/(1.*n) in real(4)
        zf[k] = gridpoint(k, n, gr, z0)
        zf[k] = zf[k] * lz
    for k in range(1, n + 1):
        dzf[k] = zf[k] - zf[k - 1]
    dzf[0] = dzf[1]
    dzf[n + 1] = dzf[n]
    for k in range(0, n + 1):
        dzc[k] = .5 * (dzf[k] + dzf[k + 1])
    dzc[n + 1] = dzc[n]
    zc[0] = -dzc[0] / 2.0
    zf[0] = 0.0
    for k in range(1, n + 2):
        zc[k] = zc[k - 1] + dzc[k - 1]
        zf[k] = zf[k - 1] + dzf[k]
    return dzc, dzf, zc, zf
