"""Host-side model of the distributed tridiagonal z solve (cales_b200/csrc/zdist.cu), statement for statement in numpy, against
dense solves: the algebra the CUDA kernels implement (block spikes from forward / backward pivots, the 2P x 2P interface
system, its pinned variant for the singular mean mode, cyclic coupling for periodic z) is checked here without a GPU; the
kernels themselves are checked against the oracle's gaussel / gaussel_periodic in tests/test_gpu_kernels.py."""
import numpy as np
import pytest

EPS = 2.220446049250313e-16


def thomas(a, b, c, r):                         # dgtsv_homebrewed, src/solver.f90:153-179
    n = len(b); z = np.zeros(n); p = np.zeros(n); d = 0.; pl = 0.
    for l in range(n):
        z[l] = 1. / (b[l] - a[l] * d + EPS); d = c[l] * z[l]; pl = (r[l] - a[l] * pl) * z[l]; p[l] = pl
    for l in range(n - 2, -1, -1):
        p[l] = p[l] - c[l] * z[l] * p[l + 1]
    return p


def distribute(n, P):                           # decomp_2d.f90:1132-1145
    q, r = divmod(n, P)
    st = [i * q + min(i, r) for i in range(P)]
    return st + [n]


def zdist(a, b, c, lam, r, P, periodic, pin):
    n = len(a); bb = b + lam; zs = distribute(n, P)
    vf = np.zeros(P); vl = np.zeros(P); wf = np.zeros(P); wl = np.zeros(P); V = [None] * P; W = [None] * P; Y = [None] * P
    for s in range(P):                          # zd_build_k, first part
        z0, z1 = zs[s], zs[s + 1]; m = z1 - z0
        A, B, C = a[z0:z1], bb[z0:z1], c[z0:z1]
        has_prev, has_next = periodic or s > 0, periodic or s < P - 1
        d = 0.; zf = np.zeros(m)
        for l in range(m):
            zf[l] = 1. / (B[l] - A[l] * d + EPS); d = C[l] * zf[l]
        w = np.zeros(m); w[m - 1] = C[m - 1] * zf[m - 1] if has_next else 0.
        for l in range(m - 2, -1, -1):
            w[l] = -(C[l] * zf[l]) * w[l + 1]
        d = 0.; zb = np.zeros(m)
        for l in range(m - 1, -1, -1):
            zb[l] = 1. / (B[l] - C[l] * d + EPS); d = A[l] * zb[l]
        v = np.zeros(m); v[0] = A[0] * zb[0] if has_prev else 0.
        for l in range(1, m):
            v[l] = -(A[l] * zb[l]) * v[l - 1]
        V[s], W[s] = v, w; vf[s], vl[s], wf[s], wl[s] = v[0], v[-1], w[0], w[-1]
        Y[s] = thomas(A, B, C, r[z0:z1])        # phase 1: the local solve
    U = 2 * P; M = np.zeros((U, U)); rhs = np.zeros(U)
    for s in range(P):                          # zd_build_k, second part (there: rows of the inverse, tabulated)
        ip, inx = 2 * ((s - 1) % P) + 1, 2 * ((s + 1) % P)
        for e in range(2):
            row = 2 * s + e
            if pin and row == 1:
                M[1, 1] = 1.; continue
            M[row, row] += 1.; M[row, ip] += (vl if e else vf)[s]; M[row, inx] += (wl if e else wf)[s]
            rhs[row] = Y[s][-1 if e else 0]
    u = np.linalg.solve(M, rhs)
    x = np.zeros(n)
    for s in range(P):                          # zd_reduce_k + zd_correct_k
        xp = u[2 * ((s - 1) % P) + 1] if (periodic or s > 0) else 0.
        xn = u[2 * ((s + 1) % P)] if (periodic or s < P - 1) else 0.
        x[zs[s]:zs[s + 1]] = Y[s] - V[s] * xp - W[s] * xn
    return x


def system(n, periodic, rng):
    dz = 1 + 0.5 * rng.random(n + 2)
    a = 1 / dz[1:n + 1] / (0.5 * (dz[0:n] + dz[1:n + 1])); c = 1 / dz[1:n + 1] / (0.5 * (dz[1:n + 1] + dz[2:n + 2])); b = -(a + c)
    if periodic:
        a[0] = c[-1] = 1 / dz[1] / (0.5 * (dz[1] + dz[n])); b[0] = -(a[0] + c[0]); b[-1] = -(a[-1] + c[-1])
    else:
        b[0] += a[0]; b[-1] += c[-1]
    return a, b, c


def dense(a, b, c, lam, periodic):
    n = len(a); A = np.diag(b + lam) + np.diag(a[1:], -1) + np.diag(c[:-1], 1)
    if periodic:
        A[0, n - 1] += a[0]; A[n - 1, 0] += c[n - 1]
    return A


@pytest.mark.parametrize("periodic", [0, 1])
@pytest.mark.parametrize("P", [2, 3, 4, 8])
@pytest.mark.parametrize("n", [16, 37, 64, 256])
def test_zdist_model_vs_dense(n, P, periodic):
    rng = np.random.default_rng(100 * n + 10 * P + periodic)
    a, b, c = system(n, periodic, rng)
    for lam in (-3.7, -1e-3, -1e-6):
        r = rng.standard_normal(n)
        xr = np.linalg.solve(dense(a, b, c, lam, periodic), r)
        x = zdist(a, b, c, lam, r, P, periodic, False)
        assert np.abs(x - xr).max() <= 1e-9 * max(1., 1e-4 / abs(lam)) * np.abs(xr).max()
    # the singular mean mode: compatible right-hand side, solution up to its additive constant
    A = dense(a, b, c, 0., periodic)
    wl = np.linalg.svd(A)[0][:, -1]
    r = rng.standard_normal(n); r -= wl * (wl @ r)
    x = zdist(a, b, c, 0., r, P, periodic, True)
    assert np.abs(A @ x - r).max() <= 1e-9 * np.abs(r).max()
    xr = np.linalg.lstsq(A, r, rcond=None)[0]
    assert np.abs((x - x.mean()) - (xr - xr.mean())).max() <= 1e-8 * np.abs(xr - xr.mean()).max()


# ---- the same algorithm across REAL processes (torch.distributed / gloo): every rank holds only its block of the right-hand
# side, solves it locally, contributes its first and last value to an all-gather (the push of two boundary planes into every
# rank's gather buffer in zdist.cu), forms its two interface unknowns from the gathered values and corrects its block -------
def _spikes(a, bb, c, z0, z1, has_prev, has_next):
    A, B, C = a[z0:z1], bb[z0:z1], c[z0:z1]; m = z1 - z0
    d = 0.; zf = np.zeros(m)
    for l in range(m):
        zf[l] = 1. / (B[l] - A[l] * d + EPS); d = C[l] * zf[l]
    w = np.zeros(m); w[m - 1] = C[m - 1] * zf[m - 1] if has_next else 0.
    for l in range(m - 2, -1, -1):
        w[l] = -(C[l] * zf[l]) * w[l + 1]
    d = 0.; zb = np.zeros(m)
    for l in range(m - 1, -1, -1):
        zb[l] = 1. / (B[l] - C[l] * d + EPS); d = A[l] * zb[l]
    v = np.zeros(m); v[0] = A[0] * zb[0] if has_prev else 0.
    for l in range(1, m):
        v[l] = -(A[l] * zb[l]) * v[l - 1]
    return v, w


def _zdist_rank(rank, P, n, periodic, port, q):
    try:
        import os
        import torch
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=P)
        rng = np.random.default_rng(7 * n + P + periodic)        # the coefficients are global (every rank has a, b, c, lambda)
        a, b, c = system(n, periodic, rng)
        ncol = 5
        lams = -np.abs(rng.standard_normal(ncol)) - 1e-3
        R = rng.standard_normal((ncol, n))
        zs = distribute(n, P); z0, z1 = zs[rank], zs[rank + 1]
        mine = R[:, z0:z1].copy()                                # this rank's block of the right-hand side: all it ever reads of R
        x_loc = np.zeros_like(mine)
        ends = torch.zeros(ncol, 2, dtype=torch.float64)
        Y, VW = [], []
        for col in range(ncol):
            bb = b + lams[col]
            y = thomas(a[z0:z1], bb[z0:z1], c[z0:z1], mine[col])
            Y.append(y); ends[col, 0] = y[0]; ends[col, 1] = y[-1]
            VW.append(_spikes(a, bb, c, z0, z1, periodic or rank > 0, periodic or rank < P - 1))
        gathered = [torch.zeros_like(ends) for _ in range(P)]
        dist.all_gather(gathered, ends)                          # 2 values per column and rank cross the network, nothing else
        for col in range(ncol):
            bb = b + lams[col]
            U = 2 * P; M = np.zeros((U, U)); rhs = np.zeros(U)
            for s in range(P):                                   # interface system from the spikes of every block (coefficients only)
                v, w = _spikes(a, bb, c, zs[s], zs[s + 1], periodic or s > 0, periodic or s < P - 1)
                ip, inx = 2 * ((s - 1) % P) + 1, 2 * ((s + 1) % P)
                for e in range(2):
                    row = 2 * s + e
                    M[row, row] += 1.; M[row, ip] += v[-1 if e else 0]; M[row, inx] += w[-1 if e else 0]
                    rhs[row] = float(gathered[s][col, e])
            u = np.linalg.solve(M, rhs)
            xp = u[2 * ((rank - 1) % P) + 1] if (periodic or rank > 0) else 0.
            xn = u[2 * ((rank + 1) % P)] if (periodic or rank < P - 1) else 0.
            v, w = VW[col]
            x_loc[col] = Y[col] - v * xp - w * xn
        # check against the dense solve of the whole system (every rank can form it: test only)
        for col in range(ncol):
            xr = np.linalg.solve(dense(a, b, c, lams[col], periodic), R[col])
            assert np.abs(x_loc[col] - xr[z0:z1]).max() <= 1e-9 * max(1., 1e-4 / abs(lams[col])) * np.abs(xr).max()
        dist.barrier(); dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))


@pytest.mark.parametrize("P,n,periodic,port", [(2, 24, 0, 29621), (2, 24, 1, 29622), (4, 37, 1, 29623), (3, 20, 0, 29624)])
def test_zdist_across_gloo_processes(P, n, periodic, port):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_zdist_rank, args=(r, P, n, periodic, port, q)) for r in range(P)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(P)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
