"""fld.bin restart files (cales_b200/checkpoint.py vs src/load.f90:20-187).  CPU only: the format is host code."""
import os

import numpy as np
import pytest

from cales_b200 import checkpoint as ck
from oracle import decomp as od


def pencil(ng, dims, r):
    """x-pencil box of rank r (initmpi.f90:178-191 through the oracle's 2decomp partition)."""
    coord = (r // dims[1], r % dims[1])
    lo, hi, n = od.partition(ng, (1, 2, 3), dims, coord)
    return lo, hi, n


def reference_reader(filename, ng):
    """Re-statement of utils/read_binary_data/python/read_restart_file.py:44-58 (the reader the reference ships)."""
    disp = int(np.prod(ng))
    data = np.zeros([ng[0], ng[1], ng[2], 4])
    offset = 0
    with open(filename, "rb") as f:
        for q in range(4):
            f.seek(offset)
            fld = np.fromfile(f, dtype="float64", count=disp)
            data[:, :, :, q] = np.reshape(fld, (ng[0], ng[1], ng[2]), order="F")
            offset += 8 * disp
        f.seek(offset)
        info = np.fromfile(f, dtype="float64", count=2)
    return data, float(info[0]), int(info[1])


def halo(a):
    out = np.full(tuple(s + 2 for s in a.shape), np.nan, order="F")
    out[1:-1, 1:-1, 1:-1] = a
    return out


@pytest.mark.parametrize("ng,dims", [((8, 6, 5), (1, 1)), ((9, 7, 10), (2, 3)), ((16, 5, 7), (3, 1))])
def test_write_by_ranks_read_by_reference_reader(tmp_path, ng, dims):
    rng = np.random.default_rng(3)
    glob = [np.asfortranarray(rng.standard_normal(ng)) for _ in range(4)]
    fn = str(tmp_path / "fld.bin")
    nr = dims[0] * dims[1]
    for r in range(nr):                                   # ranks write one after the other into the shared file
        lo, hi, n = pencil(ng, dims, r)
        box = tuple(slice(lo[q] - 1, hi[q]) for q in range(3))
        f = [halo(g[box]) for g in glob]
        ck.load_all("w", fn, ng, lo, hi, *f, time=1.25, istep=77, rank=r, barrier=lambda: None)   # (ranks emulated one after the other)
    assert os.path.getsize(fn) == ck.expected_size(ng) == (np.prod(ng) * 4 + 2) * 8
    data, time, istep = reference_reader(fn, ng)
    for q in range(4):
        assert np.array_equal(data[:, :, :, q], glob[q])
    assert time == 1.25 and istep == 77
    # read back with a different decomposition: interiors filled, halos untouched
    dims2 = (dims[1], dims[0])
    for r in range(dims2[0] * dims2[1]):
        lo, hi, n = pencil(ng, dims2, r)
        box = tuple(slice(lo[q] - 1, hi[q]) for q in range(3))
        f = [np.full(tuple(n[q] + 2 for q in range(3)), -7., order="F") for _ in range(4)]
        t, i = ck.load_all("r", fn, ng, lo, hi, *f)
        assert (t, i) == (1.25, 77)
        for q in range(4):
            assert np.array_equal(f[q][1:-1, 1:-1, 1:-1], glob[q][box])
            assert f[q][0, 0, 0] == -7. and f[q][-1, -1, -1] == -7.


def test_size_check_and_errors(tmp_path):
    ng = (4, 4, 4)
    fn = str(tmp_path / "fld.bin")
    f = [halo(np.zeros(ng, order="F")) for _ in range(4)]
    with pytest.raises(ck.CheckpointError, match="not found"):
        ck.load_all("r", fn, ng, (1, 1, 1), ng, *f)
    with open(fn, "wb") as fh:
        fh.write(b"\0" * (ck.expected_size(ng) - 8))
    with pytest.raises(ck.CheckpointError, match="incorrect size"):      # load.f90:46-52
        ck.load_all("r", fn, ng, (1, 1, 1), ng, *f)
    with pytest.raises(ck.CheckpointError, match="outside"):
        ck.load_all("w", fn, ng, (1, 1, 1), (5, 4, 4), *f)
    with pytest.raises(ck.CheckpointError, match="does not match"):
        ck.load_all("w", fn, ng, (1, 1, 1), (4, 4, 3), *f, barrier=lambda: None)
    with pytest.raises(ck.CheckpointError, match="needs a barrier"):      # a sub-box write without a barrier could race rank 0's truncate
        g = [np.zeros((4, 6, 6), order="F") for _ in range(4)]
        ck.load_all("w", fn, ng, (1, 1, 1), (2, ng[1], ng[2]), *g, rank=1)


def test_alias(tmp_path):
    ng = (3, 3, 3)
    f = [halo(np.ones(ng, order="F") * q) for q in range(4)]
    ck.load_all("w", str(tmp_path / "fld_0001.bin"), ng, (1, 1, 1), ng, *f, time=0.5, istep=3)
    ck.gen_alias(str(tmp_path), "fld_0001.bin", "fld.bin")             # main.f90:605
    ck.gen_alias(str(tmp_path), "fld_0001.bin", "fld.bin")             # idempotent
    data, t, i = reference_reader(str(tmp_path / "fld.bin"), ng)
    assert t == 0.5 and i == 3 and np.array_equal(data[:, :, :, 2], f[2][1:-1, 1:-1, 1:-1])


def test_compare_fld_tool_on_oracle_restart_files(tmp_path):
    """tools/compare_fld.py (the off-box pin of INTEGRATION.md) end to end: two restart files of the same oracle run agree; the
    additive constant of the pressure is ignored; a 1e-6 perturbation of one velocity, a different step count or a different
    time are reported as FAIL with a non-zero exit code."""
    import subprocess
    import sys
    import oracle.param as op
    from oracle.main import Sim
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ng = (12, 10, 8)
    s = Sim(op.deck_tgv(ng=ng))
    for _ in range(3):
        s.step()
    f = [getattr(s, nm)[0] for nm in ("U", "V", "W", "P")]

    def write(name, fields, time, istep):
        fn = str(tmp_path / name)
        ck.load_all("w", fn, ng, (1, 1, 1), ng, *fields, time=time, istep=istep)
        return fn

    def run(a, b):
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "compare_fld.py"), a, b, "--ng"] + [str(x) for x in ng],
                           capture_output=True, text=True)
        return r.returncode, r.stdout
    ref = write("ref.bin", f, s.time, s.istep)
    shifted = [f[0], f[1], f[2], f[3] + 123.456]
    rc, out = run(write("same.bin", shifted, s.time, s.istep), ref)
    assert rc == 0 and "PASS" in out, out
    bad = [f[0], f[1] * (1. + 1e-6), f[2], f[3]]
    rc, out = run(write("bad.bin", bad, s.time, s.istep), ref)
    assert rc == 1 and "FAIL" in out and "v: relative L-inf difference" in out, out
    assert run(write("step.bin", f, s.time, s.istep + 1), ref)[0] == 1
    assert run(write("time.bin", f, s.time * 1.001, s.istep), ref)[0] == 1
