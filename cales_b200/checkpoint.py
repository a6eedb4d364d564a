"""Restart files in the reference's `fld.bin` format, so runs of this library and of a site-built CaLES interoperate
(the only route to parity against the real Fortran binary, which cannot be built in this image).

Mirrors `load_all` / `io_field`, src/load.f90:20-187:
  * the file holds u, v, w, p one after the other, each the GLOBAL halo-free array `(ng1, ng2, ng3)` in Fortran order and
    native real(rp) (8-byte) words, followed by `[time, real(istep)]` (load.f90:146-151, 175-178);
  * every rank reads / writes only its own sub-box `lo:hi` of each field (the MPI subarray view of `io_field`,
    load.f90:162-186); here each rank maps the file and touches that sub-box, which gives the same bytes on the shared
    file system of one node;
  * reading checks the file size first and fails like the reference does (load.f90:44-52).
The Python reader shipped with the reference, utils/read_binary_data/python/read_restart_file.py:44-58, reads the files
written here unchanged (tests/test_checkpoint.py re-states it).

Host code only: fields are haloed Fortran-ordered numpy arrays `(0:n1+1, 0:n2+1, 0:n3+1)` (nh = 1, main.f90:365, 609)."""
import os

import numpy as np

RP = np.dtype("float64")          # real(rp), precision.f90:14-20 (double precision build)


class CheckpointError(RuntimeError):
    pass


def expected_size(ng):
    """(product(ng)*4 + 2) * sizeof(real(rp))   (load.f90:45)."""
    return (int(ng[0]) * int(ng[1]) * int(ng[2]) * 4 + 2) * RP.itemsize


def _box(ng, lo, hi):
    n = [int(hi[q]) - int(lo[q]) + 1 for q in range(3)]
    for q in range(3):
        if lo[q] < 1 or hi[q] > ng[q] or n[q] < 1:
            raise CheckpointError("sub-box lo=%s hi=%s outside the global grid %s" % (list(lo), list(hi), list(ng)))
    return n, tuple(slice(int(lo[q]) - 1, int(hi[q])) for q in range(3))


def load_all(io, filename, ng, lo, hi, u, v, w, p, time=0.0, istep=0, rank=0, nh=(1, 1, 1), barrier=None):
    """`load_all(io, filename, comm, ng, nh, lo, hi, u, v, w, p, time, istep)`, src/load.f90:20.

    io = 'r': fills the interiors of u, v, w, p (in place) and returns (time, istep);
    io = 'w': writes them; rank 0 creates the file and writes `[time, istep]`; `barrier` (a callable, e.g.
    torch.distributed.barrier) separates the creation of the file from the other ranks' writes -- MPI_FILE_OPEN is
    collective in the reference.  Returns (time, istep)."""
    ng = [int(x) for x in ng]
    n, box = _box(ng, lo, hi)
    nglob = ng[0] * ng[1] * ng[2]
    inner = tuple(slice(int(nh[q]), int(nh[q]) + n[q]) for q in range(3))
    flds = (u, v, w, p)
    for a in flds:
        if tuple(a.shape) != tuple(n[q] + 2 * int(nh[q]) for q in range(3)):
            raise CheckpointError("field of shape %s does not match lo/hi/nh (%s expected)" %
                                  (tuple(a.shape), tuple(n[q] + 2 * int(nh[q]) for q in range(3))))
    good = expected_size(ng)
    if io == "r":
        if not os.path.exists(filename):
            raise CheckpointError("checkpoint file %s not found" % filename)
        size = os.path.getsize(filename)
        if size != good:                                                   # load.f90:46-52
            raise CheckpointError("*** Simulation aborted due a checkpoint file with incorrect size ***\n"
                                  "    file: %s | expected size: %d | actual size: %d" % (filename, good, size))
        mm = np.memmap(filename, dtype=RP, mode="r")
        for q, a in enumerate(flds):
            g = mm[q * nglob:(q + 1) * nglob].reshape(ng, order="F")
            a[inner] = g[box]
        time, istep = float(mm[4 * nglob]), int(np.rint(mm[4 * nglob + 1]))   # istep = nint(fldinfo(2)), load.f90:101
        del mm
        return time, istep
    if io != "w":
        raise CheckpointError("io must be 'r' or 'w'")
    if barrier is None and (rank != 0 or any(int(lo[q]) != 1 or int(hi[q]) != ng[q] for q in range(3))):
        # this rank holds a sub-box: without a barrier nothing orders rank 0's create-and-truncate against the other ranks' writes
        raise CheckpointError("writing a checkpoint from several ranks needs a barrier (MPI_FILE_OPEN is collective, load.f90:106-109)")
    if rank == 0:
        with open(filename, "wb") as f:                                    # MPI_MODE_CREATE + set size 0 (load.f90:106-109)
            f.truncate(good)
    if barrier is not None:
        barrier()
    mm = np.memmap(filename, dtype=RP, mode="r+")
    if mm.size * RP.itemsize != good:
        raise CheckpointError("checkpoint file %s has the wrong size for ng=%s" % (filename, ng))
    for q, a in enumerate(flds):
        g = mm[q * nglob:(q + 1) * nglob].reshape(ng, order="F")
        g[box] = a[inner]
    if rank == 0:
        mm[4 * nglob] = time                                               # fldinfo = [time, 1._rp*istep], load.f90:147
        mm[4 * nglob + 1] = float(istep)
    mm.flush()
    del mm
    if barrier is not None:
        barrier()
    return time, istep


def gen_alias(datadir, filename, alias):
    """`gen_alias`, src/utils.f90 (main.f90:605): a symbolic link `alias` -> `filename` next to it (rank 0 only)."""
    dst = os.path.join(datadir, alias)
    if os.path.lexists(dst):
        os.remove(dst)
    os.symlink(filename, dst)
