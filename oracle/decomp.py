"""2-D pencil decomposition maps and a serial emulation of the MPI ranks.

Follows dependencies/2decomp-fft/src/decomp_2d.f90:1046-1149 (`partition`, `distribute`),
decomp_2d_init_fin.f90:120-142 (cartesian coordinates: rank = coord(1)*p_col + coord(2)),
src/initmpi.f90:34-206 (lo/hi/n, n_x_fft, n_y_fft, lo_z/hi_z/n_z, nb, is_bound) and
src/bound.f90:619-696 (`updthalo`).  cuDecomp's `cudecompGetPencilInfo`
(cuDecomp/src/cudecomp.cc:776-836) yields the same lo/hi (0-based there, 1-based here)."""
import numpy as np

PROC_NULL = -1


def distribute(data1, proc):
    """decomp_2d.f90:1096-1147 (NEW_DISTRIBUTION branch).  Returns 1-based st, en and sz."""
    sz = [data1 // proc] * proc
    for i in range(1, data1 % proc + 1):
        sz[i - 1] = sz[i - 1] + 1
    st = [0] * proc
    en = [0] * proc
    for i in range(1, proc + 1):
        st[i - 1] = 1 + (i - 1) * sz[i - 1]
        en[i - 1] = 0 + i * sz[i - 1]
    for i in range(data1 % proc + 1, proc + 1):
        st[i - 1] = st[i - 1] + data1 % proc
        en[i - 1] = en[i - 1] + data1 % proc
    return st, en, sz


def partition(ng, pdim, dims, coord):
    """decomp_2d.f90:1046-1094.  pdim(i)=1: local; 2: split over dims(1); 3: split over dims(2)."""
    lstart, lend, lsize = [0] * 3, [0] * 3, [0] * 3
    for i in range(3):
        gsize = ng[i]
        if pdim[i] == 1:
            lstart[i], lend[i], lsize[i] = 1, gsize, gsize
        else:
            a = pdim[i] - 2
            st, en, sz = distribute(gsize, dims[a])
            lstart[i], lend[i], lsize[i] = st[coord[a]], en[coord[a]], sz[coord[a]]
    return lstart, lend, lsize


_PDIM = {1: (1, 2, 3), 2: (2, 1, 3), 3: (2, 3, 1)}     # decomp_2d.f90 get_decomp_info: x-, y-, z-pencil


class Rank:
    pass


class World:
    """All ranks of a run, emulated serially.  `ipencil` in {1,2,3} <- _DECOMP_X/_Y/_Z."""

    def __init__(self, ng, dims, cbcpre, ipencil=1):
        self.ng = tuple(int(x) for x in ng)
        self.dims = tuple(int(x) for x in dims)
        self.ipencil = ipencil
        self.nranks = self.dims[0] * self.dims[1]
        periods = [cbcpre[0, d] + cbcpre[1, d] == "PP" for d in range(3)]     # initmpi.f90:76-77
        self.periods = periods
        ipencil_t = [d for d in (1, 2, 3) if d != ipencil]                    # initmpi.f90:63
        self.ranks = []
        for myid in range(self.nranks):
            r = Rank()
            r.id = myid
            r.coord = (myid // self.dims[1], myid % self.dims[1])             # MPI_CART_COORDS, row-major
            r.xstart, r.xend, r.xsize = partition(self.ng, _PDIM[1], self.dims, r.coord)
            r.ystart, r.yend, r.ysize = partition(self.ng, _PDIM[2], self.dims, r.coord)
            r.zstart, r.zend, r.zsize = partition(self.ng, _PDIM[3], self.dims, r.coord)
            st, en = {1: (r.xstart, r.xend), 2: (r.ystart, r.yend), 3: (r.zstart, r.zend)}[ipencil]
            r.lo = list(st); r.hi = list(en)                                  # initmpi.f90:178-191
            r.n = [r.hi[i] - r.lo[i] + 1 for i in range(3)]
            r.n_x_fft = list(r.xsize); r.n_y_fft = list(r.ysize)
            r.lo_z = list(r.zstart); r.hi_z = list(r.zend); r.n_z = list(r.zsize)
            nb = np.full((2, 3), PROC_NULL, dtype=np.int64)                   # initmpi.f90:201-203
            for a in range(2):                                                # MPI_CART_SHIFT(comm_cart,a,1,...)
                idir = ipencil_t[a] - 1
                for ib, disp in ((0, -1), (1, +1)):
                    c = list(r.coord)
                    c[a] += disp
                    if c[a] < 0 or c[a] >= self.dims[a]:
                        if periods[idir]:
                            c[a] %= self.dims[a]
                        else:
                            continue
                    nb[ib, idir] = c[0] * self.dims[1] + c[1]
            r.nb = nb
            r.is_bound = (nb == PROC_NULL)                                    # initmpi.f90:204
            self.ranks.append(r)

    # -- field helpers -------------------------------------------------------------------
    def zeros(self):
        return [np.zeros((r.n[0] + 2, r.n[1] + 2, r.n[2] + 2), order="F") for r in self.ranks]

    def scatter(self, g):
        """Split a global halo-free array (ng) into per-rank haloed arrays (interior filled)."""
        out = self.zeros()
        for r, a in zip(self.ranks, out):
            a[1:-1, 1:-1, 1:-1] = g[r.lo[0] - 1:r.hi[0], r.lo[1] - 1:r.hi[1], r.lo[2] - 1:r.hi[2]]
        return out

    def gather(self, arrs):
        g = np.zeros(self.ng, order="F")
        for r, a in zip(self.ranks, arrs):
            g[r.lo[0] - 1:r.hi[0], r.lo[1] - 1:r.hi[1], r.lo[2] - 1:r.hi[2]] = a[1:-1, 1:-1, 1:-1]
        return g

    # -- communication -------------------------------------------------------------------
    def updthalo(self, arrs, idir):
        """bound.f90:619-696 with nh=1: full-extent faces (ghost rows of the other directions
        included), nothing along the pencil axis, PROC_NULL neighbours leave ghosts untouched."""
        if idir + 1 == self.ipencil:
            return
        sl = [slice(None)] * 3
        for r in self.ranks:
            p = arrs[r.id]
            n = r.n[idir]
            nb0, nb1 = r.nb[0, idir], r.nb[1, idir]
            if nb1 != PROC_NULL:           # recv p(hi+1) from nb(1), which sends its p(lo)
                src = arrs[nb1]
                d = list(sl); d[idir] = n + 1
                s = list(sl); s[idir] = 1
                p[tuple(d)] = src[tuple(s)]
            if nb0 != PROC_NULL:           # recv p(lo-1) from nb(0), which sends its p(hi)
                src = arrs[nb0]
                d = list(sl); d[idir] = 0
                s = list(sl); s[idir] = self.ranks[nb0].n[idir]
                p[tuple(d)] = src[tuple(s)]

    def updthalo_all(self, arrs):
        for idir in range(3):
            self.updthalo(arrs, idir)

    @staticmethod
    def allreduce_sum(vals):
        """MPI_ALLREDUCE(SUM): accumulated in rank order (one legal order among many)."""
        s = vals[0]
        for v in vals[1:]:
            s = s + v
        return s

    # -- pencil transposes (2decomp transpose_x_to_y etc.): same data, different pencil --------
    def pencil_slices(self, r, which):
        st, en = {"x": (r.xstart, r.xend), "y": (r.ystart, r.yend), "z": (r.zstart, r.zend)}[which]
        return tuple(slice(st[i] - 1, en[i]) for i in range(3))

    def to_pencils(self, g, which):
        return [np.asfortranarray(g[self.pencil_slices(r, which)]) for r in self.ranks]

    def from_pencils(self, arrs, which):
        g = np.zeros(self.ng, order="F")
        for r, a in zip(self.ranks, arrs):
            g[self.pencil_slices(r, which)] = a
        return g

    def transpose(self, arrs, src, dst):
        return self.to_pencils(self.from_pencils(arrs, src), dst)
