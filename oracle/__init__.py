"""CPU oracle for the CaLES per-RK3-substep hot path.

TEST INFRASTRUCTURE ONLY.  This package is a numpy/scipy restatement of the
reference's Fortran algorithm (every function cites the file:line it follows,
paths relative to the reference tree).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it; the product path (``cales_b200``) never
does.

PARITY UNPINNED: the reference (Fortran 2008 + MPI + FFTW, or nvfortran +
OpenACC + cuDecomp) cannot be compiled in this image and ships no golden
vectors or tests for this path, so this oracle is pinned only by
 (i) analytical identities (discrete-Laplacian residual of the solver,
     divergence after projection, Thomas vs dense solve, transform round
     trips, 2-D Taylor-Green decay, polynomial checks of mom_xyz_ad), and
 (ii) the cuDecomp/2decomp transpose-test convention for the index maps
     (payload = global linear index, compared exactly),
 (iii) physics known-answer runs (tests/test_oracle_physics.py: discrete
     Poiseuille steady state and its pressure gradient, duct series solution
     at second order, Smagorinsky / van Driest / dynamic-model closed forms), and
 (iv) a second restatement written independently from the same Fortran lines
     (oracle/c: C + OpenMP; static and dynamic Smagorinsky, tri-periodic and
     plane-channel decks incl. the wall model), which reproduces this one bit
     for bit wherever no transform is involved (tests/test_oracle_c.py).
Transforms use scipy.fft (pocketfft), whose unnormalised rfft/DCT/DST kinds
have the same definitions as the FFTW R2HC/HC2R/REDFT/RODFT kinds the
reference plans (src/fft.f90:192-245).

Array convention: every 3-D field is a Fortran-ordered float64 ndarray of
shape (n1+2, n2+2, n3+2); python index [i,j,k] is the Fortran element
(i,j,k) of an array declared (0:n1+1,0:n2+1,0:n3+1).  Directions and
velocity components are 0-based here (x=0,y=1,z=2); ``cbc[ib,idir,ivel]``
mirrors ``cbcvel(ib,idir+1,ivel+1)``.
"""
