! Drop-in wrappers: the hot-path procedures of CaLES with their ORIGINAL Fortran signatures, implemented by
! libcales_b200.so through the bind(C) interfaces of cales_b200_c.f90 (generated from include/cales_b200.h).
! `program cans` (src/main.f90) only changes its `use` lines (INTEGRATION.md section 3); everything it allocates,
! maps to the device (`!$acc enter data`) and passes stays as it is.
!
!   replaced procedure (reference file:line)              wrapper below
!   initmpi            src/initmpi.f90:34                 initmpi   (+ cales_b200_start for the NCCL id / stream)
!   initsolver, fftend src/initsolver.f90:17, fft.f90:145  initsolver, fftend
!   rk                 src/rk.f90:17                      rk
!   bulk_forcing       src/mom.f90:311                    bulk_forcing
!   bulk_mean          src/utils.f90:16                   bulk_mean
!   bounduvw, boundp   src/bound.f90:18, 156              bounduvw, boundp
!   cmpt_rhs_b         src/bound.f90:447                  cmpt_rhs_b
!   updt_rhs_b         src/bound.f90:562                  updt_rhs_b
!   fillps             src/fillps.f90:14                  fillps
!   solver             src/solver.f90:20                  solver
!   solver_gaussel_z   src/solver.f90:182                 solver_gaussel_z
!   correc             src/correc.f90:14                  correc
!   updatep            src/updatep.f90:14                 updatep
!   cmpt_sgs           src/sgs.f90:21                     cmpt_sgs
!   chkdt, chkdiv      src/chkdt.f90:17, chkdiv.f90:16    chkdt, chkdiv
!
! Conventions (include/cales_b200.h:10-29): 3-D fields, grid vectors, `bound` planes, lambdaxy/a/b/c of `solver` are
! DEVICE arrays -> passed as c_loc() inside `host_data use_device`; small descriptor vectors are host arrays; logicals
! travel as 0/1 integers; (0:1,3) tables are passed in Fortran order (reshape to rank 1).  Status codes are checked by
! `chk`, which stops with cales_last_error() -- the reference's hot-path routines return nothing and ignore `istat`.
!
! This file cannot be compiled in the development image (no Fortran compiler); it is kept in step with the C ABI by
! construction (the interface module is generated) and by tests/test_fortran_iface.py (every wrapper calls an existing
! export with the right number of arguments).
module cales_b200_iface
  use, intrinsic :: iso_c_binding
  use cales_b200_c
  use mod_precision, only: rp
  use mod_typedef  , only: bound
  implicit none
  private
  public :: cales_b200_start,cales_b200_stop,initmpi,initsolver,fftend,rk,bulk_forcing,bulk_mean,bounduvw,boundp, &
            cmpt_rhs_b,updt_rhs_b,fillps,solver,solver_gaussel_z,correc,updatep,cmpt_sgs,chkdt,chkdiv
  type(c_ptr), save :: ctx = c_null_ptr
  ! set by cales_b200_start before initmpi is called
  character(kind=c_char), save :: nccl_uid(CALES_UNIQUE_ID_BYTES)
  type(c_ptr)   , save :: stream = c_null_ptr
  integer(c_int), save :: myid_c = 0, nproc_c = 1, ipencil_c = 1, diffusion_c = CALES_DIFF_EXPLICIT
contains
  subroutine chk(istat,what)
    integer(c_int)  , intent(in) :: istat
    character(len=*), intent(in) :: what
    character(kind=c_char), pointer :: msg(:)
    integer :: i
    if(istat == CALES_OK) return
    call c_f_pointer(cales_last_error(ctx),msg,[512])
    write(*,'(3a,i0,a)',advance='no') 'cales_b200: ',what,' failed (',istat,'): '
    do i = 1,512
      if(msg(i) == c_null_char) exit
      write(*,'(a)',advance='no') msg(i)
    end do
    write(*,*)
    error stop
  end subroutine chk
  pure function l2i(l) result(i)
    logical, intent(in) :: l(:)
    integer(c_int) :: i(size(l))
    i = merge(1_c_int,0_c_int,l)
  end function l2i
  function dev_bound(x,y,z) result(cb)
    ! device addresses of the three planes of a `bound` (src/typedef.f90:10-14).  The planes arrive as plain dummy arrays, as
    ! in the reference's own set_bc(...,bcu%x,...) calls (src/bound.f90:58), so that host_data applies to them here.
    real(rp), intent(in), dimension(:,:,:), contiguous, target :: x,y,z
    type(cales_bound) :: cb
    !$acc host_data use_device(x,y,z)
    cb%x = c_loc(x); cb%y = c_loc(y); cb%z = c_loc(z)
    !$acc end host_data
  end function dev_bound
  !
  ! ---- start-up: NCCL id broadcast (the bootstrap of cuDecomp, dependencies/cuDecomp/src/cudecomp.cc:66-80) ---------
  subroutine cales_b200_start(myid,nproc,ipencil,is_impdiff,is_impdiff_1d,cuda_stream)
    use mpi
    integer    , intent(in) :: myid,nproc,ipencil
    logical    , intent(in) :: is_impdiff,is_impdiff_1d
    type(c_ptr), intent(in) :: cuda_stream            ! acc_get_cuda_stream(1): the reference's OpenACC queue 1
    integer :: ierr
    myid_c = myid; nproc_c = nproc; ipencil_c = ipencil; stream = cuda_stream
    diffusion_c = merge(merge(CALES_DIFF_IMPLICIT_1D,CALES_DIFF_IMPLICIT_3D,is_impdiff_1d),CALES_DIFF_EXPLICIT,is_impdiff)
    nccl_uid(:) = c_null_char
    if(nproc > 1) then
      if(myid == 0) call chk(cales_get_unique_id(nccl_uid),'cales_get_unique_id')
      call MPI_BCAST(nccl_uid,CALES_UNIQUE_ID_BYTES,MPI_CHARACTER,0,MPI_COMM_WORLD,ierr)
    end if
  end subroutine cales_b200_start
  subroutine cales_b200_stop()
    if(c_associated(ctx)) call chk(cales_finalize(ctx),'cales_finalize')
    ctx = c_null_ptr
  end subroutine cales_b200_stop
  !
  subroutine initmpi(ng,dims,sgstype,cbcvel,cbcpre,lo,hi,n,n_x_fft,n_y_fft,lo_z,hi_z,n_z,nb,is_bound)   ! src/initmpi.f90:34
    integer         , intent(in   ), dimension(3)       :: ng
    integer         , intent(inout), dimension(2)       :: dims
    character(len=*), intent(in   )                     :: sgstype
    character(len=1), intent(in   ), dimension(0:1,3,3) :: cbcvel
    character(len=1), intent(in   ), dimension(0:1,3)   :: cbcpre
    integer         , intent(out  ), dimension(3)       :: lo,hi,n,n_x_fft,n_y_fft,lo_z,hi_z,n_z
    integer         , intent(out  ), dimension(0:1,3)   :: nb
    logical         , intent(out  ), dimension(0:1,3)   :: is_bound
    integer(c_int) :: nb_c(6),isb_c(6)
    character(kind=c_char) :: cbc_c(6)
    cbc_c = reshape(cbcpre,[6])
    ! device = -1: the library picks local rank modulo device count, as cuDecomp's users do
    call chk(cales_init(ctx,ng,dims,ipencil_c,cbc_c,myid_c,nproc_c,nccl_uid,-1_c_int,stream,diffusion_c),'cales_init')
    call chk(cales_get_decomp(ctx,lo,hi,n,n_x_fft,n_y_fft,lo_z,hi_z,n_z,nb_c,isb_c),'cales_get_decomp')
    nb       = reshape(nb_c,[2,3])
    is_bound = reshape(isb_c == 1,[2,3])
  end subroutine initmpi
  !
  subroutine initsolver(ng,n_x_fft,n_y_fft,lo_z,hi_z,dli,dzci,dzfi,cbc,lambdaxy,c_or_f,a,b,c,arrplan,normfft)
    ! src/initsolver.f90:17
    integer , intent(in), dimension(3) :: ng,n_x_fft,n_y_fft,lo_z,hi_z
    real(rp), intent(in), dimension(3 ) :: dli
    real(rp), intent(in), dimension(0:), target :: dzci,dzfi           ! host copies (global vectors 0:ng(3)+1)
    character(len=1), intent(in), dimension(0:1,3) :: cbc
    real(rp), intent(out), dimension(lo_z(1):,lo_z(2):), target :: lambdaxy
    character(len=1), intent(in), dimension(3) :: c_or_f
    real(rp), intent(out), dimension(:), target :: a,b,c
    integer , intent(out), dimension(2,2) :: arrplan                   ! arrplan(1,1) carries the library's plan handle
    real(rp), intent(out), target :: normfft
    integer(c_int) :: plan(1)
    character(kind=c_char) :: cbc_c(6),cf_c(3)
    cbc_c = reshape(cbc,[6]); cf_c = c_or_f
    call chk(cales_initsolver(ctx,ng,n_x_fft,n_y_fft,lo_z,hi_z,dli,c_loc(dzci),c_loc(dzfi),cbc_c,cf_c, &
                              c_loc(lambdaxy),c_loc(a),c_loc(b),c_loc(c),plan,c_loc(normfft)),'cales_initsolver')
    arrplan(:,:) = plan(1)
  end subroutine initsolver
  subroutine fftend(arrplan)                                                                            ! src/fft.f90:145
    integer, intent(in), dimension(:,:) :: arrplan
    call chk(cales_fftend(ctx,int(arrplan(1,1),c_int)),'cales_fftend')
  end subroutine fftend
  !
  subroutine rk(rkpar,n,dli,dzci,dzfi,grid_vol_ratio_c,grid_vol_ratio_f,visc,dt,p, &
                is_forced,velf,bforce,visct,u,v,w,f)                                                   ! src/rk.f90:17
    real(rp), intent(in), dimension(2) :: rkpar
    integer , intent(in), dimension(3) :: n
    real(rp), intent(in), dimension(3) :: dli
    real(rp), intent(in), dimension(0:), target :: dzci,dzfi
    real(rp), intent(in), dimension(0:), target :: grid_vol_ratio_c,grid_vol_ratio_f
    real(rp), intent(in) :: visc,dt
    real(rp), intent(in), dimension(0:,0:,0:), target :: p
    logical , intent(in), dimension(3)        :: is_forced
    real(rp), intent(in), dimension(3)        :: velf,bforce
    real(rp), intent(in), dimension(0:,0:,0:), target :: visct
    real(rp), intent(inout), dimension(0:,0:,0:), target :: u,v,w
    real(rp), intent(out), dimension(3) :: f
    !$acc host_data use_device(dzci,dzfi,grid_vol_ratio_c,grid_vol_ratio_f,p,visct,u,v,w)
    call chk(cales_rk(ctx,rkpar,n,dli,c_loc(dzci),c_loc(dzfi),c_loc(grid_vol_ratio_c),c_loc(grid_vol_ratio_f),visc,dt, &
                      c_loc(p),l2i(is_forced),velf,bforce,c_loc(visct),c_loc(u),c_loc(v),c_loc(w),f),'cales_rk')
    !$acc end host_data
  end subroutine rk
  subroutine bulk_forcing(n,is_forced,f,u,v,w)                                                          ! src/mom.f90:311
    integer , intent(in   ), dimension(3) :: n
    logical , intent(in   ), dimension(3) :: is_forced
    real(rp), intent(in   ), dimension(3) :: f
    real(rp), intent(inout), dimension(0:,0:,0:), target :: u,v,w
    !$acc host_data use_device(u,v,w)
    call chk(cales_bulk_forcing(ctx,n,l2i(is_forced),f,c_loc(u),c_loc(v),c_loc(w)),'cales_bulk_forcing')
    !$acc end host_data
  end subroutine bulk_forcing
  subroutine bulk_mean(n,grid_vol_ratio,p,mean)                                                         ! src/utils.f90:16
    integer , intent(in), dimension(3) :: n
    real(rp), intent(in), dimension(0:), target :: grid_vol_ratio
    real(rp), intent(in), dimension(0:,0:,0:), target :: p
    real(rp), intent(out), target :: mean
    !$acc host_data use_device(grid_vol_ratio,p)
    call chk(cales_bulk_mean(ctx,n,c_loc(grid_vol_ratio),c_loc(p),c_loc(mean)),'cales_bulk_mean')
    !$acc end host_data
  end subroutine bulk_mean
  !
  subroutine bounduvw(cbc,n,bcu,bcv,bcw,bcu_mag,bcv_mag,bcw_mag,nb,is_bound,lwm,l,dl,zc,zf,dzc,dzf, &
                      visc,h,index_wm,is_updt_wm,is_correc,u,v,w)                                      ! src/bound.f90:18
    character(len=1), intent(in), dimension(0:1,3,3) :: cbc
    integer         , intent(in), dimension(3) :: n
    type(bound)     , intent(inout) :: bcu,bcv,bcw
    type(bound)     , intent(in) :: bcu_mag,bcv_mag,bcw_mag
    integer , intent(in), dimension(0:1,3) :: nb
    logical , intent(in), dimension(0:1,3) :: is_bound
    integer , intent(in), dimension(0:1,3) :: lwm,index_wm
    real(rp), intent(in), dimension(3) :: l,dl
    real(rp), intent(in), dimension(0:), target :: zc,zf,dzc,dzf
    real(rp), intent(in) :: visc,h
    logical , intent(in) :: is_updt_wm,is_correc
    real(rp), intent(inout), dimension(0:,0:,0:), target :: u,v,w
    character(kind=c_char) :: cbc_c(18)
    type(cales_bound) :: du,dv,dw,dum,dvm,dwm
    cbc_c = reshape(cbc,[18])
    du  = dev_bound(bcu%x,bcu%y,bcu%z); dv  = dev_bound(bcv%x,bcv%y,bcv%z); dw  = dev_bound(bcw%x,bcw%y,bcw%z)
    dum = dev_bound(bcu_mag%x,bcu_mag%y,bcu_mag%z); dvm = dev_bound(bcv_mag%x,bcv_mag%y,bcv_mag%z)
    dwm = dev_bound(bcw_mag%x,bcw_mag%y,bcw_mag%z)
    !$acc host_data use_device(zc,zf,dzc,dzf,u,v,w)
    call chk(cales_bounduvw(ctx,cbc_c,n,du,dv,dw,dum,dvm,dwm,reshape(nb,[6]),l2i(reshape(is_bound,[6])),reshape(lwm,[6]),l,dl, &
                            c_loc(zc),c_loc(zf),c_loc(dzc),c_loc(dzf),visc,h,reshape(index_wm,[6]), &
                            merge(1_c_int,0_c_int,is_updt_wm),merge(1_c_int,0_c_int,is_correc), &
                            c_loc(u),c_loc(v),c_loc(w)),'cales_bounduvw')
    !$acc end host_data
  end subroutine bounduvw
  subroutine boundp(cbc,n,bcp,nb,is_bound,dl,dzc,p)                                                    ! src/bound.f90:156
    character(len=1), intent(in), dimension(0:1,3) :: cbc
    integer         , intent(in), dimension(3) :: n
    type(bound)     , intent(in) :: bcp
    integer , intent(in), dimension(0:1,3) :: nb
    logical , intent(in), dimension(0:1,3) :: is_bound
    real(rp), intent(in), dimension(3 ) :: dl
    real(rp), intent(in), dimension(0:), target :: dzc
    real(rp), intent(inout), dimension(0:,0:,0:), target :: p
    character(kind=c_char) :: cbc_c(6)
    type(cales_bound) :: dp
    cbc_c = reshape(cbc,[6])
    dp = dev_bound(bcp%x,bcp%y,bcp%z)
    !$acc host_data use_device(dzc,p)
    call chk(cales_boundp(ctx,cbc_c,n,dp,reshape(nb,[6]),l2i(reshape(is_bound,[6])),dl,c_loc(dzc),c_loc(p)),'cales_boundp')
    !$acc end host_data
  end subroutine boundp
  !
  subroutine cmpt_rhs_b(ng,dl,dzc,dzf,cbc,bc,c_or_f,rhsbx,rhsby,rhsbz)                                  ! src/bound.f90:447
    ! dzc,dzf: the GLOBAL vectors (0:ng(3)+1) on the host, as in the reference's call (main.f90:317, 425-477);
    ! n is taken from the context's decomposition
    integer , intent(in), dimension(3) :: ng
    real(rp), intent(in), dimension(3 ) :: dl
    real(rp), intent(in), dimension(0:), target :: dzc,dzf
    character(len=1), intent(in), dimension(0:1,3) :: cbc
    type(bound)     , intent(in) :: bc
    character(len=1), intent(in), dimension(3) :: c_or_f
    real(rp), intent(out), dimension(:,:,0:), optional, target :: rhsbx
    real(rp), intent(out), dimension(:,:,0:), optional, target :: rhsby
    real(rp), intent(out), dimension(:,:,0:), optional, target :: rhsbz
    integer(c_int) :: lo(3),hi(3),n(3),nxf(3),nyf(3),loz(3),hiz(3),nz(3),nb_c(6),isb_c(6)
    type(c_ptr) :: px,py,pz
    type(cales_bound) :: db
    character(kind=c_char) :: cbc_c(6),cf_c(3)
    cbc_c = reshape(cbc,[6]); cf_c = c_or_f
    db = dev_bound(bc%x,bc%y,bc%z)
    call chk(cales_get_decomp(ctx,lo,hi,n,nxf,nyf,loz,hiz,nz,nb_c,isb_c),'cales_get_decomp')
    px = c_null_ptr; py = c_null_ptr; pz = c_null_ptr
    !$acc host_data use_device(rhsbx,rhsby,rhsbz) if_present
    if(present(rhsbx)) px = c_loc(rhsbx)
    if(present(rhsby)) py = c_loc(rhsby)
    if(present(rhsbz)) pz = c_loc(rhsbz)
    call chk(cales_cmpt_rhs_b(ctx,ng,n,dl,c_loc(dzc),c_loc(dzf),cbc_c,db,cf_c,px,py,pz),'cales_cmpt_rhs_b')
    !$acc end host_data
  end subroutine cmpt_rhs_b
  subroutine updt_rhs_b(c_or_f,cbc,n,is_bound,rhsbx,rhsby,rhsbz,p)                                      ! src/bound.f90:562
    character(len=1), intent(in), dimension(3    ) :: c_or_f
    character(len=1), intent(in), dimension(0:1,3) :: cbc
    integer , intent(in), dimension(3) :: n
    logical , intent(in), dimension(0:1,3) :: is_bound
    real(rp), intent(in), dimension(:,:,0:), optional, target :: rhsbx,rhsby,rhsbz
    real(rp), intent(inout), dimension(0:,0:,0:), target :: p
    type(c_ptr) :: px,py,pz
    character(kind=c_char) :: cbc_c(6),cf_c(3)
    cbc_c = reshape(cbc,[6]); cf_c = c_or_f
    px = c_null_ptr; py = c_null_ptr; pz = c_null_ptr
    !$acc host_data use_device(rhsbx,rhsby,rhsbz,p) if_present
    if(present(rhsbx)) px = c_loc(rhsbx)
    if(present(rhsby)) py = c_loc(rhsby)
    if(present(rhsbz)) pz = c_loc(rhsbz)
    call chk(cales_updt_rhs_b(ctx,cf_c,cbc_c,n,l2i(reshape(is_bound,[6])),px,py,pz,c_loc(p)),'cales_updt_rhs_b')
    !$acc end host_data
  end subroutine updt_rhs_b
  !
  subroutine fillps(n,dli,dzfi,dti,u,v,w,p)                                                             ! src/fillps.f90:14
    integer , intent(in ), dimension(3) :: n
    real(rp), intent(in ), dimension(3 ) :: dli
    real(rp), intent(in ), dimension(0:), target :: dzfi
    real(rp), intent(in ) :: dti
    real(rp), intent(in ), dimension(0:,0:,0:), target :: u,v,w
    real(rp), intent(out), dimension(0:,0:,0:), target :: p
    !$acc host_data use_device(dzfi,u,v,w,p)
    call chk(cales_fillps(ctx,n,dli,c_loc(dzfi),dti,c_loc(u),c_loc(v),c_loc(w),c_loc(p)),'cales_fillps')
    !$acc end host_data
  end subroutine fillps
  subroutine solver(n,ng,arrplan,normfft,lambdaxy,a,b,c,bc,c_or_f,p)                                    ! src/solver.f90:20
    integer , intent(in), dimension(3) :: n,ng
    integer , intent(in), dimension(2,2) :: arrplan
    real(rp), intent(in) :: normfft
    real(rp), intent(in), dimension(:,:), target :: lambdaxy
    real(rp), intent(in), dimension(:), target :: a,b,c
    character(len=1), dimension(0:1,3), intent(in) :: bc
    character(len=1), intent(in), dimension(3) :: c_or_f
    real(rp), intent(inout), dimension(0:,0:,0:), target :: p
    character(kind=c_char) :: bc_c(6),cf_c(3)
    bc_c = reshape(bc,[6]); cf_c = c_or_f
    !$acc host_data use_device(lambdaxy,a,b,c,p)
    call chk(cales_solver(ctx,n,ng,int(arrplan(1,1),c_int),normfft,c_loc(lambdaxy),c_loc(a),c_loc(b),c_loc(c),bc_c,cf_c, &
                          c_loc(p)),'cales_solver')
    !$acc end host_data
  end subroutine solver
  subroutine solver_gaussel_z(n,a,b,c,bcz,c_or_f,p)                                                     ! src/solver.f90:182
    integer , intent(in), dimension(3) :: n
    real(rp), intent(in), dimension(:), target :: a,b,c
    character(len=1), dimension(0:1), intent(in) :: bcz
    character(len=1), intent(in), dimension(3) :: c_or_f
    real(rp), intent(inout), dimension(0:,0:,0:), target :: p
    character(kind=c_char) :: bcz_c(2),cf_c(3)
    bcz_c = bcz; cf_c = c_or_f
    !$acc host_data use_device(a,b,c,p)
    call chk(cales_solver_gaussel_z(ctx,n,c_loc(a),c_loc(b),c_loc(c),bcz_c,cf_c,c_loc(p)),'cales_solver_gaussel_z')
    !$acc end host_data
  end subroutine solver_gaussel_z
  subroutine correc(n,dli,dzci,dt,p,u,v,w)                                                              ! src/correc.f90:14
    integer , intent(in), dimension(3) :: n
    real(rp), intent(in), dimension(3 ) :: dli
    real(rp), intent(in), dimension(0:), target :: dzci
    real(rp), intent(in) :: dt
    real(rp), intent(in   ), dimension(0:,0:,0:), target :: p
    real(rp), intent(inout), dimension(0:,0:,0:), target :: u,v,w
    !$acc host_data use_device(dzci,p,u,v,w)
    call chk(cales_correc(ctx,n,dli,c_loc(dzci),dt,c_loc(p),c_loc(u),c_loc(v),c_loc(w)),'cales_correc')
    !$acc end host_data
  end subroutine correc
  subroutine updatep(n,dli,dzci,dzfi,alpha,pp,p)                                                        ! src/updatep.f90:14
    integer , intent(in   ), dimension(3) :: n
    real(rp), intent(in   ), dimension(3 ) :: dli
    real(rp), intent(in   ), dimension(0:), target :: dzci,dzfi
    real(rp), intent(in   ) :: alpha
    real(rp), intent(in   ), dimension(0:,0:,0:), target :: pp
    real(rp), intent(inout), dimension(0:,0:,0:), target :: p
    !$acc host_data use_device(dzci,dzfi,pp,p)
    call chk(cales_updatep(ctx,n,dli,c_loc(dzci),c_loc(dzfi),alpha,c_loc(pp),c_loc(p)),'cales_updatep')
    !$acc end host_data
  end subroutine updatep
  !
  subroutine cmpt_sgs(sgstype,n,ng,lo,hi,cbcvel,cbcsgs,bcs,nb,is_bound,lwm,l,dl,dli,zc,zf,dzc,dzf, &
                      dzci,dzfi,visc,h,index_wm,u,v,w,bcuf,bcvf,bcwf,bcu_mag,bcv_mag,bcw_mag,visct)     ! src/sgs.f90:21
    character(len=*), intent(in) :: sgstype
    integer , intent(in ), dimension(3) :: n,ng,lo,hi
    character(len=1), intent(in), dimension(0:1,3,3) :: cbcvel
    character(len=1), intent(in), dimension(0:1,3)   :: cbcsgs
    type(bound), intent(in   ) :: bcs
    type(bound), intent(inout) :: bcuf,bcvf,bcwf
    type(bound), intent(in   ) :: bcu_mag,bcv_mag,bcw_mag
    integer , intent(in ), dimension(0:1,3)      :: nb,lwm,index_wm
    logical , intent(in ), dimension(0:1,3)      :: is_bound
    real(rp), intent(in ), dimension(3)          :: l,dl,dli
    real(rp), intent(in ), dimension(0:), target :: zc,zf,dzc,dzf,dzci,dzfi
    real(rp), intent(in )                        :: visc,h
    real(rp), intent(in ), dimension(0:,0:,0:), target :: u,v,w
    real(rp), intent(out), dimension(0:,0:,0:), target :: visct
    character(kind=c_char) :: name_c(len_trim(sgstype)+1),cbcvel_c(18),cbcsgs_c(6)
    type(cales_bound) :: ds,duf,dvf,dwf,dum,dvm,dwm
    integer :: i
    do i = 1,len_trim(sgstype)
      name_c(i) = sgstype(i:i)
    end do
    name_c(len_trim(sgstype)+1) = c_null_char
    cbcvel_c = reshape(cbcvel,[18]); cbcsgs_c = reshape(cbcsgs,[6])
    ds  = dev_bound(bcs%x,bcs%y,bcs%z)
    duf = dev_bound(bcuf%x,bcuf%y,bcuf%z); dvf = dev_bound(bcvf%x,bcvf%y,bcvf%z); dwf = dev_bound(bcwf%x,bcwf%y,bcwf%z)
    dum = dev_bound(bcu_mag%x,bcu_mag%y,bcu_mag%z); dvm = dev_bound(bcv_mag%x,bcv_mag%y,bcv_mag%z)
    dwm = dev_bound(bcw_mag%x,bcw_mag%y,bcw_mag%z)
    !$acc host_data use_device(zc,zf,dzc,dzf,dzci,dzfi,u,v,w,visct)
    call chk(cales_cmpt_sgs(ctx,name_c,n,ng,lo,hi,cbcvel_c,cbcsgs_c,ds,reshape(nb,[6]), &
                            l2i(reshape(is_bound,[6])),reshape(lwm,[6]),l,dl,dli,c_loc(zc),c_loc(zf),c_loc(dzc),c_loc(dzf), &
                            c_loc(dzci),c_loc(dzfi),visc,h,reshape(index_wm,[6]),c_loc(u),c_loc(v),c_loc(w), &
                            duf,dvf,dwf,dum,dvm,dwm, &
                            c_loc(visct)),'cales_cmpt_sgs')
    !$acc end host_data
  end subroutine cmpt_sgs
  !
  subroutine chkdt(n,dl,dzci,dzfi,visc,visct,u,v,w,dtmax)                                               ! src/chkdt.f90:17
    integer , intent(in), dimension(3) :: n
    real(rp), intent(in), dimension(3) :: dl
    real(rp), intent(in), dimension(0:), target :: dzci,dzfi
    real(rp), intent(in) :: visc
    real(rp), intent(in), dimension(0:,0:,0:), target :: visct,u,v,w
    real(rp), intent(out), target :: dtmax
    !$acc host_data use_device(dzci,dzfi,visct,u,v,w)
    call chk(cales_chkdt(ctx,n,dl,c_loc(dzci),c_loc(dzfi),visc,c_loc(visct),c_loc(u),c_loc(v),c_loc(w),c_loc(dtmax)), &
             'cales_chkdt')
    !$acc end host_data
  end subroutine chkdt
  subroutine chkdiv(lo,hi,dli,dzfi,u,v,w,divtot,divmax)                                                 ! src/chkdiv.f90:16
    integer , intent(in), dimension(3) :: lo,hi
    real(rp), intent(in), dimension(3) :: dli
    real(rp), intent(in), dimension(lo(3)-1:), target :: dzfi
    real(rp), intent(in), dimension(lo(1)-1:,lo(2)-1:,lo(3)-1:), target :: u,v,w
    real(rp), intent(out), target :: divtot,divmax
    !$acc host_data use_device(dzfi,u,v,w)
    call chk(cales_chkdiv(ctx,lo,hi,dli,c_loc(dzfi),c_loc(u),c_loc(v),c_loc(w),c_loc(divtot),c_loc(divmax)),'cales_chkdiv')
    !$acc end host_data
  end subroutine chkdiv
end module cales_b200_iface
