!***************************************************************************************************
! xnet_jacobian_b200.F90
!
! Link-time replacement for xnet_jacobian_dense.F90 (source/Makefile:82-85 selects exactly one file
! that defines `Module xnet_jacobian`).  Same module name, same public procedures, same optional
! keyword arguments; every procedure forwards to libxnet_b200.so through iso_c_binding
! (include/xnet_b200.h).  No OpenACC/HIP/MAGMA/cuBLAS dispatch and no CPU fallback: a failing
! library call ends the run through xnet_terminate, the reference's one fatal path.
!
! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Fortran compiler.  To use it,
!   cp fortran/xnet_jacobian_b200.F90 fortran/xnet_b200f.F90 fortran/xnet_integrate_b200.F90 $XNET/source/
!   add `xnet_b200f.o` to BASE and `MATRIX_SOLVER = b200` to Makefile.opt, link with
!   -L$REPO/xnet_b200 -lxnet_b200 -lcudart      (see INTEGRATION.md)
!
! Run-time switches (environment, read once in read_jacobian_data):
!   XNET_B200_ARITH  = strict | fast (default strict: IEEE operation order and DGETRF's pivots, the mode that
!                      reproduces the CPU build; fast = FMA, CUDA libm, threshold pivoting)
!   XNET_B200_HELM   = path of the Helmholtz table (default helm_table.dat in the run directory, where
!                      actual_eos_init reads it)
! Threads: every call serialises on one launch mutex and one stream inside the library (the contexts of different
! OpenMP threads may share a device); calls on disjoint zone slices are correct from any thread, they do not overlap.
!***************************************************************************************************

Module xnet_jacobian
  Use, Intrinsic :: iso_c_binding
  Use xnet_types, Only: dp
  Use xnet_b200f
  Implicit None
  Real(dp), Allocatable, Target :: jac(:,:,:)     ! host mirror, filled only when idiag >= 5
  Integer, Allocatable, Target :: indx(:,:)       ! host mirror of the LAPACK ipiv (idiag >= 6)
  Integer, Allocatable, Target :: info(:)
  Integer :: msize
  Real(dp), Allocatable, Target :: diag0(:), mult1(:)
  Type(C_PTR), Save :: xb_net = C_NULL_PTR, xb_ctx = C_NULL_PTR

Contains

  Subroutine xb_check(rc,where)
    Use xnet_util, Only: xnet_terminate
    Integer(C_INT), Intent(in) :: rc
    Character(*), Intent(in) :: where
    If ( rc /= 0 ) Call xnet_terminate('xnet_b200: '//where//' failed',int(rc))
  End Subroutine xb_check

  Subroutine read_jacobian_data(data_dir)
    ! Replaces xnet_jacobian_dense.F90:54-119.  Hands the module state that read_nuclear_data /
    ! read_reaction_data / read_ffn_data built to the library (it copies; Fortran keeps ownership).
    Use nuclear_data, Only: ny, aa, zz, mex, be, mex_n, mex_p, g, t9i
    Use reaction_data
    Use xnet_ffn, Only: ffnsum, ffnenu, ffn_ec, ffn_beta, ffn_qval, has_logft
    Use xnet_match, Only: mflx, nflx, iwflx, qflx, ifl1, ifl2, ifl3, ifl4
    Use xnet_controls
    Implicit None
    Character(*), Intent(in) :: data_dir
    Type(xb_net_arrays), Target :: arr
    Type(xb_ctl_t), Target :: ctl
    Character(len=512) :: env
    Integer :: envlen, envstat

    msize = ny
    If ( iheat > 0 ) msize = ny + 1
    Allocate (diag0(nzevolve),mult1(nzevolve),info(nzevolve))
    diag0 = 0.0 ; mult1 = 1.0 ; info = 0
    If ( idiag >= 5 ) Allocate (jac(msize,msize,nzevolve),indx(msize,nzevolve))

    arr%ny = ny ; arr%nffn = nffn ; arr%nnnu = nnnu
    arr%nreac = nreac(1:4) ; arr%nan = nan(1:4)
    arr%rc = [ c_loc(rc1), c_loc(rc2), c_loc(rc3), c_loc(rc4) ]
    arr%ni = [ c_loc(n1i), c_loc(n2i), c_loc(n3i), c_loc(n4i) ]
    arr%ni_ld = [ 5, 6, 6, 6 ]
    arr%iwk = [ c_loc(iwk1), c_loc(iwk2), c_loc(iwk3), c_loc(iwk4) ]
    arr%irev = [ c_loc(irev1), c_loc(irev2), c_loc(irev3), c_loc(irev4) ]
    arr%la = c_loc(la) ; arr%le = c_loc(le)
    arr%mu = [ c_loc(mu1), c_loc(mu2), c_loc(mu3), c_loc(mu4) ]
    arr%a = [ c_loc(a1), c_loc(a2), c_loc(a3), c_loc(a4) ]
    arr%aa = c_loc(aa) ; arr%zz = c_loc(zz) ; arr%mex = c_loc(mex) ; arr%be = c_loc(be)
    arr%mex_n = mex_n ; arr%mex_p = mex_p
    arr%g = c_loc(g) ; arr%t9i = c_loc(t9i)
    arr%ffnsum = c_loc(ffnsum) ; arr%ffnenu = c_loc(ffnenu)
    ! logft weak rates (read_ffn_data_logft, xnet_ffn.F90:96-137): log(ft) and beta-decay tables, Q values, capture type
    If ( abs(iweak0) == 2 ) Then
      arr%ffn_ec = c_loc(ffn_ec) ; arr%ffn_beta = c_loc(ffn_beta)
      arr%ffn_qval = c_loc(ffn_qval) ; arr%has_logft = c_loc(has_logft)
    EndIf
    ! reaction pairs (read_match_data, xnet_match.F90:230-257, runs before read_jacobian_data in net.F90)
    arr%mflx = mflx
    arr%ifl = [ c_loc(ifl1), c_loc(ifl2), c_loc(ifl3), c_loc(ifl4) ]
    arr%nflx = c_loc(nflx) ; arr%iwflx = c_loc(iwflx) ; arr%qflx = c_loc(qflx)
    Call xb_check(xb_net_from_arrays(c_loc(arr),xb_net),'xb_net_from_arrays')

    ! Screening and self-heating need the electron EoS: the library reads the same Helmholtz table that eos_initialize
    ! (net.F90:138 -> actual_eos_init) reads on the host
    If ( iscrn > 0 .or. iheat > 0 ) Then
      Call get_environment_variable('XNET_B200_HELM',env,envlen,envstat)
      If ( envstat /= 0 .or. envlen == 0 ) env = 'helm_table.dat'
      Call xb_check(xb_net_load_helm(xb_net,trim(env)//C_NULL_CHAR),'xb_net_load_helm')
    EndIf

    ctl%iweak0 = iweak0 ; ctl%iscrn = iscrn ; ctl%iheat = iheat ; ctl%isolv = isolv
    ctl%kstmx = kstmx ; ctl%kitmx = kitmx ; ctl%ijac = ijac ; ctl%iconvc = iconvc
    ctl%changemx = changemx ; ctl%changemxt = changemxt ; ctl%yacc = yacc ; ctl%tolm = tolm
    ctl%tolc = tolc ; ctl%tolt9 = tolt9 ; ctl%ymin = ymin ; ctl%tdel_maxmult = tdel_maxmult
    ctl%arith = 1 ; ctl%solver = 0
    Call get_environment_variable('XNET_B200_ARITH',env,envlen,envstat)
    If ( envstat == 0 .and. envlen > 0 ) Then
      If ( env(1:4) == 'fast' ) ctl%arith = 0
    EndIf
    Call xb_check(xb_ctx_create(xb_net,c_loc(ctl),int(nzevolve,C_INT),int(mod(myid,max(1,ngpu)),C_INT),xb_ctx), &
      & 'xb_ctx_create')
    Return
  End Subroutine read_jacobian_data

  Subroutine jacobian_finalize
    Call xb_ctx_destroy(xb_ctx) ; Call xb_net_free(xb_net)
    xb_ctx = C_NULL_PTR ; xb_net = C_NULL_PTR
    Deallocate (diag0,mult1,info)
    If ( allocated(jac) ) Deallocate (jac,indx)
  End Subroutine jacobian_finalize

  Function mask_ptr(mask_in,buf) Result(p)
    ! Fortran LOGICAL(4) -> int32 (non-zero = .true.) without relying on the compiler's bit pattern
    Use xnet_controls, Only: zb_lo, zb_hi, lzactive
    Logical, Optional, Intent(in) :: mask_in(zb_lo:zb_hi)
    Integer(C_INT), Target, Intent(out) :: buf(zb_lo:zb_hi)
    Type(C_PTR) :: p
    If ( present(mask_in) ) Then
      buf = merge(1,0,mask_in)
    Else
      buf = merge(1,0,lzactive(zb_lo:zb_hi))
    EndIf
    p = c_loc(buf)
  End Function mask_ptr

  Subroutine jacobian_build(diag_in,mult_in,mask_in)
    ! Replaces jacobian_build (xnet_jacobian_dense.F90:257-434): dydotdy from the device-resident yt, csect (and, with
    ! iheat > 0, dcsect/cv) of the preceding cross_sect / yderiv calls of the b200 xnet_integrate, kept on the device, then
    ! jacobian_scale with the factors given here.
    Use xnet_controls, Only: zb_lo, zb_hi, idiag
    Real(dp), Optional, Target, Intent(in) :: diag_in(zb_lo:zb_hi), mult_in(zb_lo:zb_hi)
    Logical, Optional, Target, Intent(in) :: mask_in(zb_lo:zb_hi)
    Integer(C_INT), Target :: m(zb_lo:zb_hi)
    Type(C_PTR) :: pd, pm
    pd = C_NULL_PTR ; pm = C_NULL_PTR          ! NULL = diag0 (0.0) / mult1 (1.0), as in the reference
    If ( present(diag_in) ) pd = c_loc(diag_in)
    If ( present(mult_in) ) pm = c_loc(mult_in)
    Call xb_check(xb_jacobian_build(xb_ctx,int(zb_lo,C_INT),int(zb_hi,C_INT),mask_ptr(mask_in,m),pd,pm), &
      & 'xb_jacobian_build')
    If ( idiag >= 5 ) Call xb_check(xb_get_jac(xb_ctx,int(zb_lo,C_INT),int(zb_hi,C_INT),c_loc(jac(1,1,zb_lo))),'xb_get_jac')
  End Subroutine jacobian_build

  Subroutine jacobian_scale(diag_in,mult_in,mask_in)
    ! Replaces jacobian_scale (xnet_jacobian_dense.F90:136-255): jac = mult*dydotdy^T + diag from the dydotdy of the last
    ! build, WITHOUT re-assembly -- BDF refactors with new (diag,mult) while yt and the rates have moved on
    ! (xnet_integrate_bdf.F90:925-927: jacobian_build under `rebuild`, jacobian_scale under `refactor`).
    Use xnet_controls, Only: zb_lo, zb_hi, idiag
    Real(dp), Optional, Target, Intent(in) :: diag_in(zb_lo:zb_hi), mult_in(zb_lo:zb_hi)
    Logical, Optional, Target, Intent(in) :: mask_in(zb_lo:zb_hi)
    Integer(C_INT), Target :: m(zb_lo:zb_hi)
    Type(C_PTR) :: pd, pm
    pd = C_NULL_PTR ; pm = C_NULL_PTR
    If ( present(diag_in) ) pd = c_loc(diag_in)
    If ( present(mult_in) ) pm = c_loc(mult_in)
    Call xb_check(xb_jacobian_scale(xb_ctx,int(zb_lo,C_INT),int(zb_hi,C_INT),mask_ptr(mask_in,m),pd,pm), &
      & 'xb_jacobian_scale')
    If ( idiag >= 5 ) Call xb_check(xb_get_jac(xb_ctx,int(zb_lo,C_INT),int(zb_hi,C_INT),c_loc(jac(1,1,zb_lo))),'xb_get_jac')
  End Subroutine jacobian_scale

  Subroutine jacobian_decomp(kstep,mask_in)
    Use xnet_controls, Only: zb_lo, zb_hi
    Integer, Intent(in) :: kstep
    Logical, Optional, Target, Intent(in) :: mask_in(zb_lo:zb_hi)
    Integer(C_INT), Target :: m(zb_lo:zb_hi)
    Call xb_check(xb_jacobian_decomp(xb_ctx,int(zb_lo,C_INT),int(zb_hi,C_INT),mask_ptr(mask_in,m),c_loc(info(zb_lo))), &
      & 'xb_jacobian_decomp')
  End Subroutine jacobian_decomp

  Subroutine jacobian_bksub(kstep,yrhs,dy,t9rhs,dt9,mask_in)
    Use nuclear_data, Only: ny
    Use xnet_controls, Only: zb_lo, zb_hi
    Integer, Intent(in) :: kstep
    Real(dp), Target, Intent(in) :: yrhs(ny,zb_lo:zb_hi), t9rhs(zb_lo:zb_hi)
    Real(dp), Target, Intent(out) :: dy(ny,zb_lo:zb_hi), dt9(zb_lo:zb_hi)
    Logical, Optional, Target, Intent(in) :: mask_in(zb_lo:zb_hi)
    Integer(C_INT), Target :: m(zb_lo:zb_hi)
    Call xb_check(xb_jacobian_bksub(xb_ctx,int(zb_lo,C_INT),int(zb_hi,C_INT),mask_ptr(mask_in,m), &
      & c_loc(yrhs),c_loc(t9rhs),c_loc(dy),c_loc(dt9)),'xb_jacobian_bksub')
  End Subroutine jacobian_bksub

  Subroutine jacobian_solve(kstep,yrhs,dy,t9rhs,dt9,mask_in)
    Use nuclear_data, Only: ny
    Use xnet_controls, Only: zb_lo, zb_hi
    Integer, Intent(in) :: kstep
    Real(dp), Target, Intent(in) :: yrhs(ny,zb_lo:zb_hi), t9rhs(zb_lo:zb_hi)
    Real(dp), Target, Intent(out) :: dy(ny,zb_lo:zb_hi), dt9(zb_lo:zb_hi)
    Logical, Optional, Target, Intent(in) :: mask_in(zb_lo:zb_hi)
    Integer(C_INT), Target :: m(zb_lo:zb_hi)
    Call xb_check(xb_jacobian_solve(xb_ctx,int(zb_lo,C_INT),int(zb_hi,C_INT),mask_ptr(mask_in,m), &
      & c_loc(yrhs),c_loc(t9rhs),c_loc(dy),c_loc(dt9),c_loc(info(zb_lo))),'xb_jacobian_solve')
  End Subroutine jacobian_solve

  ! -------------------------------------------------------------------------------------------------
  ! Optional fused drop-ins: same argument lists as the reference procedures they stand in for.
  ! step_be_b200 replaces step_be (xnet_integrate_be.F90:233-482): point the call in solve_be
  ! (xnet_integrate_be.F90:98) at it to run the whole Newton loop in one launch.
  Subroutine step_be_b200(kstep,inr)
    Use nuclear_data, Only: ny
    Use xnet_abundances, Only: y, yt, ydot
    Use xnet_conditions, Only: t9, t9t, rhot, tdel, nh, t9dot
    Use xnet_controls, Only: zb_lo, zb_hi, ktot
    Integer, Intent(in) :: kstep
    Integer, Target, Intent(inout) :: inr(zb_lo:zb_hi)
    Integer(C_INT), Target :: kc(5,zb_lo:zb_hi)
    Call xb_check(xb_step_be(xb_ctx,int(zb_lo,C_INT),int(zb_hi,C_INT),int(kstep,C_INT),c_loc(inr), &
      & c_loc(y(1,zb_lo)),c_loc(yt(1,zb_lo)),c_loc(tdel(zb_lo)),c_loc(t9(zb_lo)),c_loc(t9t(zb_lo)),c_loc(rhot(zb_lo)), &
      & c_loc(nh(zb_lo)),c_loc(ydot(1,zb_lo)),c_loc(t9dot(zb_lo)),c_loc(kc)),'xb_step_be')
    ktot(3:5,zb_lo:zb_hi) = ktot(3:5,zb_lo:zb_hi) + kc(3:5,:)   ! Jacobian builds, yderiv and cross_sect calls
  End Subroutine step_be_b200

  ! flux_b200 replaces the summation loop of flux (xnet_flux.F90:104-131); flx_int is accumulated by the caller
  Subroutine flux_b200(flx,mask_in)
    Use xnet_abundances, Only: yt
    Use xnet_controls, Only: zb_lo, zb_hi
    Real(dp), Target, Intent(inout) :: flx(:,:)
    Logical, Optional, Intent(in) :: mask_in(zb_lo:zb_hi)
    Integer(C_INT), Target :: m(zb_lo:zb_hi)
    Call xb_check(xb_flux(xb_ctx,int(zb_lo,C_INT),int(zb_hi,C_INT),mask_ptr(mask_in,m),c_loc(yt(1,zb_lo)),c_loc(flx)), &
      & 'xb_flux')
  End Subroutine flux_b200

End Module xnet_jacobian
