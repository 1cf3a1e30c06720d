!***************************************************************************************************
! xnet_integrate_b200.F90
!
! The rate half of the drop-in boundary (SURVEY 8b): cross_sect and yderiv of Module xnet_integrate
! (source/xnet_integrate.F90:569-984 and :364-567) forwarded to libxnet_b200.so.  The two procedures below
! REPLACE the bodies of the reference's cross_sect / yderiv inside xnet_integrate.F90 (the module also holds
! timestep, update_iweak, ... which stay as they are): same names, same optional mask_in, same module
! variables read (yt, t9t, rhot, xext, aext, zext) and written (yet, ydot, t9dot, cv/etae/detaedt9), same
! counters (ktot(4), ktot(5)).  csect, the screening factors and, with iheat > 0, dcsect*dt9 stay device
! resident for the jacobian_build that follows (xnet_jacobian_b200.F90); xb_get_csect / xb_get_dcsect copy
! them back when the host wants to inspect them (idiag >= 5).
!
! NOT COMPILED IN THIS REPOSITORY'S CI (no Fortran compiler in the build image).  To use: in
! $XNET/source/xnet_integrate.F90 delete the bodies of cross_sect and yderiv and `Include` this file in the
! module's Contains section (or paste the two procedures), build with MATRIX_SOLVER = b200 (INTEGRATION.md).
!***************************************************************************************************

  Subroutine cross_sect(mask_in)
    ! Replaces xnet_integrate.F90:569-984 (update_eos + screening + partf + REACLIB/FFN rates).
    Use, Intrinsic :: iso_c_binding
    Use xnet_abundances, Only: yt, xext, aext, zext
    Use xnet_conditions, Only: t9t, rhot, yet, cv, etae, detaedt9, t9dot
    Use xnet_controls, Only: zb_lo, zb_hi, lzactive, ktot, iheat, iscrn
    Use xnet_jacobian, Only: xb_ctx, xb_check, mask_ptr
    Use xnet_b200f
    Use xnet_timers, Only: xnet_wtime, start_timer, stop_timer, timer_csect
    Implicit None
    Logical, Optional, Target, Intent(in) :: mask_in(zb_lo:zb_hi)
    Integer(C_INT), Target :: m(zb_lo:zb_hi)
    Type(C_PTR) :: pm
    Integer :: izb

    pm = mask_ptr(mask_in,m)
    If ( .not. any(m /= 0) ) Return
    start_timer = xnet_wtime()
    timer_csect = timer_csect - start_timer
    Call xb_check(xb_cross_sect(xb_ctx,int(zb_lo,C_INT),int(zb_hi,C_INT),pm,c_loc(yt(1,zb_lo)),c_loc(t9t(zb_lo)), &
      & c_loc(rhot(zb_lo)),c_loc(xext(zb_lo)),c_loc(aext(zb_lo)),c_loc(zext(zb_lo)),c_loc(yet(zb_lo))),'xb_cross_sect')
    ! update_eos leaves cv / etae / detaedt9 in xnet_conditions (used by timestep's temperature time scale)
    If ( iheat > 0 .or. iscrn > 0 ) Call xb_check(xb_get_eos(xb_ctx,int(zb_lo,C_INT),int(zb_hi,C_INT), &
      & c_loc(cv(zb_lo)),c_loc(etae(zb_lo)),c_loc(detaedt9(zb_lo)),C_NULL_PTR),'xb_get_eos')
    Do izb = zb_lo, zb_hi
      If ( m(izb) /= 0 ) ktot(5,izb) = ktot(5,izb) + 1
    EndDo
    stop_timer = xnet_wtime()
    timer_csect = timer_csect + stop_timer
    Return
  End Subroutine cross_sect

  Subroutine yderiv(mask_in)
    ! Replaces xnet_integrate.F90:364-567 (ydot over the la/le maps; t9dot with iheat > 0).
    Use, Intrinsic :: iso_c_binding
    Use xnet_abundances, Only: yt, ydot
    Use xnet_conditions, Only: t9dot
    Use xnet_controls, Only: zb_lo, zb_hi, lzactive, ktot, iheat
    Use xnet_jacobian, Only: xb_ctx, xb_check, mask_ptr
    Use xnet_b200f
    Use xnet_timers, Only: xnet_wtime, start_timer, stop_timer, timer_deriv
    Implicit None
    Logical, Optional, Target, Intent(in) :: mask_in(zb_lo:zb_hi)
    Integer(C_INT), Target :: m(zb_lo:zb_hi)
    Type(C_PTR) :: pm
    Integer :: izb

    pm = mask_ptr(mask_in,m)
    If ( .not. any(m /= 0) ) Return
    start_timer = xnet_wtime()
    timer_deriv = timer_deriv - start_timer
    ! masked-out zones keep their ydot: the library writes only the zones of the mask
    Call xb_check(xb_yderiv(xb_ctx,int(zb_lo,C_INT),int(zb_hi,C_INT),pm,c_loc(yt(1,zb_lo)),c_loc(ydot(1,zb_lo))),'xb_yderiv')
    If ( iheat > 0 ) Call xb_check(xb_get_eos(xb_ctx,int(zb_lo,C_INT),int(zb_hi,C_INT), &
      & C_NULL_PTR,C_NULL_PTR,C_NULL_PTR,c_loc(t9dot(zb_lo))),'xb_get_eos')
    Do izb = zb_lo, zb_hi
      If ( m(izb) /= 0 ) ktot(4,izb) = ktot(4,izb) + 1
    EndDo
    stop_timer = xnet_wtime()
    timer_deriv = timer_deriv + stop_timer
    Return
  End Subroutine yderiv
