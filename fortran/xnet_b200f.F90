!***************************************************************************************************
! xnet_b200f.F90 -- iso_c_binding interface blocks for libxnet_b200.so (include/xnet_b200.h), in the
! idiom of the reference's own vendor bindings (source/cublasf.F90:252-263: `type(c_ptr), value`,
! `integer(c_int), value`, bind(c,name=...)).  NOT compiled in this repository (no Fortran compiler
! in the build image); it is the reference-side half of the drop-in boundary.
!***************************************************************************************************
Module xnet_b200f
  Use, Intrinsic :: iso_c_binding
  Implicit None

  ! struct xb_net_arrays (include/xnet_b200.h:52-68)
  Type, Bind(C) :: xb_net_arrays
    Integer(C_INT32_T) :: ny, nffn, nnnu
    Integer(C_INT32_T) :: nreac(4), nan(4)
    Type(C_PTR) :: rc(4), ni(4)
    Integer(C_INT32_T) :: ni_ld(4)
    Type(C_PTR) :: iwk(4), irev(4)
    Type(C_PTR) :: la, le
    Type(C_PTR) :: mu(4), a(4)
    Type(C_PTR) :: aa, zz, mex, be
    Real(C_DOUBLE) :: mex_n, mex_p
    Type(C_PTR) :: g, t9i, ffnsum, ffnenu
    Type(C_PTR) :: ffn_ec = C_NULL_PTR, ffn_beta = C_NULL_PTR, ffn_qval = C_NULL_PTR, has_logft = C_NULL_PTR
    ! reaction pairs of Module xnet_match (mflx = 0: no flux output); dcflx may stay NULL
    Integer(C_INT32_T) :: mflx = 0
    Type(C_PTR) :: ifl(4) = C_NULL_PTR
    Type(C_PTR) :: nflx = C_NULL_PTR, iwflx = C_NULL_PTR, qflx = C_NULL_PTR, dcflx = C_NULL_PTR
  End Type xb_net_arrays

  ! struct xb_ctl (include/xnet_b200.h:71-77)
  Type, Bind(C) :: xb_ctl_t
    Integer(C_INT32_T) :: iweak0, iscrn, iheat, isolv, kstmx, kitmx, ijac, iconvc
    Real(C_DOUBLE) :: changemx, changemxt, yacc, tolm, tolc, tolt9, ymin, tdel_maxmult
    Integer(C_INT32_T) :: arith, solver
  End Type xb_ctl_t

  Interface
    Integer(C_INT) Function xb_net_from_arrays(arr,net) Bind(C,name="xb_net_from_arrays")
      Import ; Type(C_PTR), Value :: arr ; Type(C_PTR) :: net
    End Function
    Integer(C_INT) Function xb_net_load_helm(net,path) Bind(C,name="xb_net_load_helm")
      Import ; Type(C_PTR), Value :: net ; Character(kind=C_CHAR), Intent(in) :: path(*)
    End Function
    Subroutine xb_net_free(net) Bind(C,name="xb_net_free")
      Import ; Type(C_PTR), Value :: net
    End Subroutine
    Integer(C_INT) Function xb_ctx_create(net,ctl,nzevolve,device,ctx) Bind(C,name="xb_ctx_create")
      Import ; Type(C_PTR), Value :: net, ctl ; Integer(C_INT32_T), Value :: nzevolve, device ; Type(C_PTR) :: ctx
    End Function
    Subroutine xb_ctx_destroy(ctx) Bind(C,name="xb_ctx_destroy")
      Import ; Type(C_PTR), Value :: ctx
    End Subroutine
    Integer(C_INT) Function xb_cross_sect(ctx,zlo,zhi,mask,yt,t9t,rhot,xext,aext,zext,yet) Bind(C,name="xb_cross_sect")
      Import ; Type(C_PTR), Value :: ctx, mask, yt, t9t, rhot, xext, aext, zext, yet ; Integer(C_INT32_T), Value :: zlo, zhi
    End Function
    Integer(C_INT) Function xb_yderiv(ctx,zlo,zhi,mask,yt,ydot) Bind(C,name="xb_yderiv")
      Import ; Type(C_PTR), Value :: ctx, mask, yt, ydot ; Integer(C_INT32_T), Value :: zlo, zhi
    End Function
    Integer(C_INT) Function xb_jacobian_build(ctx,zlo,zhi,mask,diag,mult) Bind(C,name="xb_jacobian_build")
      Import ; Type(C_PTR), Value :: ctx, mask, diag, mult ; Integer(C_INT32_T), Value :: zlo, zhi
    End Function
    Integer(C_INT) Function xb_jacobian_scale(ctx,zlo,zhi,mask,diag,mult) Bind(C,name="xb_jacobian_scale")
      Import ; Type(C_PTR), Value :: ctx, mask, diag, mult ; Integer(C_INT32_T), Value :: zlo, zhi
    End Function
    Integer(C_INT) Function xb_get_csect(ctx,zlo,zhi,csect) Bind(C,name="xb_get_csect")
      Import ; Type(C_PTR), Value :: ctx, csect ; Integer(C_INT32_T), Value :: zlo, zhi
    End Function
    Integer(C_INT) Function xb_get_jac(ctx,zlo,zhi,jac) Bind(C,name="xb_get_jac")
      Import ; Type(C_PTR), Value :: ctx, jac ; Integer(C_INT32_T), Value :: zlo, zhi
    End Function
    Integer(C_INT) Function xb_jacobian_decomp(ctx,zlo,zhi,mask,info) Bind(C,name="xb_jacobian_decomp")
      Import ; Type(C_PTR), Value :: ctx, mask, info ; Integer(C_INT32_T), Value :: zlo, zhi
    End Function
    Integer(C_INT) Function xb_jacobian_bksub(ctx,zlo,zhi,mask,yrhs,t9rhs,dy,dt9) Bind(C,name="xb_jacobian_bksub")
      Import ; Type(C_PTR), Value :: ctx, mask, yrhs, t9rhs, dy, dt9 ; Integer(C_INT32_T), Value :: zlo, zhi
    End Function
    Integer(C_INT) Function xb_jacobian_solve(ctx,zlo,zhi,mask,yrhs,t9rhs,dy,dt9,info) Bind(C,name="xb_jacobian_solve")
      Import ; Type(C_PTR), Value :: ctx, mask, yrhs, t9rhs, dy, dt9, info ; Integer(C_INT32_T), Value :: zlo, zhi
    End Function
    ! self-heating state mirrors (dcsect*, cv/etae/detaedt9, t9dot)
    Integer(C_INT) Function xb_get_dcsect(ctx,zlo,zhi,dcsect) Bind(C,name="xb_get_dcsect")
      Import ; Type(C_PTR), Value :: ctx, dcsect ; Integer(C_INT32_T), Value :: zlo, zhi
    End Function
    Integer(C_INT) Function xb_get_eos(ctx,zlo,zhi,cv,etae,detaedt9,t9dot) Bind(C,name="xb_get_eos")
      Import ; Type(C_PTR), Value :: ctx, cv, etae, detaedt9, t9dot ; Integer(C_INT32_T), Value :: zlo, zhi
    End Function
    ! flux (xnet_flux.F90:67) from the device-resident rates
    Integer(C_INT) Function xb_flux(ctx,zlo,zhi,mask,yt,flx) Bind(C,name="xb_flux")
      Import ; Type(C_PTR), Value :: ctx, mask, yt, flx ; Integer(C_INT32_T), Value :: zlo, zhi
    End Function
    ! step_be (xnet_integrate_be.F90:233): the whole Newton loop in one launch
    Integer(C_INT) Function xb_step_be(ctx,zlo,zhi,kstep,inr,y,yt,tdel,t9,t9t,rhot,nh,ydot,t9dot,kcount) &
        & Bind(C,name="xb_step_be")
      Import ; Type(C_PTR), Value :: ctx, inr, y, yt, tdel, t9, t9t, rhot, nh, ydot, t9dot, kcount
      Integer(C_INT32_T), Value :: zlo, zhi, kstep
    End Function
    Integer(C_INT) Function xb_evolve(ctx,zones,out) Bind(C,name="xb_evolve")
      Import ; Type(C_PTR), Value :: ctx, zones, out
    End Function
  End Interface
End Module xnet_b200f
