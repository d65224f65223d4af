MODULE parafem_gpu
!------------------------------------------------------------------------------
! ISO_C_BINDING interfaces to libparafem_b200.so (include/parafem_b200.h).
! This is the module a ParaFEM maintainer drops next to the driver, exactly as
! xx3 does for its CUDA helpers (src/programs/dev/xx3/xx3.f90:56-148).  It is
! source only in this repository: the build image has no Fortran compiler.
!
! Link line (cf. src/programs/dev/xx3/makefile):
!   $(FC) p121_gpu.o parafem_gpu.o -L$(PARAFEM)/lib -lParaFEM_mpi \
!         -L$(PF_B200)/parafem_b200 -lparafem_b200 -o p121_gpu
!
! Conventions: every function returns 0 on success, >0 on error (the driver
! prints pf_last_error and STOPs, as xx3.f90:423-430); sizes are 64-bit
! (c_int64_t) because xx3's 32-bit byte counts overflow beyond 596 523 hex20
! elements; arrays are passed as-is (column-major, 1-based contents).
!------------------------------------------------------------------------------
  USE, INTRINSIC :: iso_c_binding
  IMPLICIT NONE

  INTERFACE

    INTEGER(c_int) FUNCTION pf_nccl_unique_id(id128) BIND(C,name='pf_nccl_unique_id')
      IMPORT; INTEGER(c_int8_t) :: id128(128)
    END FUNCTION

    ! rank = numpe-1 ; device = local GPU index ; id128 broadcast from rank 0 by MPI_BCAST
    INTEGER(c_int) FUNCTION pf_init(rank,nranks,device,id128,h) BIND(C,name='pf_init')
      IMPORT; INTEGER(c_int),VALUE :: rank,nranks,device
      INTEGER(c_int8_t) :: id128(128); TYPE(c_ptr) :: h
    END FUNCTION

    INTEGER(c_int) FUNCTION pf_finalize(h) BIND(C,name='pf_finalize')
      IMPORT; TYPE(c_ptr),VALUE :: h
    END FUNCTION

    INTEGER(c_int) FUNCTION pf_last_error(h,buf,len) BIND(C,name='pf_last_error')
      IMPORT; TYPE(c_ptr),VALUE :: h; CHARACTER(kind=c_char) :: buf(*); INTEGER(c_int),VALUE :: len
    END FUNCTION

    ! after make_ggl (p121.f90:49), before DEALLOCATE(g_g_pp) (p121.f90:86)
    INTEGER(c_int) FUNCTION pf_setup_mesh(h,nod,nodof,nip,nels_pp,g_coord_pp,g_g_pp,           &
                                          neq,ieq_start,neq_pp) BIND(C,name='pf_setup_mesh')
      IMPORT; TYPE(c_ptr),VALUE :: h; INTEGER(c_int),VALUE :: nod,nodof,nip
      INTEGER(c_int64_t),VALUE :: nels_pp,neq,ieq_start,neq_pp
      REAL(c_double) :: g_coord_pp(*); INTEGER(c_int) :: g_g_pp(*)
    END FUNCTION

    ! elements_1 / gauss_pts_1 of p121.f90:54-64 on the device
    INTEGER(c_int) FUNCTION pf_form_km_elastic(h,e,v) BIND(C,name='pf_form_km_elastic')
      IMPORT; TYPE(c_ptr),VALUE :: h; REAL(c_double),VALUE :: e,v
    END FUNCTION

    ! per-element materials (xx2.f90:169-193): prop(2,np_types) = (e,v), etype_pp(nels_pp)
    INTEGER(c_int) FUNCTION pf_form_km_elastic_mat(h,np_types,prop,etype_pp) BIND(C,name='pf_form_km_elastic_mat')
      IMPORT; TYPE(c_ptr),VALUE :: h; INTEGER(c_int),VALUE :: np_types
      REAL(c_double) :: prop(2,*); INTEGER(c_int) :: etype_pp(*)
    END FUNCTION

    ! elements_1 of p123.f90:70-84 on the device
    INTEGER(c_int) FUNCTION pf_form_kc_laplace(h,kx,ky,kz) BIND(C,name='pf_form_kc_laplace')
      IMPORT; TYPE(c_ptr),VALUE :: h; REAL(c_double),VALUE :: kx,ky,kz
    END FUNCTION

    ! p124 (transient conduction, theta method): storka_pp/storkb_pp of p124.f90:81-95 on the device,
    ! then per time step one call replacing p124.f90:143-218
    INTEGER(c_int) FUNCTION pf_form_k_transient(h,kx,ky,kz,rho,cp,theta,dtim) BIND(C,name='pf_form_k_transient')
      IMPORT; TYPE(c_ptr),VALUE :: h; REAL(c_double),VALUE :: kx,ky,kz,rho,cp,theta,dtim
    END FUNCTION
    INTEGER(c_int) FUNCTION pf_transient_start(h,val0,val_f_pp) BIND(C,name='pf_transient_start')
      IMPORT; TYPE(c_ptr),VALUE :: h; REAL(c_double),VALUE :: val0; REAL(c_double) :: val_f_pp(*)
    END FUNCTION
    INTEGER(c_int) FUNCTION pf_transient_step(h,loads_pp,tol,limit,iters,converged,elapsed_ms) &
        BIND(C,name='pf_transient_step')
      IMPORT; TYPE(c_ptr),VALUE :: h; REAL(c_double) :: loads_pp(*); REAL(c_double),VALUE :: tol
      INTEGER(c_int),VALUE :: limit; INTEGER(c_int) :: iters,converged; REAL(c_double) :: elapsed_ms
    END FUNCTION

    ! p125 (explicit transient conduction): store_pm_pp / globma_pp of p125.f90:66-82, then nsteps passes of :94-99
    INTEGER(c_int) FUNCTION pf_form_k_explicit(h,kx,ky,kz,dtim) BIND(C,name='pf_form_k_explicit')
      IMPORT; TYPE(c_ptr),VALUE :: h; REAL(c_double),VALUE :: kx,ky,kz,dtim
    END FUNCTION
    INTEGER(c_int) FUNCTION pf_explicit_start(h,val0) BIND(C,name='pf_explicit_start')
      IMPORT; TYPE(c_ptr),VALUE :: h; REAL(c_double),VALUE :: val0
    END FUNCTION
    INTEGER(c_int) FUNCTION pf_explicit_steps(h,nsteps,elapsed_ms) BIND(C,name='pf_explicit_steps')
      IMPORT; TYPE(c_ptr),VALUE :: h; INTEGER(c_int),VALUE :: nsteps; REAL(c_double) :: elapsed_ms
    END FUNCTION

    ! alternative: upload a host storkm_pp (xx3.f90:440-452)
    INTEGER(c_int) FUNCTION pf_set_storkm(h,storkm_pp) BIND(C,name='pf_set_storkm')
      IMPORT; TYPE(c_ptr),VALUE :: h; REAL(c_double) :: storkm_pp(*)
    END FUNCTION

    ! optional variants, call before forming the matrices: matrix-free operator (1 rebuild, 2 stored
    ! geometric factors) and the packed-lower-triangle storkm layout (1; 0 = storkm_pp as in p121.f90:33)
    INTEGER(c_int) FUNCTION pf_set_matrix_free(h,mode) BIND(C,name='pf_set_matrix_free')
      IMPORT; TYPE(c_ptr),VALUE :: h; INTEGER(c_int),VALUE :: mode
    END FUNCTION
    INTEGER(c_int) FUNCTION pf_set_storkm_layout(h,layout) BIND(C,name='pf_set_storkm_layout')
      IMPORT; TYPE(c_ptr),VALUE :: h; INTEGER(c_int),VALUE :: layout
    END FUNCTION

    ! p121.f90:65-69,86 / p123.f90:86-92,120-125 ; no_f_pp = GLOBAL equation numbers
    INTEGER(c_int) FUNCTION pf_build_precon(h,nfixed_pp,no_f_pp,penalty) BIND(C,name='pf_build_precon')
      IMPORT; TYPE(c_ptr),VALUE :: h; INTEGER(c_int64_t),VALUE :: nfixed_pp
      INTEGER(c_int) :: no_f_pp(*); REAL(c_double),VALUE :: penalty
    END FUNCTION

    INTEGER(c_int) FUNCTION pf_get_diag_precon(h,diag_precon_pp) BIND(C,name='pf_get_diag_precon')
      IMPORT; TYPE(c_ptr),VALUE :: h; REAL(c_double) :: diag_precon_pp(*)
    END FUNCTION

    INTEGER(c_int) FUNCTION pf_get_store(h,store_pp) BIND(C,name='pf_get_store')
      IMPORT; TYPE(c_ptr),VALUE :: h; REAL(c_double) :: store_pp(*)
    END FUNCTION

    ! p121.f90:87-104 : d=M^-1 r, p=d, x=0, PCG loop with checon_par every iteration
    INTEGER(c_int) FUNCTION pf_pcg_solve(h,r_pp,tol,limit,xnew_pp,iters,converged)             &
                                         BIND(C,name='pf_pcg_solve')
      IMPORT; TYPE(c_ptr),VALUE :: h; REAL(c_double) :: r_pp(*),xnew_pp(*)
      REAL(c_double),VALUE :: tol; INTEGER(c_int),VALUE :: limit; INTEGER(c_int) :: iters,converged
    END FUNCTION

    ! fine-grained twins of gather / elements_3 / scatter / dot_product_p (tests)
    INTEGER(c_int) FUNCTION pf_gather(h,p_pp,pmul_pp) BIND(C,name='pf_gather')
      IMPORT; TYPE(c_ptr),VALUE :: h; REAL(c_double) :: p_pp(*),pmul_pp(*)
    END FUNCTION
    INTEGER(c_int) FUNCTION pf_matvec(h,pmul_pp,utemp_pp) BIND(C,name='pf_matvec')
      IMPORT; TYPE(c_ptr),VALUE :: h; REAL(c_double) :: pmul_pp(*),utemp_pp(*)
    END FUNCTION
    INTEGER(c_int) FUNCTION pf_scatter(h,utemp_pp,u_pp) BIND(C,name='pf_scatter')
      IMPORT; TYPE(c_ptr),VALUE :: h; REAL(c_double) :: utemp_pp(*),u_pp(*)
    END FUNCTION
    INTEGER(c_int) FUNCTION pf_dot(h,a_pp,b_pp,res) BIND(C,name='pf_dot')
      IMPORT; TYPE(c_ptr),VALUE :: h; REAL(c_double) :: a_pp(*),b_pp(*),res
    END FUNCTION

    ! p121.f90:113-123 ; iel is 0-based local
    INTEGER(c_int) FUNCTION pf_centroid_stress(h,iel,e,v,sigma) BIND(C,name='pf_centroid_stress')
      IMPORT; TYPE(c_ptr),VALUE :: h; INTEGER(c_int64_t),VALUE :: iel
      REAL(c_double),VALUE :: e,v; REAL(c_double) :: sigma(6)
    END FUNCTION

    ! the same at a local point (xi,eta,zeta) of the element (the 2013 build behind p121/book/p121.res
    ! printed the stress at the last Gauss point)
    INTEGER(c_int) FUNCTION pf_point_stress(h,iel,xi,eta,zeta,e,v,sigma) BIND(C,name='pf_point_stress')
      IMPORT; TYPE(c_ptr),VALUE :: h; INTEGER(c_int64_t),VALUE :: iel
      REAL(c_double),VALUE :: xi,eta,zeta,e,v; REAL(c_double) :: sigma(6)
    END FUNCTION

    ! PCG_KM (maths.f90:1152-1323): one km(ntot,ntot) for every element, the caller's inverted preconditioner
    INTEGER(c_int) FUNCTION pf_pcg_km(h,km,diag_precon_pp,r_pp,tol,limit,xnew_pp,iters,converged)  &
                                      BIND(C,name='pf_pcg_km')
      IMPORT; TYPE(c_ptr),VALUE :: h; REAL(c_double) :: km(*),diag_precon_pp(*),r_pp(*),xnew_pp(*)
      REAL(c_double),VALUE :: tol; INTEGER(c_int),VALUE :: limit; INTEGER(c_int) :: iters,converged
    END FUNCTION

    ! p122.f90:88-93 (after pf_form_km_elastic + pf_build_precon), :115-205 (one load increment), :206-214
    INTEGER(c_int) FUNCTION pf_plastic_begin(h,phi,c,psi,e,v,dt) BIND(C,name='pf_plastic_begin')
      IMPORT; TYPE(c_ptr),VALUE :: h; REAL(c_double),VALUE :: phi,c,psi,e,v; REAL(c_double) :: dt
    END FUNCTION
    INTEGER(c_int) FUNCTION pf_plastic_increment(h,qinc,ld0_pp,valf_pp,plasits,plastol,cjits,cjtol,  &
                                                 plasiters,cjtot,elapsed_ms) BIND(C,name='pf_plastic_increment')
      IMPORT; TYPE(c_ptr),VALUE :: h; REAL(c_double),VALUE :: qinc,plastol,cjtol
      TYPE(c_ptr),VALUE :: ld0_pp,valf_pp        ! C_LOC(ld0_pp) / C_LOC(valf) or C_NULL_PTR
      INTEGER(c_int),VALUE :: plasits,cjits; INTEGER(c_int) :: plasiters,cjtot; REAL(c_double) :: elapsed_ms
    END FUNCTION
    INTEGER(c_int) FUNCTION pf_plastic_get(h,totd_pp,iel,ig,tensor6) BIND(C,name='pf_plastic_get')
      IMPORT; TYPE(c_ptr),VALUE :: h,totd_pp,tensor6; INTEGER(c_int64_t),VALUE :: iel; INTEGER(c_int),VALUE :: ig
    END FUNCTION

    ! p1210 (explicit elasto-plastic von Mises dynamics, lumped mass): begin after pf_setup_mesh, npri steps per call
    INTEGER(c_int) FUNCTION pf_vm_explicit_begin(h,e,v,sbary,rho,dtim,pload,fext_pp) BIND(C,name='pf_vm_explicit_begin')
      IMPORT; TYPE(c_ptr),VALUE :: h,fext_pp; REAL(c_double),VALUE :: e,v,sbary,rho,dtim,pload
    END FUNCTION
    INTEGER(c_int) FUNCTION pf_vm_explicit_set_form(h,form) BIND(C,name='pf_vm_explicit_set_form')
      IMPORT; TYPE(c_ptr),VALUE :: h; INTEGER(c_int),VALUE :: form    ! 0: elements_2 as written, 1: tensor-core operator form
    END FUNCTION
    INTEGER(c_int) FUNCTION pf_vm_explicit_steps(h,nsteps,elapsed_ms) BIND(C,name='pf_vm_explicit_steps')
      IMPORT; TYPE(c_ptr),VALUE :: h; INTEGER(c_int),VALUE :: nsteps; REAL(c_double) :: elapsed_ms
    END FUNCTION
    INTEGER(c_int) FUNCTION pf_vm_explicit_get(h,x1_pp,d1x1_pp,d2x1_pp,mm_pp) BIND(C,name='pf_vm_explicit_get')
      IMPORT; TYPE(c_ptr),VALUE :: h,x1_pp,d1x1_pp,d2x1_pp,mm_pp       ! C_LOC(...) or C_NULL_PTR
    END FUNCTION

    ! 0: one rank, 1: NCCL send/recv, 2: peer memory over NVLink
    INTEGER(c_int) FUNCTION pf_halo_transport(h) BIND(C,name='pf_halo_transport')
      IMPORT; TYPE(c_ptr),VALUE :: h
    END FUNCTION

  END INTERFACE

CONTAINS

  SUBROUTINE pf_check(status,h,what)
    ! xx3.f90:423-430 convention: print and stop on status > 0
    INTEGER(c_int),INTENT(IN) :: status; TYPE(c_ptr),INTENT(IN) :: h
    CHARACTER(LEN=*),INTENT(IN) :: what
    CHARACTER(kind=c_char) :: buf(512); INTEGER(c_int) :: n,i
    IF (status > 0) THEN
      n = pf_last_error(h,buf,512_c_int)
      PRINT *, "parafem_gpu: ", what, " failed, status ", status
      PRINT *, (buf(i),i=1,MIN(n,511))
      STOP
    END IF
  END SUBROUTINE pf_check

END MODULE parafem_gpu
