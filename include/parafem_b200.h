/*
 * parafem_b200.h -- C-ABI of libparafem_b200.so
 *
 * B200-native (sm_100a) replacement for the EBE-PCG hot path of ParaFEM programs
 * p121 (3-D elasticity, 20-/8-node hexahedra) and p123 (steady heat conduction,
 * 8-node hexahedra), and of the drivers that reuse the same gather / mat-vec /
 * scatter kernels: p124 (implicit transient conduction), p125 (explicit transient
 * conduction), xx2 (per-element materials); 4-node tetrahedra are served too.
 * It is, in the reference's own terms, the missing "modules/gpu" platform library
 * (parafem/src/modules/readme.txt:9-40).  The six symbols the reference's existing
 * GPU driver xx3 binds are declared in parafem_xx3_compat.h.
 *
 * Conventions (same as the reference's existing CUDA boundary,
 * parafem/src/programs/dev/xx3/cuda_helpers.cu:183-368 and the
 * `interface ... bind(C)` block at xx3.f90:56-148):
 *   - plain C symbols, callable from Fortran through ISO_C_BINDING
 *     (see fortran/parafem_gpu.f90) -- no torch / C++ types in any signature;
 *   - every function returns 0 on success and >0 on error; it never calls
 *     exit().  pf_last_error() returns the message (xx3 prints it with printf);
 *   - all host arrays are the caller's, in Fortran (column-major) layout with
 *     1-based CONTENTS (node numbers, equation numbers; 0 = restrained);
 *     the library never keeps a host pointer after a call returns;
 *   - reals are REAL(iwp) = double (precision.f90:19); node and equation numbers
 *     are default INTEGER = int32 as in the reference; SIZES are int64 because
 *     xx3's int*int byte counts overflow past 596 523 hex20 elements
 *     (cuda_helpers.cu:205,242,264);
 *   - one MPI rank <-> one process <-> one GPU; rank = numpe-1.
 *
 * Section A is the device API (needs a B200).  Section B are host-side helpers
 * that restate the ParaFEM library routines the Fortran driver would call
 * between read_p121 and make_ggl; they exist because this image has no Fortran
 * compiler and the host driver (parafem_b200/csrc/p121_b200.cpp and the Python
 * mirror parafem_b200/driver.py) has to be written in C++/Python.  They run on
 * the CPU and do not need a GPU.
 */
#ifndef PARAFEM_B200_H
#define PARAFEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pf_ctx *pf_handle;

/* ===================================================================== */
/* A. Device API                                                          */
/* ===================================================================== */

/* --- process / device / communicator ---------------------------------
 * Replaces set_gpu (xx3/cuda_helpers.cu:183-192) and find_pe_procs' device
 * side (mp_interface.f90:42-91).  For nranks > 1 rank 0 calls
 * pf_nccl_unique_id(), the driver broadcasts the 128 bytes (MPI_BCAST in
 * Fortran, torch.distributed in bench.py) and every rank passes them to
 * pf_init.  id128 may be NULL when nranks == 1.                          */
int pf_nccl_unique_id(void *id128);
int pf_init(int rank, int nranks, int device, const void *id128, pf_handle *h);
int pf_finalize(pf_handle h);
int pf_last_error(pf_handle h, char *buf, int len);  /* h may be NULL: last global error */
int pf_version(void);

/* --- mesh setup --------------------------------------------------------
 * Called once between make_ggl (p121.f90:49) and DEALLOCATE(g_g_pp)
 * (p121.f90:86).  Replaces allocate_memory_on_gpu + copy_*_data_to_gpu
 * (xx3.f90:409-463).  The gather/scatter tables (ggl_pp, toget/toput of
 * gather_scatter.f90:50-70, PRIVATE there) are rebuilt inside the library
 * from g_g_pp and the closed-form owner map of calc_neq_pp
 * (gather_scatter.f90:319-339), so the Fortran module needs no accessor.
 *   nod          nodes per element: 8 or 20 (hexahedra) or 4 (tetrahedra; shape_der nod = 4,
 *                new_library.f90:757-767, with sample('tetrahedron') nip = 1, :1329-1341);
 *                nodof dof per node (3 or 1)
 *   nip          integrating points (1 or 8; 1 for tetrahedra)
 *   nels_pp      elements of this rank
 *   g_coord_pp   (nod, ndim=3, nels_pp) coordinates        [p121.f90:33]
 *   g_g_pp       (ntot, nels_pp) global equation numbers, 0 = restrained
 *   neq          global equation count; ieq_start 1-based; neq_pp owned   */
int pf_setup_mesh(pf_handle h, int nod, int nodof, int nip, int64_t nels_pp,
                  const double *g_coord_pp, const int32_t *g_g_pp,
                  int64_t neq, int64_t ieq_start, int64_t neq_pp);

/* --- element matrices --------------------------------------------------
 * pf_form_km_elastic: elements_1/gauss_pts_1 of p121.f90:54-64 (deemat,
 * sample, shape_der, invert, beemat, BtDB*det*w) formed ON the device into
 * storkm_pp(ntot,ntot,nels_pp), which never crosses PCIe.
 * pf_form_kc_laplace: p123.f90:70-84 (kcx,kcy,kcz -> storkc_pp).
 * pf_set_storkm uploads a host storkm_pp instead (xx3's
 * copy_3d_data_to_gpu, xx3.f90:440-452); pf_get_storkm reads n elements
 * starting at 0-based local element iel0 back (parity tests).
 * pf_set_matrix_free(1) (call before pf_form_km_elastic): BASELINE config E.
 * storkm is never stored; every iteration recomputes the element operator
 * from g_coord_pp as sum_gp B^T (D (B p)) det w (p121 elements, nip = 8).
 * Only the diagonal of km is formed once, for the preconditioner.
 * pf_set_matrix_free(2): as 1, but the inverse Jacobian and det*w of every
 * Gauss point (10 doubles, 640 B per element) are stored at setup and read
 * back instead of being rebuilt -- same bits as mode 1, no FP64 divisions
 * in the loop ("partial assembly").
 * Both modes run on the FP64 tensor cores (k_apply_mf4: the two node sums as
 * mma.sync.m8n8k4.f64 with der as constant register fragments; a DMMA is a
 * k-ascending fma chain bit for bit, which is how the oracle states it).
 * pf_set_storkm_layout(h, 1) (call before forming / uploading the matrices):
 * keep only the lower triangle of every element matrix, packed by columns
 * (ntot(ntot+1)/2 doubles: 14 640 B instead of 28 800 B per 20-node brick),
 * which halves the storkm stream of every iteration.  The product keeps the
 * reference's summation order with K(i,j) = L(max(i,j),min(i,j)); it is
 * bit-identical to MATMUL on the symmetrised matrix, and differs from the
 * reference's unsymmetrised storkm_pp only where BtDB rounding breaks the
 * symmetry (relative 1e-16).  pf_get_storkm returns the symmetrised matrices.
 * Layout 0 (default) is the reference's storkm_pp, bit for bit.             */
int pf_set_storkm_layout(pf_handle h, int layout);
int pf_form_km_elastic(pf_handle h, double e, double v);
int pf_form_kc_laplace(pf_handle h, double kx, double ky, double kz);
/* Per-element materials (xx2 / rfemsolve): elements_3 of
 * programs/dev/xx2/xx2.f90:169-193, e = prop(1,etype_pp(iel)), v = prop(2,etype_pp(iel)).
 * prop(2,np_types) column-major, etype_pp(nels_pp) 1-based material numbers (the last
 * column of the .d element lines, read_elements input.f90:1434-1583).               */
int pf_form_km_elastic_mat(pf_handle h, int np_types, const double *prop, const int32_t *etype_pp);
int pf_set_storkm(pf_handle h, const double *storkm_pp);
int pf_get_storkm(pf_handle h, int64_t iel0, int64_t n, double *out);
int pf_set_matrix_free(pf_handle h, int on);

/* --- diagonal preconditioner -------------------------------------------
 * p121.f90:65-69,86 / p123.f90:86-92,120-125.  no_f_pp are the GLOBAL
 * 1-based equation numbers of this rank's fixed freedoms (p123's no_f_pp),
 * nfixed_pp may be 0.  After the call the device holds 1/diag; store_pp
 * (p123.f90:123) can be fetched with pf_get_store.  pf_get_diag_precon
 * returns the INVERTED diagonal (diag_precon_pp after p121.f90:86).       */
int pf_build_precon(pf_handle h, int64_t nfixed_pp, const int32_t *no_f_pp,
                    double penalty);
int pf_get_diag_precon(pf_handle h, double *diag_precon_pp);
int pf_get_store(pf_handle h, double *store_pp);

/* --- the solve -----------------------------------------------------------
 * pf_pcg_solve = p121.f90:87-104 / p123.f90:132-151: d=M^-1 r, p=d, x=0,
 * then the PCG loop with checon_par's stopping rule evaluated every
 * iteration.  r_pp in (host, neq_pp), xnew_pp out (host, neq_pp).
 * The three-call form keeps the vectors resident (bench `value`):
 * pf_pcg_load_rhs (H2D) / pf_pcg_run (device only; elapsed_ms from CUDA
 * events on the solver stream) / pf_pcg_get_x (D2H).                      */
int pf_pcg_solve(pf_handle h, const double *r_pp, double tol, int limit,
                 double *xnew_pp, int *iters, int *converged);
int pf_pcg_load_rhs(pf_handle h, const double *r_pp);
int pf_pcg_run(pf_handle h, double tol, int limit, int *iters, int *converged,
               double *elapsed_ms);
int pf_pcg_get_x(pf_handle h, double *xnew_pp);
/* checon_par ratio max|xnew-x|/max|xnew| of every iteration of the last run */
/* PCG_KM (maths.f90:1152-1323), SURVEY 8a row a14: the solver for meshes of congruent elements -- one
 * km(ntot,ntot) for every element (utemp_pp = MATMUL(km,pmul_pp)), the inverted diagonal preconditioner
 * diag_precon_pp(neq_pp) supplied by the caller as in the Fortran argument list; r_pp in, xnew_pp out.  The element
 * matrix lives in registers on the device: no storkm stream at all.                                          */
int pf_pcg_km(pf_handle h, const double *km, const double *diag_precon_pp, const double *r_pp, double tol, int limit,
              double *xnew_pp, int *iters, int *converged);
int pf_get_ratio_history(pf_handle h, double *out, int maxn, int *n);

/* --- transient conduction: program p124 ---------------
 * Same three kernels as p123, two element-matrix sets and one PCG solve per
 * time step (implicit theta method), everything resident on the device.
 * pf_form_k_transient: elements_3/gauss_pts of p124.f90:81-95 for 8-node bricks,
 *   kc += MATMUL(MATMUL(TRANSPOSE(deriv),kay),deriv)*det*w, pm += fun fun^T det*w*rho*cp,
 *   storka_pp = pm + kc*theta*dtim (the PCG matrix: pf_get_storkm, pf_build_precon,
 *   pf_apply see it), storkb_pp = pm - kc*(1-theta)*dtim (pf_get_storkb).
 *   One material (p12meshgen forces np_types = 1, p12meshgen.f90:837).
 * pf_build_precon as for p123 (no_f_pp = fixed freedoms, penalty 1e20; p124.f90:97-124).
 * pf_transient_start: x_pp = val0, x_pp(l) = val_f(k) on the fixed freedoms
 *   (p124.f90:168-173); val_f_pp (nfixed_pp, order of no_f_pp) may be NULL when none.
 * pf_transient_step: one pass of the timesteps loop (p124.f90:139-218): right-hand side
 *   loads_pp + storkb*xnew (storka*x on the first step after pf_transient_start), fixed-freedom
 *   rows as the reference writes them, then PCG from x = 0.  loads_pp (host, neq_pp) =
 *   val*dtim at the loaded freedoms, NULL = none.  xnew_pp stays on the device
 *   (pf_pcg_get_x reads it); elapsed_ms = the whole step on the solver stream.       */
int pf_form_k_transient(pf_handle h, double kx, double ky, double kz, double rho, double cp,
                        double theta, double dtim);
int pf_get_storkb(pf_handle h, int64_t iel0, int64_t n, double *out);
int pf_transient_start(pf_handle h, double val0, const double *val_f_pp);
int pf_transient_step(pf_handle h, const double *loads_pp, double tol, int limit, int *iters,
                      int *converged, double *elapsed_ms);

/* --- explicit transient conduction: program p125 ------
 * The gather / mat-vec / scatter kernels without a solver: forward Euler with a lumped mass.
 * pf_form_k_explicit: elements_1 of p125.f90:66-82 for 8-node bricks: store_pm_pp = diag(mass) - kc*dtim
 *   with mass(i) = SUM(pm(i,:)) (pf_get_storkm returns it) and globma_pp = 1/scatter(mass)
 *   (pf_get_diag_precon returns it).
 * pf_explicit_start: loads_pp = val0 (p125.f90:83).
 * pf_explicit_steps: nsteps passes of p125.f90:94-99, loads_pp = scatter(store_pm*gather(loads_pp))*globma_pp,
 *   resident on the device; pf_pcg_get_x reads the field; elapsed_ms on the solver stream.              */
int pf_form_k_explicit(pf_handle h, double kx, double ky, double kz, double dtim);
int pf_explicit_start(pf_handle h, double val0);
int pf_explicit_steps(pf_handle h, int nsteps, double *elapsed_ms);

/* p1210 (programs/5th_ed/p1210/p1210.f90: forced vibration of an elastic-plastic von Mises solid, lumped mass,
 * explicit integration).  No element matrices, no PCG: after pf_setup_mesh (20-node bricks, nip 8)
 *   pf_vm_explicit_begin : element tables for (e, v), the diagonal mass matrix mm_pp (p1210.f90:93-104, scattered with
 *                          the reverse halo exchange), fext_pp(neq_pp) from load() (:106-111), zero state (:112);
 *   pf_vm_explicit_steps : nsteps passes of time_steps (:114-150) on the device: x1 += dtim*d1x1 + dtim**2/2*d2x1,
 *                          gather, the Gauss-point stress update (invar / vmpl, tensor_pp / etensor_pp resident),
 *                          scatter of -bload, bdylds = (bdylds + fext*pload)/mm, velocity and acceleration updates;
 *   pf_vm_explicit_get   : x1_pp, d1x1_pp, d2x1_pp, mm_pp (any may be NULL).                                      */
int pf_vm_explicit_begin(pf_handle h, double e, double v, double sbary, double rho, double dtim, double pload,
                         const double *fext_pp);
/* form 0 (default): the Gauss-point loop as p1210.f90:120-147 writes it, bit for bit (k_p1210_elements); form 1: the
 * same update in operator form on the FP64 tensor cores (the matrix-free kernel's pipeline with the elasto-plastic
 * point update in the middle) -- another rounding (1e-15), ~20x faster, pinned to the same golden fields.        */
int pf_vm_explicit_set_form(pf_handle h, int form);
int pf_vm_explicit_steps(pf_handle h, int nsteps, double *elapsed_ms);
int pf_vm_explicit_get(pf_handle h, double *x1_pp, double *d1x1_pp, double *d2x1_pp, double *mm_pp);
/* --- p129: forced vibration, implicit theta method, consistent mass ----------
 * programs/5th_ed/p129/p129.f90.  The driver keeps its input section, the harmonic load factor and its output.
 *   pf_form_dynamic   elements_2 (p129.f90:83-98): store_km_pp and the consistent store_mm_pp (ecmat, shape_fun; 20-node
 *                     bricks with nip = 8 or the 27-point rule, 8-node bricks with nip = 8), combined once into the three
 *                     matrix sets the time loop uses: store_mm*c3 + store_km*c4 (PCG, :128), store_km*c2 + store_mm*c3
 *                     (:114) and store_mm/theta (:120), c1..c4 as :80-82.  Then pf_build_precon (:99-105).
 *   pf_dynamic_start  fext_pp (load, :106-111); x0 = d1x0 = d2x0 = 0 (:112).
 *   pf_dynamic_step   one pass of timesteps (:113-149): both right-hand-side products, loads = u + vu + fext*load_factor
 *                     with load_factor = theta*dtim*cos(omega*t) + c1*cos(omega*(t-dtim)) from the caller, the PCG solve
 *                     from x = 0, and the displacement / velocity / acceleration update.
 *   pf_dynamic_get    x1_pp, d1x1_pp, d2x1_pp of the last step (any pointer may be NULL).                         */
int pf_form_dynamic(pf_handle h, double e, double v, double rho, double alpha1, double beta1, double theta, double dtim);
int pf_dynamic_start(pf_handle h, const double *fext_pp);
int pf_dynamic_step(pf_handle h, double load_factor, double tol, int limit, int *iters, int *converged, double *elapsed_ms);
int pf_dynamic_get(pf_handle h, double *x_pp, double *d1x_pp, double *d2x_pp);

/* --- p122: 3-D elasto-plasticity ------------------------
 * programs/5th_ed/p122/p122.f90: Mohr-Coulomb solid, viscoplastic strain method.  The driver keeps its input
 * section and output; the device holds storkm_pp, evpt_pp / tensor_pp (nst,nip,nels_pp) and every vector of the
 * load-increment loop.
 *   pf_plastic_begin      p122.f90:88-93 after pf_form_km_elastic(e,v) + pf_build_precon(fixed freedoms, 1e20):
 *                         tensor_pp = totd_pp = oldis_pp = x_pp = 0, dt (returned).
 *   pf_plastic_increment  one pass of load_increments, p122.f90:115-205: plastic iterations until checon_par on the
 *                         displacement increment (plastol) or plasits; each one builds loads_pp (ld0_pp*qinc and /
 *                         or store_pp*valf*qinc on the fixed freedoms of this rank, + bdylds_pp), r = loads - A*x,
 *                         a PCG solve restarted from the current x (cjits, cjtol; fixed rows u = p*store on the
 *                         first plastic iteration, 0 afterwards) and the Gauss-point update of elements_4 (invar,
 *                         mocouf, mocouq, formm -- new_library.f90:1813-1916, 2364-2490, 85-198) whose body
 *                         loads are scattered into bdylds_pp.  ld0_pp (neq_pp) may be NULL (no loaded nodes);
 *                         valf_pp holds the values of this rank's fixed freedoms in the order given to
 *                         pf_build_precon.
 *   pf_plastic_get        totd_pp and tensor_pp(:,ig,iel) for the log lines of p122.f90:206-214.            */
int pf_plastic_begin(pf_handle h, double phi, double c, double psi, double e, double v, double *dt);
int pf_plastic_increment(pf_handle h, double qinc, const double *ld0_pp, const double *valf_pp, int plasits,
                         double plastol, int cjits, double cjtol, int *plasiters, int *cjtot, double *elapsed_ms);
int pf_plastic_get(pf_handle h, double *totd_pp, int64_t iel, int ig, double *tensor6);

/* --- fine-grained entry points (kernel-level parity tests) --------------
 * Same argument meaning as the reference routines they replace:
 *   pf_gather  = gather(p_pp,pmul_pp)            gather_scatter.f90:547-688
 *   pf_matvec  = elements_3 loop                 p121.f90:93-97
 *   pf_scatter = scatter(u_pp,utemp_pp), u_pp zeroed first  :694-850
 *   pf_apply   = the three fused as the solver runs them (u = A p)
 *   pf_dot     = dot_product_p                   maths.f90:168-216
 *   pf_norm    = norm_p                          maths.f90:222-265
 *   pf_sum     = sum_p                           maths.f90:271-315
 * All pointers are host arrays.                                           */
int pf_gather(pf_handle h, const double *p_pp, double *pmul_pp);
int pf_matvec(pf_handle h, const double *pmul_pp, double *utemp_pp);
int pf_scatter(pf_handle h, const double *utemp_pp, double *u_pp);
int pf_apply(pf_handle h, const double *p_pp, double *u_pp);
int pf_dot(pf_handle h, const double *a_pp, const double *b_pp, double *result);
int pf_norm(pf_handle h, const double *a_pp, double *result);
int pf_sum(pf_handle h, const double *a_pp, double *result);

/* --- post-solve ---------------------------------------
 * pf_centroid_stress: p121.f90:113-123 for local 0-based element iel
 * (gather of xnew into eld, one shape_der/beemat at the centroid,
 * sigma = dee*bee*eld).  Uses the x of the last solve.                    */
int pf_centroid_stress(pf_handle h, int64_t iel, double e, double v, double *sigma6);
/* The same at the local point (xi, eta, zeta) of the element.  The 2013 build of p121 that produced
 * examples/5th_ed/p121/book/p121.res printed the stress at the LAST Gauss point of the rule,
 * (-1/sqrt(3), -1/sqrt(3), -1/sqrt(3)), not at the centroid (p121.res:9-10). */
int pf_point_stress(pf_handle h, int64_t iel, double xi, double eta, double zeta, double e, double v, double *sigma6);

/* --- measurement ---------------------------------------------------------
 * pf_set_profile(1) brackets every launch of the named kernels with CUDA
 * events on the solver stream; pf_get_kernel_ms returns the sum and the
 * count since the last pf_pcg_run/pf_reset_profile.
 * which: 0 mat-vec (storkm stream), 1 scatter, 2 vector updates+reductions,
 *        3 halo exchange.                                                 */
int pf_set_profile(pf_handle h, int on);
int pf_reset_profile(pf_handle h);
int pf_get_kernel_ms(pf_handle h, int which, double *total_ms, int64_t *launches);
int64_t pf_kernel_launches(pf_handle h); /* all kernel launches since pf_init */
/* DFMA micro-benchmark on this device: the FP64 roofline denominator of the
 * matrix-free variant ("FP64 peak is not in MEASURED_PEAKS.json").  */
int pf_measure_fp64(pf_handle h, double *tflops);
/* The same for the FP64 tensor pipe (mma.sync.m8n8k4.f64, 512 flop per warp instruction): the roofline
 * denominator of the tensor-core matrix-free kernels (k_apply_mf4 / k_apply_mf3).              */
int pf_measure_fp64_tensor(pf_handle h, double *tflops);
/* The mat-vec kernel of the current problem, `reps` launches back to back between one pair of events
 * (ms per launch): the kernel time without per-launch event records, for launches of ~0.1 ms.   */
int pf_measure_matvec(pf_handle h, int reps, double *ms_per_launch);
/* Read-only HBM stream through the same bulk-copy ring as the mat-vec, no arithmetic (GB/s):
 * MEASURED_PEAKS.json's hbm_gbs is a copy (read + write); this is the read-stream ceiling.     */
int pf_measure_hbm_read(pf_handle h, double *gbs);
int pf_device_info(pf_handle h, int *sm_count, int64_t *free_bytes, int64_t *total_bytes);
/* How the halo exchange of this handle travels: 0 one rank (none), 1 NCCL send/recv + all-gather,
 * 2 peer memory over NVLink (CUDA IPC mappings; gather_scatter.f90:547-850 replaced either way). */
int pf_halo_transport(pf_handle h);
/* Device time (CUDA events on the solver stream) of the iteration loop of the last pf_pcg_run /
 * pf_pcg_solve: the window of timest(3) in p121.f90:89,107-108. */
int pf_get_last_solve_ms(pf_handle h, double *ms);

/* ===================================================================== */
/* B. Host helpers (CPU; restate ParaFEM library routines)                */
/* ===================================================================== */

/* calc_nels_pp partitioner 1 (gather_scatter.f90:217-238) and calc_neq_pp
 * (gather_scatter.f90:319-339).  numpe is 1-based; *_start 1-based.       */
void pf_calc_nels_pp(int64_t nels, int npes, int numpe, int64_t *nels_pp, int64_t *iel_start);
void pf_calc_neq_pp(int64_t neq, int npes, int numpe, int64_t *neq_pp, int64_t *ieq_start);
/* calc_nels_pp partitioner 2 = read_nels_pp (input.f90:3108-3196): <job>.psize holds
 * "npes n_1 ... n_npes" (elements pre-sorted by partition, e.g. by METIS).  Any contiguous
 * element ranges are accepted by pf_setup_mesh; the equation partition stays calc_neq_pp. */
int pf_read_psize(const char *job, int npes, int numpe, int64_t *nels_pp, int64_t *iel_start);

/* p12meshgen cubes (tools/preprocessing/p12meshgen/p12meshgen.f90:118-236,
 * 658-701; geometry.f90 geometry_20bxz :175-286, geometry_8bxz :70-169,
 * cube_bc20 :425-500, cube_bc8 :589-650, box_bc8 :652-696; loading.f90
 * load_p121 :386-546).
 * round_mode 0: full-precision values; 1: values as they survive the deck
 * text formats (coordinates E14.6, loads E16.8).                          */
int pf_p121_sizes(int nxe, int nye, int nze, int nod,
                  int64_t *nn, int64_t *nr, int64_t *loaded_nodes);
int pf_p123_sizes(int nxe, int nye, int nze, int64_t *nn, int64_t *nr, int64_t *nres);
/* elements iel_start .. iel_start+nels_pp-1 (1-based) in S&G node order:
 * g_num_pp(nod,nels_pp), g_coord_pp(nod,3,nels_pp)                         */
int pf_cube_elements(int nxe, int nze, int nod, double aa, double bb, double cc,
                     int64_t iel_start, int64_t nels_pp, int round_mode,
                     int32_t *g_num_pp, double *g_coord_pp);
/* rest(nr,nodof+1) column-major: kind 0 = cube_bc20/cube_bc8 (p121),
 * kind 1 = box_bc8 (p123)                                                 */
int pf_cube_rest(int kind, int nxe, int nye, int nze, int nod, int64_t nr, int32_t *rest);
/* load_p121 scaled as p12meshgen does; node(loaded), val(3,loaded)        */
int pf_p121_loads(int nxe, int nze, int nod, double aa, double bb, int round_mode,
                  int32_t *node, double *val);

/* rearrange + find_g3 (new_library.f90:3059-3112, 3130-3212) and
 * rearrange_2 + find_g4 (:3118-3124, 3249-3271) restated as "number the free
 * freedoms in ascending node order": nf(nodof,nn), 0 = restrained.         */
int pf_form_nf(int64_t nn, int nodof, int64_t nr, const int32_t *rest,
               int32_t *nf, int64_t *neq);
/* Node numbers come from deck files: every routine that indexes with one takes nn and returns status 5 for a
 * number outside 1..nn instead of reading out of bounds (the readers return 8 for the connectivity, pf_form_nf 2
 * for the restraint list, pf_read_dat 4 for sizes no deck can have).                                      */
int pf_find_g(int nod, int nodof, int64_t nels_pp, int64_t nn, const int32_t *g_num_pp,
              const int32_t *nf, int32_t *g_g_pp);
/* load + scatter_noadd (loading.f90:36-142): r_pp(neq_pp) from nodal loads */
int pf_load(int nodof, int64_t loaded_nodes, int64_t nn, const int32_t *node, const double *val,
            const int32_t *nf, int64_t ieq_start, int64_t neq_pp, double *r_pp);
/* abaqus2sg (new_library.f90:3515-3682) for hexahedra, in place            */
int pf_abaqus2sg(int nod, int64_t nels, int32_t *g_num);

/* deck readers (formats: SURVEY Appendix A; input.f90:3234-3396 read_p121,
 * :3620-3806 read_p123, :288-445 read_g_coord_pp, :935-1080 read_g_num_pp,
 * :2570-2632 read_rest, :2350-2411 read_loads).  job = path without suffix */
typedef struct {
  int program;      /* 121 or 123 */
  int meshgen, partitioner, nip, nod, limit;
  int64_t nels, nn, nr, loaded, fixed, nres;
  double e, v, kx, ky, kz, tol;
  /* p124 (read_p124, input.f90:3997-4169) and xx2 (read_xx2, :5391-5562) */
  int np_types, nstep, npri, pad_;
  double val0, dtim, theta;
  double rho, cp;   /* p124 material (the .mat file, read_material input.f90:3067-3102) */
} pf_deck_info;
/* program: 121, 123, 124, or 2 for the dev program xx2 (per-element materials) */
int pf_read_dat(const char *job, int program, pf_deck_info *info);
/* pf_read_d + the material number of every element (read_elements, input.f90:1434-1583); etype may be NULL */
int pf_read_d_mat(const char *job, int64_t nn, int64_t nels, int nod,
                  double *g_coord /*(3,nn)*/, int32_t *g_num /*(nod,nels)*/, int32_t *etype /*(nels)*/);
/* <job>.mat (read_material input.f90:3067-3102): prop(nprops,np_types) */
int pf_read_mat(const char *job, int nprops, int np_types, double *prop);
int pf_read_d(const char *job, int64_t nn, int64_t nels, int nod,
              double *g_coord /*(3,nn)*/, int32_t *g_num /*(nod,nels)*/);
int pf_read_bnd(const char *job, int64_t nr, int nodof, int32_t *rest);
int pf_read_lds(const char *job, int64_t loaded, int nodof, int32_t *node, double *val);
/* read_fixed (input.f90:2483-2564): node(fixed), sense(fixed) (1-based freedom of the node), valf(fixed) */
int pf_read_fix(const char *job, int64_t fixed, int32_t *node, int32_t *sense, double *valf);
/* g_coord_pp(nod,3,nels_pp) from g_coord(3,nn) and g_num_pp                */
int pf_coords_pp(int nod, int64_t nels_pp, int64_t nn, const int32_t *g_num_pp,
                 const double *g_coord, double *g_coord_pp);

/* p12meshgen's output side for p121 (p12meshgen.f90:244-323): <job>.d/.bnd/.lds/.dat in the
 * reference's formats ((I12,3E14.6) nodes, Abaqus-ordered 20-node bricks with meshgen = 2,
 * (I15,3I6) restraints, E16.8 loads).  g_coord(3,nn), g_num(nod,nels) in S&G order,
 * rest(nr,4) column-major, node(loaded), val(3,loaded).                      */
int pf_write_deck_p121(const char *job, int nod, int64_t nels, int64_t nn, int64_t nr, int nip,
                       int64_t loaded, double e, double v, double tol, int limit, const double *g_coord,
                       const int32_t *g_num, const int32_t *rest, const int32_t *node, const double *val);

/* p12meshgen's output side for the scalar programs on 8-node bricks: p123 (p12meshgen.f90:703-797), p124
 * (:879-987, plus <job>.mat) and p125 (:1075-1160), selected by info->program; formats as written there
 * (elements in Abaqus order with meshgen = 2; I10 / I12 node fields and I8 / I12 restraint fields differ between
 * the programs).  Reproduces the shipped p124_demo / p125_demo decks byte for byte.  g_coord(3,nn), g_num(8,nels)
 * in S&G order, rest(nr,2) column-major; .lds / .fix as p12meshgen writes them (freedom nres, 10.0 / 100.0) when
 * info->loaded / info->fixed > 0.                                                                    */
int pf_write_deck_scalar(const char *job, const pf_deck_info *info, const double *g_coord,
                         const int32_t *g_num, const int32_t *rest);

/* Output section (p121.f90:124-138): calc_nodes_pp (gather_scatter.f90:2139-2234);
 * nodal values of the owned equations for nodes node_start..node_start+nodes_pp-1
 * (what scatter_nodes, gather_scatter.f90:1786-1929, yields for a conforming field:
 * restrained freedoms 0), out(nodof,nodes_pp); and the EnSight Gold ASCII writer of
 * dismsh_ensi_p (output.f90:2983-3111): 4 header lines, then component-major values,
 * one per line, Fortran e12.<decimals> (5 in the current source, 4 in the shipped
 * p121_demo.ensi.DISPL-000001).                                               */
void pf_calc_nodes_pp(int64_t nn, int npes, int numpe, int64_t *nodes_pp, int64_t *node_start);
/* calc_npes_pp (gather_scatter.f90:349-394): the reference's overestimate of a rank's neighbour count (dimensions of
 * toget / toput before make_ggl).  Not needed by pf_setup_mesh, which counts the neighbours exactly.             */
int pf_calc_npes_pp(int npes);
int pf_nodal_values(int nodof, int64_t nn, const int32_t *nf, int64_t ieq_start, int64_t neq_pp,
                    const double *x_pp, int64_t node_start, int64_t nodes_pp, double *out);
int pf_write_ensi(const char *path, int numvar, int64_t nn, const double *values, int decimals);

/* Halo tables (make_ggl, gather_scatter.f90:1387-1780, rebuilt from g_g_pp).
 * Local numbering of this rank's gather buffer: slot 0 = restrained dump
 * slot, 1..neq_pp = owned equations, then the remote equations the local
 * elements touch, grouped by owner rank ascending, ascending inside a group.
 * ggl_pp(ntot,nels_pp) receives slot numbers; halo_eq receives the global
 * numbers of the remote slots (capacity cap); halo_cnt[npes] per owner.
 * Returns the number of remote slots in *nhalo (call with cap = 0 to size). */
int pf_make_ggl(int ntot, int64_t nels_pp, const int32_t *g_g_pp, int64_t neq,
                int npes, int numpe, int32_t *ggl_pp, int64_t cap,
                int32_t *halo_eq, int64_t *halo_cnt, int64_t *nhalo);

/* binary decks: <job>.bin.ensi.geo = the EnSight Gold "C Binary" geometry file written by
 * p12meshgenbin (mesh_ensi_geo_bin, input.f90:7986-8164) and read by read_g_coord_pp_be (input.f90:632-790) and
 * read_g_num_pp_be (:1254-1420).  Coordinates are single precision in the file; the connectivity is in EnSight's
 * node order (for 8-node bricks the order abaqus2sg expects; pf_ensi2sg restores S&G order for 8 / 20 / 4 nodes). */
int pf_write_geo_bin(const char *job, int nod, int64_t nn, int64_t nels, const double *g_coord, const int32_t *g_num_sg);
int pf_geo_bin_sizes(const char *job, int64_t *nn, int64_t *nels, int *nod);
int pf_read_geo_bin(const char *job, int64_t nn, int64_t nels, int nod, double *g_coord, int32_t *g_num);
int pf_ensi2sg(int nod, int64_t nels, int32_t *g_num);

/* Tables of the halo exchanges fused into the PCG kernels (peer transport, device.cu).  Forward: which owned
 * equations peers gather (one bit each in `bits`, the equations ascending in `slot0`, per equation its
 * destinations rank[ptr[i]..ptr[i+1]) / dst = index in that rank's p_ext) from the put list of pf_setup_mesh
 * (put_slot grouped by destination rank, put_off[nranks+1]) -- replaces the send side of gather's exchange
 * (gather_scatter.f90:600-640).  Reverse: the accumulate entries of every reduction chunk -- the receive side of
 * scatter's exchange (gather_scatter.f90:790-840). */
int pf_make_put_tables(int nranks, int64_t neq_pp, const int64_t *put_off, const int32_t *put_slot,
                       const int64_t *fwd_dst_off, uint32_t *bits, int32_t *slot0, uint32_t *ptr, int32_t *rank,
                       int64_t *dst, int64_t *n_unique);
int pf_make_acc_chunks(int64_t neq_pp, int chunk, int64_t nacc, const int32_t *acc_slot, uint32_t *chunk_ptr);

#ifdef __cplusplus
}
#endif
#endif
