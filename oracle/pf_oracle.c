/*
 * pf_oracle.c -- CPU ORACLE for the EBE-PCG hot path of ParaFEM p121 / p123.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the
 * __graft_entry__.smoke() check and bench.py's cpu_baseline / --impl reference
 * legs may load it.  Nothing under parafem_b200/ links, imports or calls it.
 *
 * It is a plain-C restatement (not a copy) of the reference's algorithm; the
 * reference itself is Fortran 90 + MPI and cannot be compiled in this image
 * (no gfortran, no MPI -- see DESIGN.md); the one compilable source of the
 * reference on this path, xx3/cuda_helpers.cu, is built into oracle/_ref by
 * `make ref` and compared on the GPU in tests/test_gpu_xx3_compat.py.
 * Parity is otherwise pinned against the reference's own golden outputs
 * (tests/golden/, tests/test_oracle_golden.py): xx3-tiny 79 iterations +
 * 756x3 displacements, p121_demo 98 360 equations / 295 iterations / x(1) /
 * centroid stresses / EnSight displacement field, p121 book 777 520 equations,
 * p123 potentials; and for the drivers that reuse the same kernels: p124_demo
 * (all fifteen temperature / iteration rows + nodal temperature files), p125_demo
 * (all ten pressure rows + nodal files), xx2-tiny (five materials: 59 iterations +
 * displacements), xx11 (fixed-freedom path: 125 temperatures).
 *
 * Reference files followed (under /root/reference/parafem/src):
 *   programs/5th_ed/p121/p121.f90 (whole), programs/5th_ed/p123/p123.f90 (whole)
 *   programs/5th_ed/p122/p122.f90:197-231 (orc_p122_elements; the load-increment loop around it is in
 *   oracle/p122_oracle.py), new_library.f90: formm :85-198, invar :1813-1916, mocouf :2364-2417, mocouq :2423-2490
 *   programs/5th_ed/p1210/p1210.f90:93-104 (orc_p1210_mass), :113-150 (orc_p1210_elements, orc_p1210_run),
 *   new_library.f90: vmpl :2196-2286
 *   programs/5th_ed/p124/p124.f90:81-95,139-232, programs/5th_ed/p125/p125.f90:66-99,
 *   programs/dev/xx2/xx2.f90:169-193 (the time loops / material loop are composed from
 *   these C functions in oracle/__init__.py: p124(), p125(), form_km_elastic_mat())
 *   modules/shared/new_library.f90: shape_fun :204-484 (3-D nod = 8 branch :397-422)
 *   modules/shared/new_library.f90: shape_der :745-896, beemat :918-1000,
 *       sample :1397-1520, deemat :1604-1691, rearrange :3059-3112,
 *       rearrange_2 :3118-3124, find_g3 :3130-3212, find_g4 :3249-3271
 *   modules/mpi/maths.f90: dot_product_p :168-216, invert :490-565,
 *       determinant :571-629, checon_par :999-1069
 *   modules/mpi/gather_scatter.f90: calc_nels_pp :146-257, calc_neq_pp :263-343,
 *       gather :547-688, scatter :694-850
 *
 * Arithmetic contract (compile with -O2 -ffp-contract=off, never -ffast-math):
 *   - every MATMUL is C(i,j) = sum_k A(i,k)*B(k,j), k ascending, separate
 *     multiply and add (what gfortran's inlined MATMUL does without FMA);
 *   - scatter adds element contributions in ascending element order inside a
 *     rank (gather_scatter.f90:759-761), the owner adds its own partial sum
 *     first and then the other ranks' partial sums in ASCENDING RANK ORDER (the
 *     reference adds them in arrival order, :829-832 -- any order is a legal
 *     reference execution, this one is reproducible);
 *   - reductions: red_mode 0 = sequential DOT_PRODUCT per rank then ranks
 *     ascending (one legal MPI_ALLREDUCE order); red_mode 1 = the fixed blocked
 *     tree documented at orc_dot_blocked (another legal order, and the one the
 *     CUDA path uses, so GPU-vs-oracle comparisons can be exact).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* element library                                                            */
/* ------------------------------------------------------------------------- */

/* sample('hexahedron'), new_library.f90:1397-1433.  points(nip,3) col-major. */
int orc_sample_hex(int nip, double *points, double *weights) {
  if (nip == 1) {
    points[0] = points[1] = points[2] = 0.0; weights[0] = 8.0; return 0;
  }
  if (nip == 8) {
    const double r = 1.0 / sqrt(3.0);
    /* sign pattern of the 8 points in the reference's order */
    static const int sx[8] = {1, 1, 1, 1, -1, -1, -1, -1};
    static const int sy[8] = {1, 1, -1, -1, 1, -1, 1, -1};
    static const int sz[8] = {1, -1, 1, -1, 1, 1, -1, -1};
    for (int i = 0; i < 8; ++i) {
      points[0 * 8 + i] = sx[i] > 0 ? r : -r;
      points[1 * 8 + i] = sy[i] > 0 ? r : -r;
      points[2 * 8 + i] = sz[i] > 0 ? r : -r;
      weights[i] = 1.0;
    }
    return 0;
  }
  if (nip == 27) {
    /* new_library.f90:1491-1517.  wt = (/5./9.*v,8./9.*v,5./9.*v/): the outer factors are DEFAULT-REAL literals
     * (single-precision quotients, widened), v(9) = (/5/9*w,8/9*w,5/9*w/) with w = (5/9,8/9,5/9) in REAL(iwp) (:1053-1054) */
    const double r15 = 0.2 * sqrt(15.0);
    const double w[3] = {5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0};
    const double f59 = (double)(5.f / 9.f), f89 = (double)(8.f / 9.f);
    double v[9];
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) v[3 * a + b] = w[a] * w[b];   /* (5/9*w, 8/9*w, 5/9*w) */
    for (int blk = 0; blk < 3; ++blk) {
      const double fy = blk == 0 ? -r15 : blk == 1 ? 0.0 : r15, fw = blk == 1 ? f89 : f59;
      for (int q = 0; q < 9; ++q) {
        const int i = 9 * blk + q;                 /* 0-based point */
        points[0 * 27 + i] = (q % 3 == 0) ? -r15 : (q % 3 == 1) ? 0.0 : r15;      /* s(1:7:3,1), s(2:8:3,1), s(3:9:3,1) */
        points[2 * 27 + i] = (q / 3 == 0) ? r15 : (q / 3 == 1) ? 0.0 : -r15;      /* s(1:3,3), s(4:6,3), s(7:9,3) */
        points[1 * 27 + i] = fy;
        weights[i] = fw * v[q];
      }
    }
    return 0;
  }
  return 1;
}

/* sample('tetrahedron') (new_library.f90:1328-1378), weights already multiplied by 1/6: nip = 1 the centroid;
 * nip = 4 and 5 are written there with DEFAULT-REAL (single-precision) literals -- .58541020, .13819660, .25/6.,
 * 1./6., -.8, 9./20. -- which Fortran evaluates in single precision and then widens: restated with floats. */
int orc_sample_tet(int nip, double *points, double *weights) {
#define S(i, j) points[((j)-1) * nip + ((i)-1)]
  if (nip == 1) {
    S(1, 1) = S(1, 2) = S(1, 3) = 0.25; weights[0] = 1.0 / 6.0;
    return 0;
  }
  for (int q = 0; q < 3 * nip; ++q) points[q] = 0.0;
  if (nip == 4) {
    const double a = (double).58541020f, b = (double).13819660f;
    S(1, 1) = a; S(1, 2) = b; S(1, 3) = b;
    S(2, 2) = a; S(2, 3) = b; S(2, 1) = b;
    S(3, 3) = a; S(3, 1) = b; S(3, 2) = b;
    S(4, 1) = b; S(4, 2) = b; S(4, 3) = b;
    const float w = .25f / 6.f;
    for (int i = 0; i < 4; ++i) weights[i] = (double)w;
    return 0;
  }
  if (nip == 5) {
    const float sixth = 1.f / 6.f;
    S(1, 1) = (double).25f; S(1, 2) = (double).25f; S(1, 3) = (double).25f;
    S(2, 1) = (double).5f; S(2, 2) = (double)sixth; S(2, 3) = S(2, 2);
    S(3, 2) = (double).5f; S(3, 3) = (double)sixth; S(3, 1) = S(3, 3);
    S(4, 3) = (double).5f; S(4, 1) = (double)sixth; S(4, 2) = S(4, 1);
    S(5, 1) = (double)sixth; S(5, 2) = S(5, 1); S(5, 3) = S(5, 1);
    weights[0] = (double)(-.8f);
    weights[1] = (double)(9.f / 20.f);
    weights[2] = weights[3] = weights[4] = weights[1];
    for (int i = 0; i < 5; ++i) weights[i] = weights[i] / (double)6.f;      /* wt = wt/6. : REAL(iwp) / default real */
    return 0;
  }
#undef S
  return 1;
}
/* the rule of an element with nod nodes: tetrahedron for nod = 4, hexahedron otherwise */
static int sample_for(int nod, int nip, double *points, double *weights) {
  return nod == 4 ? orc_sample_tet(nip, points, weights) : orc_sample_hex(nip, points, weights);
}

/* shape_der for 3-D nod = 8 / 20 at Gauss point i (0-based); der(3,nod) col-major,
 * new_library.f90:745-755, 769-794, 865-896 */
int orc_shape_der(int nod, const double *points, int nip, int i, double *der) {
  const double xi = points[0 * nip + i], eta = points[1 * nip + i], zeta = points[2 * nip + i];
#define DER(a, l) der[((l)-1) * 3 + ((a)-1)]
  if (nod == 4) {   /* 4-node tetrahedron, new_library.f90:757-767: constant derivatives */
    for (int q = 0; q < 12; ++q) der[q] = 0.0;
    DER(1, 1) = 1.0; DER(2, 2) = 1.0; DER(3, 3) = 1.0;
    DER(1, 4) = -1.0; DER(2, 4) = -1.0; DER(3, 4) = -1.0;
    return 0;
  }
  if (nod == 8) {
    const double etam = 1.0 - eta, xim = 1.0 - xi, zetam = 1.0 - zeta;
    const double etap = eta + 1.0, xip = xi + 1.0, zetap = zeta + 1.0;
    DER(1, 1) = -0.125 * etam * zetam; DER(1, 2) = -0.125 * etam * zetap;
    DER(1, 3) = 0.125 * etam * zetap;  DER(1, 4) = 0.125 * etam * zetam;
    DER(1, 5) = -0.125 * etap * zetam; DER(1, 6) = -0.125 * etap * zetap;
    DER(1, 7) = 0.125 * etap * zetap;  DER(1, 8) = 0.125 * etap * zetam;
    DER(2, 1) = -0.125 * xim * zetam;  DER(2, 2) = -0.125 * xim * zetap;
    DER(2, 3) = -0.125 * xip * zetap;  DER(2, 4) = -0.125 * xip * zetam;
    DER(2, 5) = 0.125 * xim * zetam;   DER(2, 6) = 0.125 * xim * zetap;
    DER(2, 7) = 0.125 * xip * zetap;   DER(2, 8) = 0.125 * xip * zetam;
    DER(3, 1) = -0.125 * xim * etam;   DER(3, 2) = 0.125 * xim * etam;
    DER(3, 3) = 0.125 * xip * etam;    DER(3, 4) = -0.125 * xip * etam;
    DER(3, 5) = -0.125 * xim * etap;   DER(3, 6) = 0.125 * xim * etap;
    DER(3, 7) = 0.125 * xip * etap;    DER(3, 8) = -0.125 * xip * etap;
    return 0;
  }
  if (nod == 20) {
    static const int xii[20] = {-1, -1, -1, 0, 1, 1, 1, 0, -1, -1, 1, 1, -1, -1, -1, 0, 1, 1, 1, 0};
    static const int etai[20] = {-1, -1, -1, -1, -1, -1, -1, -1, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1};
    static const int zetai[20] = {-1, 0, 1, 1, 1, 0, -1, -1, -1, 1, 1, -1, -1, 0, 1, 1, 1, 0, -1, -1};
    for (int l = 1; l <= 20; ++l) {
      const double xl = xii[l - 1], el = etai[l - 1], zl = zetai[l - 1];
      const double xi0 = xi * xl, eta0 = eta * el, zeta0 = zeta * zl;
      if (l == 4 || l == 8 || l == 16 || l == 20) {
        DER(1, l) = -.5 * xi * (1. + eta0) * (1. + zeta0);
        DER(2, l) = .25 * el * (1. - xi * xi) * (1. + zeta0);
        DER(3, l) = .25 * zl * (1. - xi * xi) * (1. + eta0);
      } else if (l >= 9 && l <= 12) {
        DER(1, l) = .25 * xl * (1. - eta * eta) * (1. + zeta0);
        DER(2, l) = -.5 * eta * (1. + xi0) * (1. + zeta0);
        DER(3, l) = .25 * zl * (1. + xi0) * (1. - eta * eta);
      } else if (l == 2 || l == 6 || l == 14 || l == 18) {
        DER(1, l) = .25 * xl * (1. + eta0) * (1. - zeta * zeta);
        DER(2, l) = .25 * el * (1. + xi0) * (1. - zeta * zeta);
        DER(3, l) = -.5 * zeta * (1. + xi0) * (1. + eta0);
      } else {
        DER(1, l) = .125 * xl * (1. + eta0) * (1. + zeta0) * (2. * xi0 + eta0 + zeta0 - 1.);
        DER(2, l) = .125 * el * (1. + xi0) * (1. + zeta0) * (xi0 + 2. * eta0 + zeta0 - 1.);
        DER(3, l) = .125 * zl * (1. + xi0) * (1. + eta0) * (xi0 + eta0 + 2. * zeta0 - 1.);
      }
    }
    return 0;
  }
#undef DER
  return 1;
}

#define M3(m, r, c) (m)[((c)-1) * 3 + ((r)-1)]

/* determinant, 3x3 branch, maths.f90:617-619 */
double orc_determinant3(const double *jac) {
  double det = M3(jac, 1, 1) * (M3(jac, 2, 2) * M3(jac, 3, 3) - M3(jac, 3, 2) * M3(jac, 2, 3));
  det = det - M3(jac, 1, 2) * (M3(jac, 2, 1) * M3(jac, 3, 3) - M3(jac, 3, 1) * M3(jac, 2, 3));
  det = det + M3(jac, 1, 3) * (M3(jac, 2, 1) * M3(jac, 3, 2) - M3(jac, 3, 1) * M3(jac, 2, 2));
  return det;
}

/* invert, 3x3 branch (adjugate / det), maths.f90:526-547 */
void orc_invert3(double *m) {
  double det = orc_determinant3(m);
  double j11 = M3(m, 2, 2) * M3(m, 3, 3) - M3(m, 3, 2) * M3(m, 2, 3);
  double j21 = -(M3(m, 2, 1) * M3(m, 3, 3)) + M3(m, 3, 1) * M3(m, 2, 3);
  double j31 = M3(m, 2, 1) * M3(m, 3, 2) - M3(m, 3, 1) * M3(m, 2, 2);
  double j12 = -(M3(m, 1, 2) * M3(m, 3, 3)) + M3(m, 3, 2) * M3(m, 1, 3);
  double j22 = M3(m, 1, 1) * M3(m, 3, 3) - M3(m, 3, 1) * M3(m, 1, 3);
  double j32 = -(M3(m, 1, 1) * M3(m, 3, 2)) + M3(m, 3, 1) * M3(m, 1, 2);
  double j13 = M3(m, 1, 2) * M3(m, 2, 3) - M3(m, 2, 2) * M3(m, 1, 3);
  double j23 = -(M3(m, 1, 1) * M3(m, 2, 3)) + M3(m, 2, 1) * M3(m, 1, 3);
  double j33 = M3(m, 1, 1) * M3(m, 2, 2) - M3(m, 2, 1) * M3(m, 1, 2);
  M3(m, 1, 1) = j11 / det; M3(m, 1, 2) = j12 / det; M3(m, 1, 3) = j13 / det;
  M3(m, 2, 1) = j21 / det; M3(m, 2, 2) = j22 / det; M3(m, 2, 3) = j23 / det;
  M3(m, 3, 1) = j31 / det; M3(m, 3, 2) = j32 / det; M3(m, 3, 3) = j33 / det;
}

/* deemat, nst = 6 branch, new_library.f90:1671-1686; dee(6,6) col-major */
void orc_deemat6(double *dee, double e, double v) {
  const double v2 = v / (1.0 - v);
  const double vv = (1.0 - 2.0 * v) / (1.0 - v) * 0.5;
  memset(dee, 0, 36 * sizeof(double));
  for (int i = 0; i < 3; ++i) dee[i * 6 + i] = 1.0;
  for (int i = 3; i < 6; ++i) dee[i * 6 + i] = vv;
  dee[1 * 6 + 0] = v2; dee[0 * 6 + 1] = v2; dee[2 * 6 + 0] = v2;
  dee[0 * 6 + 2] = v2; dee[2 * 6 + 1] = v2; dee[1 * 6 + 2] = v2;
  const double den = 2.0 * (1.0 + v) * vv;
  for (int i = 0; i < 36; ++i) dee[i] = dee[i] * e / den;
}

/* beemat, nst = 6 branch, new_library.f90:976-993; bee(6,3*nod) col-major */
void orc_beemat6(double *bee, const double *deriv, int nod) {
  memset(bee, 0, (size_t)6 * 3 * nod * sizeof(double));
#define BEE(r, c) bee[((c)-1) * 6 + ((r)-1)]
  for (int m = 1; m <= nod; ++m) {
    const int n = 3 * m, k = n - 1, l = k - 1;
    const double x = deriv[(m - 1) * 3 + 0], y = deriv[(m - 1) * 3 + 1], z = deriv[(m - 1) * 3 + 2];
    BEE(1, l) = x; BEE(4, k) = x; BEE(6, n) = x;
    BEE(2, k) = y; BEE(4, l) = y; BEE(5, n) = y;
    BEE(3, n) = z; BEE(5, k) = z; BEE(6, l) = z;
  }
#undef BEE
}

/* Cartesian derivatives at one Gauss point: shape_der, jac = der*coord,
 * det, invert, deriv = jac^-1 * der  (p121.f90:58-59).  coord(nod,3). */
static double gauss_point(int nod, const double *points, int nip, int ig, const double *coord,
                          double *der, double *deriv) {
  double jac[9];
  orc_shape_der(nod, points, nip, ig, der);
  for (int b = 0; b < 3; ++b)
    for (int a = 0; a < 3; ++a) {
      double s = 0.0;
      for (int m = 0; m < nod; ++m) s += der[m * 3 + a] * coord[b * nod + m];
      jac[b * 3 + a] = s;
    }
  double det = orc_determinant3(jac);
  orc_invert3(jac);
  for (int m = 0; m < nod; ++m)
    for (int a = 0; a < 3; ++a) {
      double s = 0.0;
      for (int b = 0; b < 3; ++b) s += jac[b * 3 + a] * der[m * 3 + b];
      deriv[m * 3 + a] = s;
    }
  return det;
}

/* elements_1 / gauss_pts_1 of p121.f90:54-64.
 * g_coord_pp(nod,3,nels), storkm_pp(ntot,ntot,nels), ntot = 3*nod. */
int orc_form_km_elastic(int64_t nels, int nod, int nip, const double *g_coord_pp, double e,
                        double v, double *storkm_pp) {
  if ((nod != 4 && nod != 8 && nod != 20) || (nod != 4 && nip != 1 && nip != 8 && nip != 27) || (nod == 4 && nip != 1 && nip != 4 && nip != 5)) return 1;
  const int ntot = 3 * nod;
  double points[81], weights[27], dee[36];
  sample_for(nod, nip, points, weights);
  orc_deemat6(dee, e, v);
#pragma omp parallel
  {
    double der[60], deriv[60];
    double *bee = malloc(sizeof(double) * 6 * ntot);
    double *btd = malloc(sizeof(double) * ntot * 6);
#pragma omp for schedule(static)
    for (int64_t iel = 0; iel < nels; ++iel) {
      double *km = storkm_pp + iel * ntot * ntot;
      memset(km, 0, sizeof(double) * ntot * ntot);
      for (int ig = 0; ig < nip; ++ig) {
        const double det = gauss_point(nod, points, nip, ig, g_coord_pp + iel * nod * 3, der, deriv);
        orc_beemat6(bee, deriv, nod);
        /* btd = MATMUL(TRANSPOSE(bee),dee): (ntot,6) col-major */
        for (int k = 0; k < 6; ++k)
          for (int i = 0; i < ntot; ++i) {
            double s = 0.0;
            for (int l = 0; l < 6; ++l) s += bee[i * 6 + l] * dee[k * 6 + l];
            btd[k * ntot + i] = s;
          }
        /* km += MATMUL(btd,bee)*det*weights(ig) */
        for (int j = 0; j < ntot; ++j)
          for (int i = 0; i < ntot; ++i) {
            double s = 0.0;
            for (int k = 0; k < 6; ++k) s += btd[k * ntot + i] * bee[j * 6 + k];
            km[j * ntot + i] = km[j * ntot + i] + s * det * weights[ig];
          }
      }
    }
    free(bee); free(btd);
  }
  return 0;
}

/* elements_1 of p123.f90:70-84: kcx,kcy,kcz outer products; storkc_pp(8,8,nels) */
int orc_form_kc_laplace(int64_t nels, int nod, int nip, const double *g_coord_pp, double kx,
                        double ky, double kz, double *storkc_pp) {
  if ((nod != 8 && nod != 4) || (nod == 8 && nip != 1 && nip != 8) || (nod == 4 && nip != 1 && nip != 4 && nip != 5)) return 1;
  double points[81], weights[27];
  sample_for(nod, nip, points, weights);
  const int nn2 = nod * nod;
#pragma omp parallel for schedule(static)
  for (int64_t iel = 0; iel < nels; ++iel) {
    double der[24], deriv[24], kc[3][64];
    memset(kc, 0, sizeof kc);
    for (int ig = 0; ig < nip; ++ig) {
      const double det = gauss_point(nod, points, nip, ig, g_coord_pp + iel * nod * 3, der, deriv);
      for (int a = 0; a < 3; ++a)
        for (int j = 0; j < nod; ++j)
          for (int i = 0; i < nod; ++i)
            kc[a][j * nod + i] = kc[a][j * nod + i] + deriv[i * 3 + a] * deriv[j * 3 + a] * det * weights[ig];
    }
    double *out = storkc_pp + iel * nn2;
    for (int q = 0; q < nn2; ++q) out[q] = kc[0][q] * kx + kc[1][q] * ky + kc[2][q] * kz;
  }
  return 0;
}

/* shape_fun, 3-D nod = 8 (new_library.f90:397-422) at Gauss point i (0-based) */
int orc_shape_fun(int nod, const double *points, int nip, int i, double *fun) {
  if (nod != 8 && nod != 20) return 1;
  const double xi = points[0 * nip + i], eta = points[1 * nip + i], zeta = points[2 * nip + i];
  if (nod == 20) {   /* new_library.f90:449-468 */
    static const int xii[20] = {-1, -1, -1, 0, 1, 1, 1, 0, -1, -1, 1, 1, -1, -1, -1, 0, 1, 1, 1, 0};
    static const int etai[20] = {-1, -1, -1, -1, -1, -1, -1, -1, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1};
    static const int zetai[20] = {-1, 0, 1, 1, 1, 0, -1, -1, -1, 1, 1, -1, -1, 0, 1, 1, 1, 0, -1, -1};
    for (int l = 1; l <= 20; ++l) {
      const double xi0 = xi * xii[l - 1], eta0 = eta * etai[l - 1], zeta0 = zeta * zetai[l - 1];
      if (l == 4 || l == 8 || l == 16 || l == 20) fun[l - 1] = .25 * (1. - xi * xi) * (1. + eta0) * (1. + zeta0);
      else if (l >= 9 && l <= 12) fun[l - 1] = .25 * (1. + xi0) * (1. - eta * eta) * (1. + zeta0);
      else if (l == 2 || l == 6 || l == 14 || l == 18) fun[l - 1] = .25 * (1. + xi0) * (1. + eta0) * (1. - zeta * zeta);
      else fun[l - 1] = .125 * (1. + xi0) * (1. + eta0) * (1. + zeta0) * (xi0 + eta0 + zeta0 - 2);
    }
    return 0;
  }
  const double etam = 1.0 - eta, xim = 1.0 - xi, zetam = 1.0 - zeta;
  const double etap = eta + 1.0, xip = xi + 1.0, zetap = zeta + 1.0;
  fun[0] = 0.125 * xim * etam * zetam; fun[1] = 0.125 * xim * etam * zetap;
  fun[2] = 0.125 * xip * etam * zetap; fun[3] = 0.125 * xip * etam * zetam;
  fun[4] = 0.125 * xim * etap * zetam; fun[5] = 0.125 * xim * etap * zetap;
  fun[6] = 0.125 * xip * etap * zetap; fun[7] = 0.125 * xip * etap * zetam;
  return 0;
}

/* elements_3 / gauss_pts of p124.f90:81-95 (also p125.f90's kc / pm):
 *   kc += MATMUL(MATMUL(TRANSPOSE(deriv),kay),deriv)*det*weights(i)
 *   pm += MATMUL(TRANSPOSE(funny),funny)*det*weights(i)*rho*cp
 *   storka = pm + kc*theta*dtim ; storkb = pm - kc*(1-theta)*dtim
 * kay = diag(kx,ky,kz).  storka/storkb (8,8,nels); either may be NULL.  kc_out / pm_out (may be
 * NULL) receive the raw kc and pm (p125 uses them directly). */
int orc_form_k_transient(int64_t nels, int nod, int nip, const double *g_coord_pp, double kx,
                         double ky, double kz, double rho, double cp, double theta, double dtim,
                         double *storka_pp, double *storkb_pp, double *kc_out, double *pm_out) {
  if (nod != 8 || (nip != 1 && nip != 8)) return 1;
  double points[81], weights[27], kay[9] = {0};
  orc_sample_hex(nip, points, weights);
  kay[0] = kx; kay[4] = ky; kay[8] = kz;                       /* kay(a,b) at [b*3+a] */
  const double omt = 1.0 - theta;
#pragma omp parallel for schedule(static)
  for (int64_t iel = 0; iel < nels; ++iel) {
    double der[24], deriv[24], fun[8], kc[64], pm[64], t1[24];
    memset(kc, 0, sizeof kc); memset(pm, 0, sizeof pm);
    for (int ig = 0; ig < nip; ++ig) {
      const double det = gauss_point(nod, points, nip, ig, g_coord_pp + iel * nod * 3, der, deriv);
      orc_shape_fun(nod, points, nip, ig, fun);
      for (int b = 0; b < 3; ++b)                               /* t1 = MATMUL(TRANSPOSE(deriv),kay): (nod,3) */
        for (int m = 0; m < 8; ++m) {
          double s = 0.0;
          for (int a = 0; a < 3; ++a) s += deriv[m * 3 + a] * kay[b * 3 + a];
          t1[b * 8 + m] = s;
        }
      for (int j = 0; j < 8; ++j)
        for (int i = 0; i < 8; ++i) {
          double s = 0.0;
          for (int b = 0; b < 3; ++b) s += t1[b * 8 + i] * deriv[j * 3 + b];
          kc[j * 8 + i] = kc[j * 8 + i] + s * det * weights[ig];
          double f = 0.0;
          f += fun[i] * fun[j];
          pm[j * 8 + i] = pm[j * 8 + i] + f * det * weights[ig] * rho * cp;
        }
    }
    for (int q = 0; q < 64; ++q) {
      if (storka_pp) storka_pp[iel * 64 + q] = pm[q] + kc[q] * theta * dtim;
      if (storkb_pp) storkb_pp[iel * 64 + q] = pm[q] - kc[q] * omt * dtim;
      if (kc_out) kc_out[iel * 64 + q] = kc[q];
      if (pm_out) pm_out[iel * 64 + q] = pm[q];
    }
  }
  return 0;
}

/* elements_2 of p129.f90:85-98, consistent mass: emm = sum_gp ecmat(fun)*det*w*rho with ecmat = MATMUL(nt,tn)
 * (new_library.f90:1536-1563: nt((i-1)*nodof+j,j) = fun(i)), i.e. emm(3i+j,3i'+j') = fun(i)*fun(i') for j == j'.
 * store_mm_pp(ntot,ntot,nels). */
int orc_form_mass(int64_t nels, int nod, int nip, const double *g_coord_pp, double rho, double *store_mm_pp) {
  if ((nod != 8 && nod != 20) || (nip != 8 && nip != 27)) return 1;
  const int ntot = 3 * nod;
  double points[81], weights[27];
  orc_sample_hex(nip, points, weights);
#pragma omp parallel for schedule(static)
  for (int64_t iel = 0; iel < nels; ++iel) {
    double der[60], deriv[60], fun[20];
    double *mm = store_mm_pp + iel * ntot * ntot;
    for (int q = 0; q < ntot * ntot; ++q) mm[q] = 0.0;
    for (int ig = 0; ig < nip; ++ig) {
      const double det = gauss_point(nod, points, nip, ig, g_coord_pp + iel * nod * 3, der, deriv);
      orc_shape_fun(nod, points, nip, ig, fun);
      for (int b = 0; b < ntot; ++b)
        for (int a = 0; a < ntot; ++a) {
          double ecm = 0.0;                              /* MATMUL(nt,tn): k ascending, separate multiply and add */
          for (int k = 0; k < 3; ++k) {
            const double nt = (a % 3 == k) ? fun[a / 3] : 0.0, tn = (b % 3 == k) ? fun[b / 3] : 0.0;
            ecm = ecm + nt * tn;
          }
          ecm = ecm * det * weights[ig] * rho;
          mm[b * ntot + a] = mm[b * ntot + a] + ecm;
        }
    }
  }
  return 0;
}

/* centroid stresses, p121.f90:113-123: one point at (0,0,0); sigma = dee*(bee*eld) */
int orc_point_stress(int nod, const double *coord, const double *eld, double e, double v, double xi, double eta,
                     double zeta, double *sigma);
int orc_centroid_stress(int nod, const double *coord, const double *eld, double e, double v,
                        double *sigma) {
  return orc_point_stress(nod, coord, eld, e, v, 0.0, 0.0, 0.0, sigma);
}
/* the same at any local point: the 2013 build of p121 behind examples/5th_ed/p121/book/p121.res printed the
 * stress at the last point of the 8-point rule, not at the centroid */
int orc_point_stress(int nod, const double *coord, const double *eld, double e, double v, double xi, double eta,
                     double zeta, double *sigma) {
  if (nod != 4 && nod != 8 && nod != 20) return 1;
  const int ntot = 3 * nod;
  double points[3] = {xi, eta, zeta}, der[60], deriv[60], dee[36], eps[6];
  double *bee = malloc(sizeof(double) * 6 * ntot);
  gauss_point(nod, points, 1, 0, coord, der, deriv);
  orc_beemat6(bee, deriv, nod);
  orc_deemat6(dee, e, v);
  for (int r = 0; r < 6; ++r) {
    double s = 0.0;
    for (int c = 0; c < ntot; ++c) s += bee[c * 6 + r] * eld[c];
    eps[r] = s;
  }
  for (int r = 0; r < 6; ++r) {
    double s = 0.0;
    for (int c = 0; c < 6; ++c) s += dee[c * 6 + r] * eps[c];
    sigma[r] = s;
  }
  free(bee);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* steering: rearrange + find_g3, rearrange_2 + find_g4                       */
/* ------------------------------------------------------------------------- */


/* ------------------------------------------------------------------------- */
/* p122 (elasto-plasticity): the Gauss-point update of elements_4, p122.f90:199-229 */
/* ------------------------------------------------------------------------- */
/* invar, nst = 6 (new_library.f90:1893-1912) */
static void invar6(const double *s, double *sigm, double *dsbar, double *theta) {
  const double sq3 = sqrt(3.0);
  *sigm = (s[0] + s[1] + s[2]) / 3.0;
  const double d2 = ((s[0] - s[1]) * (s[0] - s[1]) + (s[1] - s[2]) * (s[1] - s[2]) + (s[2] - s[0]) * (s[2] - s[0])) / 6.0 +
                    s[3] * s[3] + s[4] * s[4] + s[5] * s[5];
  const double ds1 = s[0] - *sigm, ds2 = s[1] - *sigm, ds3 = s[2] - *sigm;
  const double d3 = ds1 * ds2 * ds3 - ds1 * s[4] * s[4] - ds2 * s[5] * s[5] - ds3 * s[3] * s[3] + 2.0 * s[3] * s[4] * s[5];
  *dsbar = sq3 * sqrt(d2);
  if (*dsbar < 1e-10) *theta = 0.0;
  else {
    const double r = sqrt(d2);
    double sine = -3.0 * sq3 * d3 / (2.0 * r * r * r);
    if (sine > 1.0) sine = 1.0;
    if (sine < -1.0) sine = -1.0;
    *theta = asin(sine) / 3.0;
  }
}
/* mocouf (new_library.f90:2364-2417) */
static double mocouf(double phi, double c, double sigm, double dsbar, double theta) {
  const double phir = phi * 4.0 * atan(1.0) / 180.0;
  const double snph = sin(phir), csph = cos(phir), csth = cos(theta), snth = sin(theta);
  return snph * sigm + dsbar * (csth / sqrt(3.0) - snth * snph / 3.0) - c * csph;
}
/* mocouq (new_library.f90:2423-2490) */
static void mocouq(double psi, double dsbar, double theta, double *dq1, double *dq2, double *dq3) {
  const double psir = psi * 4.0 * atan(1.0) / 180.0;
  const double snth = sin(theta), snps = sin(psir), sq3 = sqrt(3.0);
  *dq1 = snps;
  if (fabs(snth) > 0.49) {
    const double c1 = snth < 0.0 ? -1.0 : 1.0;
    *dq2 = (sq3 * 0.5 - c1 * snps * 0.5 / sq3) * sq3 * 0.5 / dsbar;
    *dq3 = 0.0;
  } else {
    const double csth = cos(theta), cs3th = cos(3.0 * theta), tn3th = tan(3.0 * theta), tnth = snth / csth;
    *dq2 = sq3 * csth / dsbar * ((1.0 + tnth * tn3th) + snps * (tn3th - tnth) / sq3) * 0.5;
    *dq3 = 0.5 * 3.0 * (sq3 * snth + snps * csth) / (cs3th * dsbar * dsbar);
  }
}
/* formm, nst = 6 (new_library.f90:145-194); m(i,j) at [j*6+i] */
static void formm6(const double *st, double *m1, double *m2, double *m3) {
  const double sx = st[0], sy = st[1], sz = st[2], txy = st[3], tyz = st[4], tzx = st[5];
  const double sigm = (sx + sy + sz) / 3.0, dx = sx - sigm, dy = sy - sigm, dz = sz - sigm;
  memset(m1, 0, 36 * sizeof(double)); memset(m2, 0, 36 * sizeof(double)); memset(m3, 0, 36 * sizeof(double));
#define M(m, i, j) (m)[((j)-1) * 6 + ((i)-1)]
  for (int i = 1; i <= 3; ++i) for (int j = 1; j <= 3; ++j) M(m1, i, j) = 1.0 / (3.0 * sigm);
  for (int i = 1; i <= 3; ++i) { M(m2, i, i) = 2.0; M(m2, i + 3, i + 3) = 6.0; }
  M(m2, 1, 2) = -1.0; M(m2, 1, 3) = -1.0; M(m2, 2, 3) = -1.0;
  M(m3, 1, 1) = dx; M(m3, 1, 2) = dz; M(m3, 1, 3) = dy; M(m3, 1, 4) = txy; M(m3, 1, 5) = -2.0 * tyz; M(m3, 1, 6) = tzx;
  M(m3, 2, 2) = dy; M(m3, 2, 3) = dx; M(m3, 2, 4) = txy; M(m3, 2, 5) = tyz; M(m3, 2, 6) = -2.0 * tzx;
  M(m3, 3, 3) = dz; M(m3, 3, 4) = -2.0 * txy; M(m3, 3, 5) = tyz; M(m3, 3, 6) = tzx;
  M(m3, 4, 4) = -3.0 * dz; M(m3, 4, 5) = 3.0 * tzx; M(m3, 4, 6) = 3.0 * tyz;
  M(m3, 5, 5) = -3.0 * dx; M(m3, 5, 6) = 3.0 * txy; M(m3, 6, 6) = -3.0 * dy;
  for (int i = 1; i <= 6; ++i)
    for (int j = i + 1; j <= 6; ++j) { M(m1, j, i) = M(m1, i, j); M(m2, j, i) = M(m2, i, j); M(m3, j, i) = M(m3, i, j); }
  for (int q = 0; q < 36; ++q) { m1[q] = m1[q] / 3.0; m2[q] = m2[q] / 3.0; m3[q] = m3[q] / 3.0; }
#undef M
}

/* elements_4 of p122.f90:197-231 over all elements: pmul (ntot,nels) = gathered displacement increment;
 * evpt, tensor (6,nip,nels) updated in place; utemp (ntot,nels) receives bload.  last = plastic_converged .OR.
 * plasiters == plasits.  Every MATMUL is the k-ascending sum from 0.0 of the arithmetic contract above. */
int orc_p122_elements(int64_t nels, int nod, int nip, const double *g_coord_pp, double e, double v, double phi,
                      double c, double psi, double dt, int last, const double *pmul, double *evpt, double *tensor,
                      double *utemp) {
  if ((nod != 8 && nod != 20) || nip != 8) return 1;
  const int ntot = 3 * nod;
  double points[81], weights[27], dee[36];
  orc_sample_hex(nip, points, weights);
  orc_deemat6(dee, e, v);
#pragma omp parallel for schedule(static)
  for (int64_t iel = 0; iel < nels; ++iel) {
    double der[60], deriv[60], bee[6 * 60], bload[60], eps[6], sigma[6], stress[6], devp[6] = {0}, evp[6], erate[6];
    double m1[36], m2[36], m3[36], flow[36];
    const double *eld = pmul + iel * ntot;
    for (int q = 0; q < ntot; ++q) bload[q] = 0.0;
    for (int ig = 0; ig < nip; ++ig) {
      double *ev = evpt + (iel * nip + ig) * 6, *te = tensor + (iel * nip + ig) * 6;
      const double det = gauss_point(nod, points, nip, ig, g_coord_pp + iel * nod * 3, der, deriv);
      orc_beemat6(bee, deriv, nod);
      for (int r = 0; r < 6; ++r) {
        double s = 0.0;
        for (int q = 0; q < ntot; ++q) s += bee[q * 6 + r] * eld[q];
        eps[r] = s - ev[r];
      }
      for (int r = 0; r < 6; ++r) {
        double s = 0.0;
        for (int q = 0; q < 6; ++q) s += dee[q * 6 + r] * eps[q];
        sigma[r] = s;
        stress[r] = s + te[r];
      }
      double sigm, dsbar, theta;
      invar6(stress, &sigm, &dsbar, &theta);
      const double f = mocouf(phi, c, sigm, dsbar, theta);
      if (last) {
        for (int r = 0; r < 6; ++r) devp[r] = stress[r];
      } else if (f >= 0.0) {
        double dq1, dq2, dq3;
        mocouq(psi, dsbar, theta, &dq1, &dq2, &dq3);
        formm6(stress, m1, m2, m3);
        for (int q = 0; q < 36; ++q) flow[q] = f * (m1[q] * dq1 + m2[q] * dq2 + m3[q] * dq3);
        for (int r = 0; r < 6; ++r) {
          double s = 0.0;
          for (int q = 0; q < 6; ++q) s += flow[q * 6 + r] * stress[q];
          erate[r] = s;
          evp[r] = s * dt;
          ev[r] = ev[r] + evp[r];
        }
        for (int r = 0; r < 6; ++r) {
          double s = 0.0;
          for (int q = 0; q < 6; ++q) s += dee[q * 6 + r] * evp[q];
          devp[r] = s;
        }
      }
      if (f >= 0.0)
        for (int q = 0; q < ntot; ++q) {
          double s = 0.0;
          for (int r = 0; r < 6; ++r) s += bee[q * 6 + r] * devp[r];
          bload[q] = bload[q] + s * det * weights[ig];
        }
      if (last) for (int r = 0; r < 6; ++r) te[r] = stress[r];
      (void)sigma; (void)erate;
    }
    for (int q = 0; q < ntot; ++q) utemp[iel * ntot + q] = bload[q];
  }
  return 0;
}

/* rest(nr,nodof+1) col-major, modified in place (new_library.f90:3059-3112) */
void orc_rearrange(int64_t nr, int nodof, int32_t *rest) {
#define REST(i, k) rest[(int64_t)((k)-1) * nr + ((i)-1)]
  int32_t m = 0;
  for (int64_t i = 1; i <= nr; ++i)
    for (int k = 2; k <= nodof + 1; ++k)
      if (REST(i, k) != 0) { m = m + 1; REST(i, k) = m; }
  for (int64_t i = 1; i <= nr; ++i) {
    int64_t k = REST(i, 1);
    for (int c = 2; c <= nodof + 1; ++c)
      if (REST(i, c) != 0) REST(i, c) = (int32_t)(REST(i, c) + nodof * (k - i));
  }
}

static int64_t rest_bsearch(int64_t nr, const int32_t *rest, int64_t l) {
  int64_t first = 1, last = nr;
  while (first != last) {
    int64_t half = (first + last) / 2;
    if (l <= REST(half, 1)) last = half; else first = half + 1;
  }
  return first;
}

/* find_g3 (new_library.f90:3130-3212) for one element; rest already rearranged */
void orc_find_g3(int nod, int nodof, const int32_t *num, int32_t *g, int64_t nr, const int32_t *rest) {
  for (int i = 1; i <= nod; ++i) {
    const int64_t l = num[i - 1];
    const int64_t only = rest_bsearch(nr, rest, l);
    if (l == REST(only, 1)) {
      for (int j = 1; j <= nodof; ++j) g[nodof * i - nodof + j - 1] = REST(only, j + 1);
    } else {
      int64_t k = (only == nr) ? 0 : 1, s1;
      for (;;) {
        s1 = only - k;
        int64_t sum = 0;
        for (int c = 2; c <= nodof + 1; ++c) sum += REST(s1, c);
        if (sum != 0 || s1 == 1) break;
        k = k + 1;
      }
      int64_t s2 = REST(s1, 2);
      for (int c = 3; c <= nodof + 1; ++c) if (REST(s1, c) > s2) s2 = REST(s1, c);
      int64_t s3 = (only == nr) ? l - REST(s1, 1) - k : l - REST(s1, 1) - (k - 1);
      for (int j = 1; j <= nodof; ++j) g[nodof * i - nodof + j - 1] = (int32_t)(s2 + nodof * s3 - (nodof - j));
    }
  }
}

/* elements_0 of p121.f90:43-45 / p123.f90: find_g3 (nodof 3) or find_g4 (nodof 1) over n elements */
void orc_find_g4(int nod, const int32_t *num, int32_t *g, int64_t nr, const int32_t *rest);
void orc_find_g_all(int nod, int nodof, int64_t n, const int32_t *g_num, int32_t *g_g, int64_t nr, const int32_t *rest) {
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < n; ++e) {
    if (nodof == 1) orc_find_g4(nod, g_num + e * nod, g_g + e * nod, nr, rest);
    else orc_find_g3(nod, nodof, g_num + e * nod, g_g + e * nod * nodof, nr, rest);
  }
}

/* rearrange_2 (new_library.f90:3118-3124); rest(nr,2) */
void orc_rearrange_2(int64_t nr, int32_t *rest) {
  int64_t m = 0;
  for (int64_t i = 1; i <= nr; ++i) REST(i, 2) = (int32_t)(REST(i, 1) - i);
  for (int64_t i = 1; i <= nr; ++i) if (REST(i, 2) != 0) { m = i - 1; break; }
  for (int64_t i = (m < 1 ? 1 : m); i <= nr; ++i) REST(i, 2) = REST(i, 2) + 1;
}

/* find_g4 (new_library.f90:3249-3271) for one element */
void orc_find_g4(int nod, const int32_t *num, int32_t *g, int64_t nr, const int32_t *rest) {
  for (int i = 1; i <= nod; ++i) {
    const int64_t l = num[i - 1];
    const int64_t only = rest_bsearch(nr, rest, l);
    if (l == REST(only, 1)) g[i - 1] = 0;
    else {
      int64_t s1 = only - 1;
      g[i - 1] = (int32_t)(REST(s1, 2) + (l - REST(s1, 1)) - 1);
    }
  }
}
#undef REST

/* ------------------------------------------------------------------------- */
/* p12meshgen cubes (tools/preprocessing/p12meshgen/p12meshgen.f90) -- the mesh  */
/* of the reference arm of bench.py and of the full-size parity tests, built     */
/* without the product's library.  In-memory (full precision), S&G node order:   */
/* what p121 / p123 hold after read_g_num_pp + abaqus2sg + read_g_coord_pp when  */
/* the deck's E14.6 rounding is left out.                                        */
/* ------------------------------------------------------------------------- */

/* geometry_8bxz (geometry.f90:118-166) / geometry_20bxz (:230-289) for elements iel0+1 .. iel0+n:
 * g_num(nod,n) and g_coord_pp(nod,3,n) -- coord(m,b) of element e at [e*3*nod + b*nod + m] */
int orc_cube_elements(int nod, int nxe, int nze, double aa, double bb, double cc, int64_t iel0, int64_t n,
                      int32_t *g_num, double *g_coord_pp) {
  if (nod != 8 && nod != 20) return 1;
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < n; ++q) {
    const int64_t iel = iel0 + q + 1;
    const int64_t iq = (iel - 1) / ((int64_t)nxe * nze) + 1;
    const int64_t iplane = iel - (iq - 1) * (int64_t)nxe * nze;
    const int64_t is = (iplane - 1) / nxe + 1;
    const int64_t ip = iplane - (is - 1) * nxe;
    int64_t num[21];
    double x[21], y[21], z[21];
    if (nod == 8) {
      num[1] = (iq - 1) * (int64_t)(nxe + 1) * (nze + 1) + is * (nxe + 1) + ip;
      num[2] = num[1] - nxe - 1; num[3] = num[2] + 1; num[4] = num[1] + 1;
      num[5] = num[1] + (int64_t)(nxe + 1) * (nze + 1);
      num[6] = num[5] - nxe - 1; num[7] = num[6] + 1; num[8] = num[5] + 1;
      const double x0 = (double)(ip - 1) * aa, x1 = (double)ip * aa;
      const double y0 = (double)(iq - 1) * bb, y1 = (double)iq * bb;
      const double z0 = -((double)is * cc), z1 = -((double)(is - 1) * cc);   /* -is*cc = -(is*cc): -0.0 on the top plane */
      x[1] = x[2] = x[5] = x[6] = x0; x[3] = x[4] = x[7] = x[8] = x1;
      y[1] = y[2] = y[3] = y[4] = y0; y[5] = y[6] = y[7] = y[8] = y1;
      z[1] = z[4] = z[5] = z[8] = z0; z[2] = z[3] = z[6] = z[7] = z1;
    } else {
      const int64_t plane = (int64_t)(2 * nxe + 1) * (nze + 1) + (int64_t)(2 * nze + 1) * (nxe + 1);
      const int64_t fac1 = plane * (iq - 1), fac2 = plane * iq;
      num[1] = fac1 + (3 * nxe + 2) * is + 2 * ip - 1;
      num[2] = fac1 + (3 * nxe + 2) * is - nxe + ip - 1;
      num[3] = num[1] - 3 * nxe - 2; num[4] = num[3] + 1; num[5] = num[4] + 1; num[6] = num[2] + 1;
      num[7] = num[1] + 2; num[8] = num[1] + 1;
      num[9] = fac2 - (int64_t)(nxe + 1) * (nze + 1) + (nxe + 1) * is + ip;
      num[10] = num[9] - nxe - 1; num[11] = num[10] + 1; num[12] = num[9] + 1;
      num[13] = fac2 + (3 * nxe + 2) * is + 2 * ip - 1;
      num[14] = fac2 + (3 * nxe + 2) * is - nxe + ip - 1;
      num[15] = num[13] - 3 * nxe - 2; num[16] = num[15] + 1; num[17] = num[16] + 1; num[18] = num[14] + 1;
      num[19] = num[13] + 2; num[20] = num[13] + 1;
      const double x0 = (double)(ip - 1) * aa, x1 = (double)ip * aa;
      x[1] = x[2] = x[3] = x[9] = x[10] = x[13] = x[14] = x[15] = x0;
      x[5] = x[6] = x[7] = x[11] = x[12] = x[17] = x[18] = x[19] = x1;
      x[4] = .5 * (x[3] + x[5]); x[8] = .5 * (x[1] + x[7]); x[16] = .5 * (x[15] + x[17]); x[20] = .5 * (x[13] + x[19]);
      const double y0 = (double)(iq - 1) * bb, y1 = (double)iq * bb;
      for (int m = 1; m <= 8; ++m) y[m] = y0;
      for (int m = 13; m <= 20; ++m) y[m] = y1;
      y[9] = .5 * (y[1] + y[13]); y[10] = .5 * (y[3] + y[15]); y[11] = .5 * (y[5] + y[17]); y[12] = .5 * (y[7] + y[19]);
      const double z0 = -((double)is * cc), z1 = -((double)(is - 1) * cc);   /* -is*cc = -(is*cc): -0.0 on the top plane */
      z[1] = z[7] = z[8] = z[9] = z[12] = z[13] = z[19] = z[20] = z0;
      z[3] = z[4] = z[5] = z[10] = z[11] = z[15] = z[16] = z[17] = z1;
      z[2] = .5 * (z[1] + z[3]); z[6] = .5 * (z[5] + z[7]); z[14] = .5 * (z[13] + z[15]); z[18] = .5 * (z[17] + z[19]);
    }
    for (int m = 1; m <= nod; ++m) {
      g_num[q * nod + m - 1] = (int32_t)num[m];
      g_coord_pp[q * 3 * nod + m - 1] = x[m];
      g_coord_pp[q * 3 * nod + nod + m - 1] = y[m];
      g_coord_pp[q * 3 * nod + 2 * nod + m - 1] = z[m];
    }
  }
  return 0;
}

/* cube_bc20 (geometry.f90:445-498), cube_bc8 (:609-647), box_bc8 (:664-694): rest(nr,nodof+1) col-major,
 * kind 20 / 8 (nodof 3) or 1 (box_bc8, nodof 1).  Returns the number of rows written (= nr of p12meshgen). */
int64_t orc_cube_rest(int kind, int nxe, int nye, int nze, int64_t nr, int32_t *rest) {
  const int nodof = kind == 1 ? 1 : 3;
  int64_t count = 0;
#define PUT(node, a, b, c) do { if (count < nr) { rest[count] = (int32_t)(node); rest[nr + count] = (a); \
      if (nodof == 3) { rest[2 * nr + count] = (b); rest[3 * nr + count] = (c); } } ++count; } while (0)
  if (kind == 1) {
    const int64_t face = (int64_t)(nxe + 1) * (nze + 1);
    for (int64_t i = 0; i <= nye - 1; ++i)
      for (int64_t j = i * face + 1; j <= (i + 1) * face; ++j)
        if (j / (nxe + 1) * (nxe + 1) == j || j < i * face + nxe + 1) PUT(j, 0, 0, 0);
    for (int64_t j = nye * face + 1; j <= (nye + 1) * face; ++j) PUT(j, 0, 0, 0);
    return count;
  }
  int64_t face1, face2, l, m, n;
  if (kind == 20) {
    face1 = 3LL * nxe * nze + 2 * (nxe + nze) + 1; face2 = (int64_t)(nxe + 1) * (nze + 1);
    l = (int64_t)nze * (nxe + 1); m = 3 * nxe + 2; n = 3LL * nxe * nze + 2 * nze;
  } else {
    face1 = (int64_t)(nxe + 1) * (nze + 1); face2 = 0;
    l = 0; m = 2 * nxe + 2; n = (int64_t)(nxe + 1) * nze;
  }
  const int64_t face = face1 + face2;
  for (int64_t i = 0; i <= nye; ++i) {
    for (int64_t j = i * face + 1; j <= i * face + face1; ++j) {
      const int64_t k = j - i * face;
      const int edge = k <= n && ((k + m - 1) / m * m == (k + m - 1) || (k + nxe + 1) / m * m == (k + nxe + 1) ||
                                  (k + nxe) / m * m == (k + nxe) || k / m * m == k);
      if (i == 0 || i == nye) {
        if (edge) PUT(j, 0, 0, 1);
        else if (k <= n) PUT(j, 1, 0, 1);
        else PUT(j, 0, 0, 0);
      } else {
        if (edge) PUT(j, 0, 1, 1);
        else if (k > n) PUT(j, 0, 0, 0);
      }
    }
    if (kind == 20 && i < nye)
      for (int64_t j = i * face + face1 + 1; j <= (i + 1) * face; ++j) {
        const int64_t k = j - (face1 + i * face);
        if (k <= l && ((k + nxe) / (nxe + 1) * (nxe + 1) == (k + nxe) || k / (nxe + 1) * (nxe + 1) == k)) PUT(j, 0, 1, 1);
        else if (k > l) PUT(j, 0, 0, 0);
      }
  }
#undef PUT
  return count;
}

/* load_p121 (loading.f90:386-546) followed by p12meshgen's scaling (p12meshgen.f90:167-168, 210-211):
 * node[loaded], val[loaded] = the z-load of every loaded node; returns loaded (call with node == NULL for the count) */
int64_t orc_load_p121(int nod, int nxe, int nze, double aa, double bb, int32_t *node, double *val) {
  const int nle = nxe / 5;
  const int64_t loaded = nod == 20 ? 3LL * nle * nle + 4 * nle + 1 : (int64_t)(nle + 1) * (nle + 1);
  if (!node) return loaded;
  int64_t c = 0;
  if (nod == 20) {
    const int64_t f1 = (int64_t)(2 * nxe + 1) * (nze + 1) + (int64_t)(nxe + 1) * nze, f2 = (int64_t)(nxe + 1) * (nze + 1);
    for (int i = 1; i <= 2 * nle + 1; ++i, ++c) {
      node[c] = i;
      val[c] = (i == 1 || i == 2 * nle + 1) ? -1.0 : (i % 2 == 0 ? 4.0 : -2.0);
    }
    for (int j = 0; j <= nle - 1; ++j) {
      for (int i = 1; i <= nle + 1; ++i, ++c) {
        node[c] = (int32_t)(i + f1 + j * (f1 + f2));
        val[c] = (i == 1 || i == nle + 1) ? 4.0 : 8.0;
      }
      for (int i = 1; i <= 2 * nle + 1; ++i, ++c) {
        node[c] = (int32_t)(i + f1 + (j + 1) * f2 + j * f1);
        if (j != nle - 1) val[c] = (i == 1 || i == 2 * nle + 1) ? -2.0 : (i % 2 == 0 ? 8.0 : -4.0);
        else val[c] = (i == 1 || i == 2 * nle + 1) ? -1.0 : (i % 2 == 0 ? 4.0 : -2.0);
      }
    }
    for (int64_t i = 0; i < c; ++i) val[i] = -val[i] * aa * bb * (25.0 / 12.0);
  } else {
    const int64_t f1 = (int64_t)(nxe + 1) * (nze + 1);
    for (int i = 1; i <= nle + 1; ++i, ++c) {
      node[c] = i;
      val[c] = (i == 1 || i == nle + 1) ? -6.25 : -12.5;
    }
    for (int j = 0; j <= nle - 1; ++j)
      for (int i = 1; i <= nle + 1; ++i, ++c) {
        node[c] = (int32_t)(i + (j + 1) * f1);
        if (j != nle - 1) val[c] = (i == 1 || i == nle + 1) ? -12.5 : -25.0;
        else val[c] = (i == 1 || i == nle + 1) ? -6.25 : -12.5;
      }
    for (int64_t i = 0; i < c; ++i) val[i] = val[i] * aa * bb;
  }
  return c == loaded ? loaded : -1;
}

/* ------------------------------------------------------------------------- */
/* partition (calc_nels_pp / calc_neq_pp, gather_scatter.f90:217-238,319-339) */
/* ------------------------------------------------------------------------- */
void orc_partition(int64_t n, int npes, int numpe /*1-based*/, int64_t *cnt, int64_t *start /*1-based*/) {
  if (npes == 1) { *cnt = n; *start = 1; return; }
  int64_t pp2 = n / npes, num1 = n - pp2 * npes, pp1 = num1 == 0 ? pp2 : pp2 + 1;
  if (numpe <= num1 || num1 == 0) { *cnt = pp1; *start = (int64_t)(numpe - 1) * pp1 + 1; }
  else { *cnt = pp2; *start = num1 * pp1 + (numpe - num1 - 1) * (pp1 - 1) + 1; }
}

/* ------------------------------------------------------------------------- */
/* per-iteration kernels                                                      */
/* ------------------------------------------------------------------------- */

/* gather: pmul(:,e) = p(g(:,e)), restrained (0) -> 0  (gather_scatter.f90:663-665) */
void orc_gather(int ntot, int64_t nels, const int32_t *g_g, const double *p /*1-based eq -> p[eq-1]*/,
                double *pmul) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nels * ntot; ++i) pmul[i] = g_g[i] > 0 ? p[g_g[i] - 1] : 0.0;
}

/* elements_3: utemp(:,e) = MATMUL(km(:,:,e), pmul(:,e))  (p121.f90:93-97) */
void orc_matvec(int ntot, int64_t nels, const double *storkm, const double *pmul, double *utemp) {
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < nels; ++e) {
    const double *km = storkm + e * ntot * ntot, *pm = pmul + e * ntot;
    double *ut = utemp + e * ntot;
    for (int i = 0; i < ntot; ++i) ut[i] = 0.0;
    for (int j = 0; j < ntot; ++j) {
      const double pj = pm[j];
      for (int i = 0; i < ntot; ++i) ut[i] = ut[i] + km[j * ntot + i] * pj;
    }
  }
}


/* ------------------------------------------------------------------------- */
/* matrix-free variant (BASELINE config E): utemp(:,e) = sum_gp B^T (D (B p)) det w   */
/* ------------------------------------------------------------------------- */
/*
 * Operator form of elements_3 without storkm.  With deriv = jac^-1 * der the element product
 * factors through two 3x3 matrices per Gauss point (ascending):
 *   jac, det, jac^-1 as in gauss_point() but with fused multiply-adds (fma chains from 0.0); order 2: the inverse is the
 *   adjugate times ONE reciprocal of det instead of nine divisions
 *   H(b,c) = sum_m der(b,m) p_c(m)          node-ascending fma chains from 0.0
 *   G(a,c) = sum_b inv(a,b) H(b,c)          product, then two fmas, b ascending
 *   eps    = (G00, G11, G22, G10+G01, G21+G12, G20+G02)      beemat's row order
 *   sigma  = D eps                          product + c-ascending fmas per row, then * det*w
 *   T(b,c) = sum_a inv(a,b) S(a,c)          S = symmetric stress tensor; product, two fmas
 * and then per dof the sum over (Gauss point, b) in the order of the device kernel in use (g_mf_order below):
 *   order 2 (k_apply_mf3)  u_c(m) = one fma chain from 0.0 over (h | b | q), Gauss point 2q+h
 *   order 1 (k_apply_mf2)  u_c(m) = [sum_{gp<4} sum_b der_gp(b,m) T_gp(b,c)] + [sum_{gp>=4} ...], each: a product, then 11 fmas
 *   order 0 (k_apply_mf)   ONE 24-term chain, Gauss point ascending, b ascending
 * These are the operation orders of the kernels in parafem_b200/csrc/kernels.cuh.  Each is a
 * different rounding of the same operator as MATMUL(storkm,pmul) (p121.f90:94), hence its
 * own oracle.
 */
/* 2: the order of k_apply_mf3 (FP64 tensor-core kernel, the default for both bricks): per freedom ONE 24-term fma chain
 *    from 0.0 over (h = 0,1 | b = 0,1,2 | q = 0..3) with Gauss point 2q+h -- the k order of its mma.sync.m8n8k4.f64
 *    instructions, each of which is a k-ascending fma chain on the accumulator (measured on the B200 bit for bit,
 *    scripts/probe/dmma_probe.cu);
 * 1: the order of k_apply_mf2 (two lanes per element, PF_MF=2lane); 0: the order of the round-1 kernel k_apply_mf
 *    (PF_MF=1lane) */
static int g_mf_order = 2;
void orc_set_mf_order(int order) { g_mf_order = (order >= 0 && order <= 2) ? order : 2; }
/* orc_invert3 with the nine divisions by det replaced by products with 1/det (the tensor-core kernels' inv3_recip) */
static void invert3_recip(double *m) {
  const double rd = 1.0 / orc_determinant3(m);
  double j11 = M3(m, 2, 2) * M3(m, 3, 3) - M3(m, 3, 2) * M3(m, 2, 3);
  double j21 = -(M3(m, 2, 1) * M3(m, 3, 3)) + M3(m, 3, 1) * M3(m, 2, 3);
  double j31 = M3(m, 2, 1) * M3(m, 3, 2) - M3(m, 3, 1) * M3(m, 2, 2);
  double j12 = -(M3(m, 1, 2) * M3(m, 3, 3)) + M3(m, 3, 2) * M3(m, 1, 3);
  double j22 = M3(m, 1, 1) * M3(m, 3, 3) - M3(m, 3, 1) * M3(m, 1, 3);
  double j32 = -(M3(m, 1, 1) * M3(m, 3, 2)) + M3(m, 3, 1) * M3(m, 1, 2);
  double j13 = M3(m, 1, 2) * M3(m, 2, 3) - M3(m, 2, 2) * M3(m, 1, 3);
  double j23 = -(M3(m, 1, 1) * M3(m, 2, 3)) + M3(m, 2, 1) * M3(m, 1, 3);
  double j33 = M3(m, 1, 1) * M3(m, 2, 2) - M3(m, 2, 1) * M3(m, 1, 2);
  M3(m, 1, 1) = j11 * rd; M3(m, 1, 2) = j12 * rd; M3(m, 1, 3) = j13 * rd;
  M3(m, 2, 1) = j21 * rd; M3(m, 2, 2) = j22 * rd; M3(m, 2, 3) = j23 * rd;
  M3(m, 3, 1) = j31 * rd; M3(m, 3, 2) = j32 * rd; M3(m, 3, 3) = j33 * rd;
}

static void mf_element(int nod, double der[8][60] /*[ig][a*20+m]*/, const double *coord, const double *dee,
                       const double *weights, const double *pm, double *ut) {
  double T[8][9];
  for (int ig = 0; ig < 8; ++ig) {
    double jac[9], inv[9], H[9], G[9], sig[6];
    for (int b = 0; b < 3; ++b)
      for (int a = 0; a < 3; ++a) {
        double s = 0.0;
        for (int m = 0; m < nod; ++m) s = fma(der[ig][a * 20 + m], coord[b * nod + m], s);
        jac[b * 3 + a] = s;
      }
    const double det = orc_determinant3(jac);
    memcpy(inv, jac, sizeof inv);
    if (g_mf_order == 2) invert3_recip(inv);   /* k_apply_mf3 / k_apply_mf4: adjugate times ONE reciprocal (inv3_recip) */
    else orc_invert3(inv);
    const double f = det * weights[ig];
    for (int b = 0; b < 3; ++b)
      for (int c = 0; c < 3; ++c) {
        double s = 0.0;
        for (int m = 0; m < nod; ++m) s = fma(der[ig][b * 20 + m], pm[3 * m + c], s);
        H[b * 3 + c] = s;
      }
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) {
        double s = inv[a] * H[c];
        s = fma(inv[3 + a], H[3 + c], s);
        s = fma(inv[6 + a], H[6 + c], s);
        G[a * 3 + c] = s;
      }
    const double eps[6] = {G[0], G[4], G[8], G[3] + G[1], G[7] + G[5], G[6] + G[2]};
    if (g_mf_order == 2) {              /* k_apply_mf3 leaves deemat's structural zeros out (mf_point_mid_iso) */
      for (int r = 0; r < 3; ++r) {
        double s = dee[r] * eps[0];
        s = fma(dee[6 + r], eps[1], s);
        s = fma(dee[12 + r], eps[2], s);
        sig[r] = s * f;
      }
      for (int r = 3; r < 6; ++r) sig[r] = (dee[r * 6 + r] * eps[r]) * f;
    } else
    for (int r = 0; r < 6; ++r) {
      double s = dee[r] * eps[0];
      for (int c = 1; c < 6; ++c) s = fma(dee[c * 6 + r], eps[c], s);
      sig[r] = s * f;
    }
    const double S[9] = {sig[0], sig[3], sig[5], sig[3], sig[1], sig[4], sig[5], sig[4], sig[2]};
    for (int b = 0; b < 3; ++b)
      for (int c = 0; c < 3; ++c) {
        double s = inv[b * 3] * S[c];
        s = fma(inv[b * 3 + 1], S[3 + c], s);
        s = fma(inv[b * 3 + 2], S[6 + c], s);
        T[ig][b * 3 + c] = s;
      }
  }
  for (int m = 0; m < nod; ++m)
    for (int c = 0; c < 3; ++c) {
      if (g_mf_order == 2) {            /* k_apply_mf3: the k order of its phase-3 mma instructions */
        double s = 0.0;
        for (int h = 0; h < 2; ++h)
          for (int b = 0; b < 3; ++b)
            for (int q = 0; q < 4; ++q) s = fma(T[2 * q + h][b * 3 + c], der[2 * q + h][b * 20 + m], s);
        ut[3 * m + c] = s;
      } else if (g_mf_order == 0) {     /* k_apply_mf (round 1): one 24-term chain */
        double s = der[0][m] * T[0][c];
        for (int ig = 0; ig < 8; ++ig)
          for (int b = 0; b < 3; ++b) {
            if (ig == 0 && b == 0) continue;
            s = fma(der[ig][b * 20 + m], T[ig][b * 3 + c], s);
          }
        ut[3 * m + c] = s;
      } else {                          /* k_apply_mf2: a 12-term chain per half of the points, then one add */
        double part[2];
        for (int hf = 0; hf < 2; ++hf) {
          double s = der[4 * hf][m] * T[4 * hf][c];
          for (int ig = 4 * hf; ig < 4 * hf + 4; ++ig)
            for (int b = 0; b < 3; ++b) {
              if (ig == 4 * hf && b == 0) continue;
              s = fma(der[ig][b * 20 + m], T[ig][b * 3 + c], s);
            }
          part[hf] = s;
        }
        ut[3 * m + c] = part[0] + part[1];
      }
    }
}

int orc_apply_mf(int64_t nels, int nod, int nip, const double *g_coord_pp, double e, double v,
                 const double *pmul, double *utemp) {
  if ((nod != 8 && nod != 20) || nip != 8) return 1;
  const int ntot = 3 * nod;
  double points[81], weights[27], dee[36], der[8][60], d3[60];
  orc_sample_hex(nip, points, weights);
  orc_deemat6(dee, e, v);
  for (int ig = 0; ig < nip; ++ig) {
    orc_shape_der(nod, points, nip, ig, d3);
    memset(der[ig], 0, sizeof der[ig]);
    for (int m = 0; m < nod; ++m)
      for (int a = 0; a < 3; ++a) der[ig][a * 20 + m] = d3[m * 3 + a];
  }
#pragma omp parallel for schedule(static)
  for (int64_t iel = 0; iel < nels; ++iel)
    mf_element(nod, der, g_coord_pp + iel * nod * 3, dee, weights, pmul + iel * ntot, utemp + iel * ntot);
  return 0;
}

typedef struct {
  int npes;
  int64_t *el0, *el1;   /* element range [el0,el1) per rank (0-based) */
  int64_t *eq0, *eq1;   /* owned equation range [eq0,eq1) per rank (0-based eq index) */
  int64_t *lo, *hi;     /* bounding range of equations touched by the rank's elements */
  double **ul;          /* per-rank partial-sum buffers over [lo,hi) */
} orc_ranks;

/* partitioner 2 (read_nels_pp, input.f90:3108-3196): elements per rank from <job>.psize instead of
 * calc_nels_pp.  orc_set_element_partition(counts, npes) makes every emulated-rank routine below use
 * those counts (npes must match; counts == NULL restores partitioner 1). */
static int64_t g_psize[64];
static int g_psize_n = 0;
void orc_set_element_partition(const int64_t *counts, int npes) {
  g_psize_n = 0;
  if (!counts || npes < 1 || npes > 64) return;
  for (int r = 0; r < npes; ++r) g_psize[r] = counts[r];
  g_psize_n = npes;
}
static void element_range(int64_t nels, int npes, int r, int64_t *e0, int64_t *e1) {
  if (g_psize_n == npes) {
    int64_t s = 0;
    for (int q = 0; q < r; ++q) s += g_psize[q];
    *e0 = s; *e1 = s + g_psize[r];
    return;
  }
  int64_t c, s;
  orc_partition(nels, npes, r + 1, &c, &s);
  *e0 = s - 1; *e1 = s - 1 + c;
}

static orc_ranks *ranks_new(int npes, int ntot, int64_t nels, const int32_t *g_g, int64_t neq) {
  orc_ranks *R = calloc(1, sizeof *R);
  R->npes = npes;
  R->el0 = malloc(sizeof(int64_t) * npes); R->el1 = malloc(sizeof(int64_t) * npes);
  R->eq0 = malloc(sizeof(int64_t) * npes); R->eq1 = malloc(sizeof(int64_t) * npes);
  R->lo = malloc(sizeof(int64_t) * npes);  R->hi = malloc(sizeof(int64_t) * npes);
  R->ul = calloc(npes, sizeof(double *));
  for (int r = 0; r < npes; ++r) {
    int64_t c, s;
    element_range(nels, npes, r, &R->el0[r], &R->el1[r]);
    orc_partition(neq, npes, r + 1, &c, &s);  R->eq0[r] = s - 1; R->eq1[r] = s - 1 + c;
    int64_t lo = neq, hi = 0;
    for (int64_t i = R->el0[r] * ntot; i < R->el1[r] * ntot; ++i)
      if (g_g[i] > 0) { if (g_g[i] - 1 < lo) lo = g_g[i] - 1; if (g_g[i] > hi) hi = g_g[i]; }
    if (hi < lo) { lo = 0; hi = 0; }
    R->lo[r] = lo; R->hi[r] = hi;
    R->ul[r] = malloc(sizeof(double) * (size_t)(hi - lo + 1));
  }
  return R;
}

static void ranks_free(orc_ranks *R) {
  for (int r = 0; r < R->npes; ++r) free(R->ul[r]);
  free(R->ul); free(R->el0); free(R->el1); free(R->eq0); free(R->eq1); free(R->lo); free(R->hi); free(R);
}

/* scatter(u,utemp) over emulated ranks, u zeroed first (p121.f90:91) */
static void ranks_scatter(orc_ranks *R, int ntot, const int32_t *g_g, const double *utemp, double *u) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int r = 0; r < R->npes; ++r) {
    double *ul = R->ul[r];
    const int64_t lo = R->lo[r];
    memset(ul, 0, sizeof(double) * (size_t)(R->hi[r] - lo));
    for (int64_t i = R->el0[r] * ntot; i < R->el1[r] * ntot; ++i)
      if (g_g[i] > 0) ul[g_g[i] - 1 - lo] = ul[g_g[i] - 1 - lo] + utemp[i];
  }
#pragma omp parallel for schedule(dynamic, 1)
  for (int o = 0; o < R->npes; ++o) {
    for (int64_t q = R->eq0[o]; q < R->eq1[o]; ++q)
      u[q] = (q >= R->lo[o] && q < R->hi[o]) ? 0.0 + R->ul[o][q - R->lo[o]] : 0.0;
    for (int r = 0; r < R->npes; ++r) {
      if (r == o) continue;
      int64_t a = R->lo[r] > R->eq0[o] ? R->lo[r] : R->eq0[o];
      int64_t b = R->hi[r] < R->eq1[o] ? R->hi[r] : R->eq1[o];
      for (int64_t q = a; q < b; ++q) u[q] = u[q] + R->ul[r][q - R->lo[r]];
    }
  }
}

/* public single-call scatter for kernel-level parity tests */
int orc_scatter(int ntot, int64_t nels, const int32_t *g_g, int64_t neq, int npes, const double *utemp,
                double *u) {
  orc_ranks *R = ranks_new(npes, ntot, nels, g_g, neq);
  ranks_scatter(R, ntot, g_g, utemp, u);
  ranks_free(R);
  return 0;
}

/* u = scatter(MATMUL(storkm, gather(x))) over npes emulated ranks: the operator product the
 * drivers form outside the PCG loop (p124.f90:150-154 with storkb, :175-178 with storka) */
int orc_apply(int ntot, int64_t nels, const int32_t *g_g, const double *storkm, int64_t neq, int npes,
              const double *x, double *u) {
  double *pmul = malloc(sizeof(double) * (size_t)(nels * ntot));
  double *utemp = malloc(sizeof(double) * (size_t)(nels * ntot));
  orc_gather(ntot, nels, g_g, x, pmul);
  orc_matvec(ntot, nels, storkm, pmul, utemp);
  orc_scatter(ntot, nels, g_g, neq, npes, utemp, u);
  free(pmul); free(utemp);
  return 0;
}

/*
 * The fixed blocked reduction tree (red_mode 1).  For a local vector of n
 * entries:
 *   chunk c   = entries [2048c, 2048c+2048)
 *   lane sum  : thread t (0..255) adds, in this order, the products at
 *               2048c + 512k + 2t and 2048c + 512k + 2t + 1 for k = 0..3,
 *               skipping indices >= n, starting from 0.0
 *   warp tree : for off = 16,8,4,2,1: v[i] = v[i] + v[i ^ off] within each
 *               group of 32 lanes
 *   block sum : s = w0; s = s + w1; ... + w7 over the 8 warps
 *   final     : thread t adds the chunk sums t, t+256, t+512, ... in order
 *               from 0.0, then the same warp tree and block sum
 * This is exactly what parafem_b200/csrc/kernels.cu's reductions compute.
 */
static double block_tree(double *v /*256*/) {
  for (int off = 16; off >= 1; off >>= 1) {
    double t[256];
    for (int i = 0; i < 256; ++i) t[i] = v[i] + v[i ^ off];
    memcpy(v, t, sizeof t);
  }
  double s = v[0];
  for (int w = 1; w < 8; ++w) s = s + v[32 * w];
  return s;
}

double orc_dot_blocked(const double *a, const double *b, int64_t n) {
  const int64_t nchunks = (n + 2047) / 2048;
  double *part = malloc(sizeof(double) * (size_t)(nchunks > 0 ? nchunks : 1));
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < nchunks; ++c) {
    double v[256];
    for (int t = 0; t < 256; ++t) {
      double acc = 0.0;
      for (int k = 0; k < 4; ++k)
        for (int h = 0; h < 2; ++h) {
          int64_t i = 2048 * c + 512 * k + 2 * t + h;
          if (i < n) acc = acc + a[i] * b[i];
        }
      v[t] = acc;
    }
    part[c] = block_tree(v);
  }
  double v[256];
  for (int t = 0; t < 256; ++t) {
    double acc = 0.0;
    for (int64_t c = t; c < nchunks; c += 256) acc = acc + part[c];
    v[t] = acc;
  }
  free(part);
  return block_tree(v);
}

static double dot_seq(const double *a, const double *b, int64_t n) {
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) s = s + a[i] * b[i];
  return s;
}

/* dot_product_p over emulated ranks: local dot, then ranks ascending */
double orc_dot_ranks(const double *a, const double *b, int64_t neq, int npes, int red_mode) {
  double s = 0.0;
  if (!red_mode && npes > 1) {
    /* every emulated rank forms its local DOT_PRODUCT at the same time (as the MPI ranks do); the partials
     * are then added in rank order: the same bits as the serial loop below */
    double *part = malloc(sizeof(double) * (size_t)npes);
#pragma omp parallel for schedule(static, 1)
    for (int r = 0; r < npes; ++r) {
      int64_t c, st;
      orc_partition(neq, npes, r + 1, &c, &st);
      part[r] = dot_seq(a + st - 1, b + st - 1, c);
    }
    for (int r = 0; r < npes; ++r) s = (r == 0) ? part[r] : s + part[r];
    free(part);
    return s;
  }
  for (int r = 0; r < npes; ++r) {
    int64_t c, st;
    orc_partition(neq, npes, r + 1, &c, &st);
    double part = red_mode ? orc_dot_blocked(a + st - 1, b + st - 1, c) : dot_seq(a + st - 1, b + st - 1, c);
    s = (r == 0) ? part : s + part;
  }
  return s;
}

/* ------------------------------------------------------------------------- */
/* the solver: p121.f90:65-69,86-104 / p123.f90:86-92,120-151                 */
/* ------------------------------------------------------------------------- */
/*
 * storkm(ntot,ntot,nels), g_g(ntot,nels) global; r(neq) the starting residual
 * (loads); fixed equations (p123): no_f[nfixed] global 1-based, val_f values,
 * penalty; npes emulated ranks (also the OpenMP width of the element loops).
 * Outputs: x(neq), *iters, *converged, ratio[limit] (may be NULL),
 * diag_out(neq) (may be NULL) = inverted preconditioner.
 * mf_coord != NULL: the element products use orc_apply_mf (matrix-free variant) instead of
 * storkm, which is then only read for the preconditioner diagonal.
 * Returns seconds spent in the iteration loop (the reference's timest(3)
 * window, p121.f90:89,107-108) through *solve_seconds.
 */
int orc_pcg(int ntot, int64_t nels, const int32_t *g_g, const double *storkm, int64_t neq,
            const double *r_in, int64_t nfixed, const int32_t *no_f, const double *val_f,
            double penalty, int npes, int red_mode, double tol, int limit, double *x_out,
            int *iters_out, int *converged_out, double *ratio, double *diag_out,
            double *solve_seconds, const double *mf_coord, int mf_nod, int mf_nip, double mf_e, double mf_v) {
  orc_ranks *R = ranks_new(npes, ntot, nels, g_g, neq);
  double *pmul = malloc(sizeof(double) * (size_t)(nels * ntot));
  double *utemp = malloc(sizeof(double) * (size_t)(nels * ntot));
  double *diag = calloc((size_t)neq, sizeof(double)), *p = calloc((size_t)neq, sizeof(double));
  double *r = malloc(sizeof(double) * (size_t)neq), *x = calloc((size_t)neq, sizeof(double));
  double *xnew = calloc((size_t)neq, sizeof(double)), *u = calloc((size_t)neq, sizeof(double));
  double *d = calloc((size_t)neq, sizeof(double));
  double *store = nfixed > 0 ? malloc(sizeof(double) * (size_t)nfixed) : NULL;
  memcpy(r, r_in, sizeof(double) * (size_t)neq);

  /* diag_precon_tmp(i,iel) = storkm(i,i,iel); scatter (p121.f90:65-69) */
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < nels; ++e)
    for (int i = 0; i < ntot; ++i) utemp[e * ntot + i] = storkm[e * ntot * ntot + i * ntot + i];
  ranks_scatter(R, ntot, g_g, utemp, diag);
  for (int64_t i = 0; i < nfixed; ++i) {      /* p123.f90:120-125 */
    int64_t j = no_f[i] - 1;
    diag[j] = diag[j] + penalty; store[i] = diag[j];
  }
  for (int64_t i = 0; i < neq; ++i) diag[i] = 1.0 / diag[i];
  if (val_f)   /* p123.f90:127-131; p124 passes its own fixed-freedom residual (val_f == NULL) */
    for (int64_t i = 0; i < nfixed; ++i) r[no_f[i] - 1] = store[i] * val_f[i];
  for (int64_t i = 0; i < neq; ++i) { d[i] = diag[i] * r[i]; p[i] = d[i]; }
  if (diag_out) memcpy(diag_out, diag, sizeof(double) * (size_t)neq);

  int iters = 0, converged = 0;
  double t0 = 0.0;
#ifdef _OPENMP
  t0 = omp_get_wtime();
#endif
  for (;;) {
    iters = iters + 1;
    orc_gather(ntot, nels, g_g, p, pmul);
    if (mf_coord) orc_apply_mf(nels, mf_nod, mf_nip, mf_coord, mf_e, mf_v, pmul, utemp);
    else orc_matvec(ntot, nels, storkm, pmul, utemp);
    ranks_scatter(R, ntot, g_g, utemp, u);
    for (int64_t i = 0; i < nfixed; ++i) u[no_f[i] - 1] = p[no_f[i] - 1] * store[i];  /* p123.f90:141-145 */
    const double up = orc_dot_ranks(r, d, neq, npes, red_mode);
    const double alpha = up / orc_dot_ranks(p, u, neq, npes, red_mode);
    double maxloads = 0.0, maxdiff = 0.0;
#pragma omp parallel for schedule(static) reduction(max : maxloads, maxdiff)
    for (int64_t i = 0; i < neq; ++i) {
      xnew[i] = x[i] + p[i] * alpha;
      r[i] = r[i] - u[i] * alpha;
      d[i] = diag[i] * r[i];
      const double al = fabs(xnew[i]), ad = fabs(xnew[i] - x[i]);
      if (al > maxloads) maxloads = al;
      if (ad > maxdiff) maxdiff = ad;
    }
    const double beta = orc_dot_ranks(r, d, neq, npes, red_mode) / up;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < neq; ++i) { p[i] = d[i] + p[i] * beta; x[i] = xnew[i]; }
    /* checon_par, maths.f90:1052-1061 */
    const double rat = maxdiff / maxloads;
    if (ratio) ratio[iters - 1] = rat;
    converged = rat <= tol;
    if (converged || iters == limit) break;
  }
#ifdef _OPENMP
  if (solve_seconds) *solve_seconds = omp_get_wtime() - t0;
#else
  if (solve_seconds) *solve_seconds = 0.0;
#endif
  memcpy(x_out, xnew, sizeof(double) * (size_t)neq);
  *iters_out = iters; *converged_out = converged;
  free(pmul); free(utemp); free(diag); free(p); free(r); free(x); free(xnew); free(u); free(d); free(store);
  ranks_free(R);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* p1210 (programs/5th_ed/p1210/p1210.f90): forced vibration of an elastic-plastic (von Mises) solid,   */
/* lumped mass, explicit integration.  20-node bricks (the lumping pattern of :100-103 is theirs).      */
/* ------------------------------------------------------------------------- */
/* elements_1 (p1210.f90:93-104): mm_tmp(:,iel) = emm, emm = volume/13 on the mid-side freedoms, an eighth of that on
 * the corner freedoms (1:19:6, 2:20:6, 3:21:6, 37:55:6, 38:56:6, 39:57:6), volume = sum_gp det*w*rho */
int orc_p1210_mass(int64_t nels, int nod, int nip, const double *g_coord_pp, double rho, double *mm_tmp) {
  if (nod != 20 || nip != 8) return 1;
  double points[81], weights[27];
  orc_sample_hex(nip, points, weights);
  for (int64_t iel = 0; iel < nels; ++iel) {
    double der[60], deriv[60], volume = 0.0;
    for (int ig = 0; ig < nip; ++ig) {
      const double det = gauss_point(nod, points, nip, ig, g_coord_pp + iel * nod * 3, der, deriv);
      volume = volume + det * weights[ig] * rho;
    }
    double *emm = mm_tmp + iel * 60;
    const double mid = volume / 13.0;
    for (int q = 0; q < 60; ++q) emm[q] = mid;
    for (int c = 0; c < 3; ++c)
      for (int q = c; q <= 18 + c; q += 6) { emm[q] = mid * .125; emm[36 + q] = mid * .125; }
  }
  return 0;
}

/* vmpl, nst = 6 (new_library.f90:2262-2284): pl(i,j) = term(i)*term(j)*ee */
static void vmpl6(double e, double v, const double *s, double *pl) {
  const double sx = s[0], sy = s[1], sz = s[2], txy = s[3], tyz = s[4], tzx = s[5];
  const double dsbar = sqrt((sx - sy) * (sx - sy) + (sy - sz) * (sy - sz) + (sz - sx) * (sz - sx) + 6.0 * (txy * txy) +
                            6.0 * (tyz * tyz) + 6.0 * (tzx * tzx)) / sqrt(2.0);
  const double ee = 1.5 * e / ((1.0 + v) * dsbar * dsbar);
  const double term[6] = {(2.0 * sx - sy - sz) / 3.0, (2.0 * sy - sz - sx) / 3.0, (2.0 * sz - sx - sy) / 3.0, txy, tyz, tzx};
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) pl[j * 6 + i] = term[i] * term[j] * ee;
}
/* the second invariant of invar (new_library.f90:1891-1897): dsbar = sqrt(3)*sqrt(d2); p1210 uses nothing else of it */
static double dsbar6(const double *s) {
  const double d2 = ((s[0] - s[1]) * (s[0] - s[1]) + (s[1] - s[2]) * (s[1] - s[2]) + (s[2] - s[0]) * (s[2] - s[0])) / 6.0 +
                    s[3] * s[3] + s[4] * s[4] + s[5] * s[5];
  return sqrt(3.0) * sqrt(d2);
}

/* elements_2 (p1210.f90:120-147) over all elements: pmul = gathered displacements, etensor / tensor (6,nip,nels) updated
 * in place, utemp = -bload.  Every MATMUL: k ascending from 0.0, separate multiply and add. */
int orc_p1210_elements(int64_t nels, int nod, int nip, const double *g_coord_pp, double e, double v, double sbary,
                       const double *pmul, double *etensor, double *tensor, double *utemp) {
  if (nod != 20 || nip != 8) return 1;
  const int ntot = 3 * nod;
  double points[81], weights[27], dee0[36];
  orc_sample_hex(nip, points, weights);
  orc_deemat6(dee0, e, v);
#pragma omp parallel for schedule(static) if (nels > 256)
  for (int64_t iel = 0; iel < nels; ++iel) {
    double der[60], deriv[60], bee[6 * 60], bload[60], eps[6], sigma[6], stressv[6], dee[36], pl[36];
    const double *eld = pmul + iel * ntot;
    for (int q = 0; q < ntot; ++q) bload[q] = 0.0;
    for (int ig = 0; ig < nip; ++ig) {
      double *et = etensor + (iel * nip + ig) * 6, *te = tensor + (iel * nip + ig) * 6;
      memcpy(dee, dee0, sizeof dee);
      const double det = gauss_point(nod, points, nip, ig, g_coord_pp + iel * nod * 3, der, deriv);
      orc_beemat6(bee, deriv, nod);
      for (int r = 0; r < 6; ++r) {
        double s = 0.0;
        for (int q = 0; q < ntot; ++q) s += bee[q * 6 + r] * eld[q];
        eps[r] = s - et[r];
      }
      for (int r = 0; r < 6; ++r) {
        double s = 0.0;
        for (int q = 0; q < 6; ++q) s += dee[q * 6 + r] * eps[q];
        sigma[r] = s;
        stressv[r] = s + te[r];
      }
      const double fnew = dsbar6(stressv) - sbary;
      if (fnew >= 0.0) {                 /* yield is violated: scale back to the surface, elasto-plastic dee */
        const double f = dsbar6(te) - sbary, fac = fnew / (fnew - f);
        for (int r = 0; r < 6; ++r) stressv[r] = te[r] + (1.0 - fac) * sigma[r];
        vmpl6(e, v, stressv, pl);
        for (int q = 0; q < 36; ++q) dee[q] = dee[q] - fac * pl[q];
      }
      for (int r = 0; r < 6; ++r) {
        double s = 0.0;
        for (int q = 0; q < 6; ++q) s += dee[q * 6 + r] * eps[q];
        sigma[r] = s + te[r];
      }
      for (int q = 0; q < ntot; ++q) {   /* eload = MATMUL(sigma,bee); bload = bload + eload*det*weights(i) */
        double s = 0.0;
        for (int r = 0; r < 6; ++r) s += sigma[r] * bee[q * 6 + r];
        bload[q] = bload[q] + s * det * weights[ig];
      }
      for (int r = 0; r < 6; ++r) { te[r] = sigma[r]; et[r] = et[r] + eps[r]; }
    }
    for (int q = 0; q < ntot; ++q) utemp[iel * ntot + q] = 0.0 - bload[q];
  }
  return 0;
}

/* The same update in OPERATOR FORM -- the arithmetic of the tensor-core kernel k_p1210_mf (the matrix-free kernel
 * k_apply_mf4 with p1210's Gauss-point update in the middle), stated as plain C:
 *   jac / H: node-ascending fma chains from 0.0; inverse = adjugate x ONE reciprocal (invert3_recip); f = det*w
 *   G = inv H (product + two fma), eps = beemat's rows of G minus etensor
 *   sigma_el = D eps with deemat's structural zeros left out (3-term chains, single products), stressv = sigma_el + tensor
 *   yield (dsbar - sbary >= 0): fac, the scaled-back stress, vmpl and dee - fac*pl exactly as elements_2; then
 *       sigma(r) = dee'(r,0)*eps(0) + fma chain over q = 1..5, + tensor;   no yield: sigma = stressv
 *   tensor = sigma, etensor += eps;  S = sigma*f;  T = inv^T S (product + two fma)
 *   u_c(m) = ONE fma chain from 0.0 over (h | b | q), Gauss point 2q+h (the k order of the phase-3 mma);  utemp = 0 - u
 * A different rounding of elements_2 (p1210.f90:120-147), pinned to the same golden fields. */
int orc_p1210_elements_mf(int64_t nels, int nod, int nip, const double *g_coord_pp, double e, double v, double sbary,
                          const double *pmul, double *etensor, double *tensor, double *utemp) {
  if (nod != 20 || nip != 8) return 1;
  const int ntot = 3 * nod;
  double points[81], weights[27], dee0[36], der[8][60], d3[60];
  orc_sample_hex(nip, points, weights);
  orc_deemat6(dee0, e, v);
  for (int ig = 0; ig < nip; ++ig) {
    orc_shape_der(nod, points, nip, ig, d3);
    for (int m = 0; m < nod; ++m)
      for (int a = 0; a < 3; ++a) der[ig][a * 20 + m] = d3[m * 3 + a];
  }
#pragma omp parallel for schedule(static) if (nels > 256)
  for (int64_t iel = 0; iel < nels; ++iel) {
    const double *coord = g_coord_pp + iel * nod * 3, *pm = pmul + iel * ntot;
    double T[8][9];
    for (int ig = 0; ig < 8; ++ig) {
      double *et = etensor + (iel * nip + ig) * 6, *te = tensor + (iel * nip + ig) * 6;
      double jac[9], inv[9], H[9], G[9], sigma[6], stressv[6], dee[36], pl[36];
      for (int b = 0; b < 3; ++b)
        for (int a = 0; a < 3; ++a) {
          double s = 0.0;
          for (int m = 0; m < nod; ++m) s = fma(der[ig][a * 20 + m], coord[b * nod + m], s);
          jac[b * 3 + a] = s;
        }
      const double det = orc_determinant3(jac);
      memcpy(inv, jac, sizeof inv);
      invert3_recip(inv);
      const double f = det * weights[ig];
      for (int b = 0; b < 3; ++b)
        for (int c = 0; c < 3; ++c) {
          double s = 0.0;
          for (int m = 0; m < nod; ++m) s = fma(der[ig][b * 20 + m], pm[3 * m + c], s);
          H[b * 3 + c] = s;
        }
      for (int a = 0; a < 3; ++a)
        for (int c = 0; c < 3; ++c) {
          double s = inv[a] * H[c];
          s = fma(inv[3 + a], H[3 + c], s);
          s = fma(inv[6 + a], H[6 + c], s);
          G[a * 3 + c] = s;
        }
      double eps[6] = {G[0], G[4], G[8], G[3] + G[1], G[7] + G[5], G[6] + G[2]};
      for (int r = 0; r < 6; ++r) eps[r] = eps[r] - et[r];
      for (int r = 0; r < 3; ++r) {
        double s = dee0[r] * eps[0];
        s = fma(dee0[6 + r], eps[1], s);
        s = fma(dee0[12 + r], eps[2], s);
        sigma[r] = s;
      }
      for (int r = 3; r < 6; ++r) sigma[r] = dee0[r * 6 + r] * eps[r];
      for (int r = 0; r < 6; ++r) stressv[r] = sigma[r] + te[r];
      const double fnew = dsbar6(stressv) - sbary;
      if (fnew >= 0.0) {
        const double fy = dsbar6(te) - sbary, fac = fnew / (fnew - fy);
        for (int r = 0; r < 6; ++r) stressv[r] = te[r] + (1.0 - fac) * sigma[r];
        vmpl6(e, v, stressv, pl);
        for (int q = 0; q < 36; ++q) dee[q] = dee0[q] - fac * pl[q];
        for (int r = 0; r < 6; ++r) {
          double s = dee[r] * eps[0];
          for (int q = 1; q < 6; ++q) s = fma(dee[q * 6 + r], eps[q], s);
          sigma[r] = s + te[r];
        }
      } else {
        for (int r = 0; r < 6; ++r) sigma[r] = stressv[r];
      }
      for (int r = 0; r < 6; ++r) { te[r] = sigma[r]; et[r] = et[r] + eps[r]; }
      double sg[6];
      for (int r = 0; r < 6; ++r) sg[r] = sigma[r] * f;
      const double S[9] = {sg[0], sg[3], sg[5], sg[3], sg[1], sg[4], sg[5], sg[4], sg[2]};
      for (int b = 0; b < 3; ++b)
        for (int c = 0; c < 3; ++c) {
          double s = inv[b * 3] * S[c];
          s = fma(inv[b * 3 + 1], S[3 + c], s);
          s = fma(inv[b * 3 + 2], S[6 + c], s);
          T[ig][b * 3 + c] = s;
        }
    }
    for (int m = 0; m < nod; ++m)
      for (int c = 0; c < 3; ++c) {
        double s = 0.0;
        for (int h = 0; h < 2; ++h)
          for (int b = 0; b < 3; ++b)
            for (int q = 0; q < 4; ++q) s = fma(T[2 * q + h][b * 3 + c], der[2 * q + h][b * 20 + m], s);
        utemp[iel * ntot + 3 * m + c] = 0.0 - s;
      }
  }
  return 0;
}

static int g_p1210_form = 0;   /* 0: elements_2 as written (orc_p1210_elements); 1: operator form (orc_p1210_elements_mf) */
void orc_set_p1210_form(int form) { g_p1210_form = form ? 1 : 0; }

/* The whole program on global arrays over npes emulated ranks (the partition only orders the scatter's sums).
 * mm_out (neq) = the assembled lumped mass; snap holds, for every npri-th step, x1 / d1x1 / d2x1 (3 x neq doubles). */
int orc_p1210_run(int64_t nels, int nod, int nip, const double *g_coord_pp, const int32_t *g_g, int64_t neq,
                  const double *fext, double e, double v, double sbary, double rho, double dtim, double pload, int nstep,
                  int npri, int npes, double *mm_out, double *snap) {
  if (nod != 20 || nip != 8 || npri < 1) return 1;
  const int ntot = 3 * nod;
#ifdef _OPENMP
  /* a handful of elements and hundreds of thousands of steps: opening parallel regions would be all the time there is */
  const int threads_before = omp_get_max_threads();
  if (nels <= 256) omp_set_num_threads(1);
#endif
  orc_ranks *R = ranks_new(npes, ntot, nels, g_g, neq);
  double *pmul = malloc(sizeof(double) * (size_t)(nels * ntot)), *utemp = malloc(sizeof(double) * (size_t)(nels * ntot));
  double *x1 = calloc((size_t)neq, 8), *d1 = calloc((size_t)neq, 8), *d2 = calloc((size_t)neq, 8);
  double *mm = calloc((size_t)neq, 8), *bdy = calloc((size_t)neq, 8);
  double *et = calloc((size_t)(nels * nip * 6), 8), *te = calloc((size_t)(nels * nip * 6), 8);
  orc_p1210_mass(nels, nod, nip, g_coord_pp, rho, utemp);
  ranks_scatter(R, ntot, g_g, utemp, mm);
  if (mm_out) memcpy(mm_out, mm, sizeof(double) * (size_t)neq);
  int64_t nout = 0;
  for (int jj = 1; jj <= nstep; ++jj) {
    for (int64_t i = 0; i < neq; ++i) x1[i] = x1[i] + (dtim * d1[i]) + (0.5 * (dtim * dtim) * d2[i]);   /* p1210.f90:117 */
    orc_gather(ntot, nels, g_g, x1, pmul);
    if (g_p1210_form) orc_p1210_elements_mf(nels, nod, nip, g_coord_pp, e, v, sbary, pmul, et, te, utemp);
    else orc_p1210_elements(nels, nod, nip, g_coord_pp, e, v, sbary, pmul, et, te, utemp);
    ranks_scatter(R, ntot, g_g, utemp, bdy);
    for (int64_t i = 0; i < neq; ++i) {                       /* :148-150 */
      double b = bdy[i] + fext[i] * pload;
      b = b / mm[i];
      d1[i] = d1[i] + (d2[i] + b) * .5 * dtim;
      d2[i] = b;
    }
    if (jj % npri == 0 && snap) {
      memcpy(snap + nout * 3 * neq, x1, sizeof(double) * (size_t)neq);
      memcpy(snap + nout * 3 * neq + neq, d1, sizeof(double) * (size_t)neq);
      memcpy(snap + nout * 3 * neq + 2 * neq, d2, sizeof(double) * (size_t)neq);
      nout = nout + 1;
    }
  }
  free(pmul); free(utemp); free(x1); free(d1); free(d2); free(mm); free(bdy); free(et); free(te);
  ranks_free(R);
#ifdef _OPENMP
  omp_set_num_threads(threads_before);
#endif
  return 0;
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
