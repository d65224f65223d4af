// kernels.cuh -- sm_100a kernels of the EBE-PCG hot path.
//
// The stored-storkm path (the reference's path, the headline) is FP64 and HBM-bound; nothing of it is
// tensor-core work.  The file is compiled with --fmad=false so that every a*b+c is a separate IEEE multiply
// and add: the kernels then produce the same bits as oracle/pf_oracle.c
// (-ffp-contract=off) because they also use the same summation orders:
//   mat-vec     u_i = sum_j K(i,j) p_j, j ascending        (p121.f90:93-97)
//   scatter     contributions in ascending element order    (gather_scatter.f90:759-761)
//   reductions  the fixed blocked tree described at block_tree() below
// The kernels are bandwidth-bound, so the lost FMA throughput costs nothing.
// The MATRIX-FREE variant (BASELINE config E: k_apply_mf / k_apply_mf2 / k_apply_mf3 / k_apply_mf4, and p1210's
// operator form) is the exception: it is FP64-flop-bound, uses explicit fma() chains and -- k_apply_mf3 / k_apply_mf4 --
// the FP64 tensor instruction mma.sync.m8n8k4.f64, which is a k-ascending fma chain bit for bit; the oracle mirrors
// those chains (orc_apply_mf, orc_p1210_elements_mf).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pf {

// ----------------------------------------------------------------------------
// device-resident solver state (one per handle)
// ----------------------------------------------------------------------------
struct State {
  double up;         // r.d at the top of the iteration (p121.f90:99)
  double pu;         // p.u
  double alpha, beta;
  double rd_new;     // r.d after the update (numerator of beta)
  double maxloads, maxdiff, ratio;
  double loc[4];     // this rank's partials: [0] dot, [1] max|xnew|, [2] max|xnew-x|
  double tol;
  int iters, limit;
  int done, converged;
  unsigned int ticket[4];  // "last block" counters, one per reducing kernel
  int fault, pad_;         // 1: a peer rank did not arrive within kSpinTimeoutNs (the solve is abandoned, `done` raised)
};

// ----------------------------------------------------------------------------
// small PTX wrappers: mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP)
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// global -> shared bulk copy completing on an mbarrier; streaming data: L2 evict_first
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// ----------------------------------------------------------------------------
// a7+a8 fused: pmul = p(ggl) gathered into shared memory, utemp = K_e * pmul
// ----------------------------------------------------------------------------
// One persistent CTA per SM.  Warp w owns ring slot w: it issues a 1-D bulk copy
// of the next tile of EPT element matrices (the Fortran layout
// storkm_pp(ntot,ntot,iel) is already contiguous per element), gathers the tile's
// p values while the copy is in flight, waits on the slot's mbarrier, multiplies,
// stores utemp and re-arms the slot.  Every byte of storkm is read once; lanes
// hold two rows each so shared-memory reads are conflict-free 128-bit loads.
struct PeerTable;
__device__ __forceinline__ void warp_wait_fwd(PeerTable *T, const State *st);

template <int NTOT, int EPT, int STAGES>
struct MatvecCfg {
  static constexpr int kTileDoubles = EPT * NTOT * NTOT;
  static constexpr int kTileBytes = kTileDoubles * 8;
  static constexpr int kPmDoubles = EPT * NTOT;
  static constexpr int kThreads = STAGES * 32;
  static constexpr size_t kSmem = (size_t)STAGES * kTileBytes + (size_t)STAGES * kPmDoubles * 8 + STAGES * 8;
};

template <int NTOT, int EPT, int STAGES, bool GATHER>
__global__ void __launch_bounds__(STAGES * 32, 1)
k_matvec(const double *__restrict__ km, const int *__restrict__ ggl, const double *__restrict__ pvec,
         double *__restrict__ utemp, long long nels, const State *st, PeerTable *T) {
  using Cfg = MatvecCfg<NTOT, EPT, STAGES>;
  if (st && *(volatile const int *)&st->done) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *tiles = reinterpret_cast<double *>(smem_raw);
  double *pm_all = tiles + (size_t)STAGES * Cfg::kTileDoubles;
  uint64_t *bars = reinterpret_cast<uint64_t *>(pm_all + STAGES * Cfg::kPmDoubles);

  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ntiles = (nels + EPT - 1) / EPT;
  const long long t0 = ntiles * blockIdx.x / gridDim.x;
  const long long t1 = ntiles * (blockIdx.x + 1) / gridDim.x;
  double *tile = tiles + (size_t)w * Cfg::kTileDoubles;
  double *pm = pm_all + w * Cfg::kPmDoubles;
  const uint32_t bar = smem_u32(&bars[w]);
  const uint32_t tile_s = smem_u32(tile);
  uint64_t policy = 0;
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    policy = policy_evict_first();
  }
  __syncwarp();

  auto issue = [&](long long t) {
    const long long e0 = t * EPT;
    const int ne = (int)((nels - e0) < EPT ? (nels - e0) : EPT);
    const uint32_t bytes = (uint32_t)ne * NTOT * NTOT * 8;
    mbar_expect_tx(bar, bytes);
    bulk_g2s(tile_s, km + e0 * (long long)(NTOT * NTOT), bytes, bar, policy);
  };

  long long t = t0 + w;
  if (t < t1 && lane == 0) issue(t);
  // N ranks, peer transport: the owners' values of this iteration must have landed in my halo segment of pvec
  // before the first gather; the first tile of element matrices is already on its way
  if (GATHER && T) warp_wait_fwd(T, st);
  uint32_t phase = 0;
  constexpr int RP = NTOT / 2;  // row pairs per element
  for (; t < t1; t += STAGES) {
    const long long e0 = t * EPT;
    const int ne = (int)((nels - e0) < EPT ? (nels - e0) : EPT);
    // gather (or copy) the tile's right-hand sides while the bulk copy flies (peer-written values: L2 loads)
    for (int s = lane; s < ne * NTOT; s += 32) {
      if (GATHER) pm[s] = T ? __ldcg(pvec + ggl[e0 * NTOT + s]) : pvec[ggl[e0 * NTOT + s]];
      else pm[s] = pvec[e0 * NTOT + s];
    }
    __syncwarp();
    mbar_wait(bar, phase);
    phase ^= 1;
    for (int s = lane; s < ne * RP; s += 32) {
      const int el = s / RP, rp = s - el * RP;
      const double *K = tile + el * (NTOT * NTOT) + 2 * rp;
      const double *pv = pm + el * NTOT;
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int j = 0; j < NTOT; j += 2) {
        const double2 pj = *reinterpret_cast<const double2 *>(pv + j);
        const double2 k0 = *reinterpret_cast<const double2 *>(K + j * NTOT);
        const double2 k1 = *reinterpret_cast<const double2 *>(K + (j + 1) * NTOT);
        a0 = a0 + k0.x * pj.x; a1 = a1 + k0.y * pj.x;
        a0 = a0 + k1.x * pj.y; a1 = a1 + k1.y * pj.y;
      }
      *reinterpret_cast<double2 *>(utemp + (e0 + el) * NTOT + 2 * rp) = make_double2(a0, a1);
    }
    __syncwarp();
    const long long tn = t + STAGES;
    if (tn < t1 && lane == 0) {
      fence_proxy_async();  // order this warp's generic reads of the slot before the async overwrite
      issue(tn);
    }
  }
}

// k_matvec2: the same product with NBUF ring slots PER WARP (round 2).  k_matvec re-arms a warp's only slot after the
// tile has been multiplied, so the slot carries no bytes while the warp computes; with 8x8 matrices (p123 / p124 / p125:
// 8 KB tiles) the copy latency is most of a warp's cycle and the 16 slots of an SM hold ~94 KB in flight on average --
// measured 0.104 ms = 0.755 of the HBM peak at config B.  Here the copy of tile t+NBUF*WARPS is issued into the slot tile
// t has just left, NBUF-1 copies per warp are always in flight while it gathers and multiplies.  Same arithmetic, same bits.
template <int NTOT, int EPT, int WARPS, int NBUF>
struct Matvec2Cfg {
  static constexpr int kTileDoubles = EPT * NTOT * NTOT, kTileBytes = kTileDoubles * 8, kPmDoubles = EPT * NTOT;
  static constexpr int kThreads = WARPS * 32;
  static constexpr size_t kSmem = (size_t)WARPS * NBUF * kTileBytes + (size_t)WARPS * kPmDoubles * 8 + (size_t)WARPS * NBUF * 8;
};

template <int NTOT, int EPT, int WARPS, int NBUF, bool GATHER>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_matvec2(const double *__restrict__ km, const int *__restrict__ ggl, const double *__restrict__ pvec,
          double *__restrict__ utemp, long long nels, const State *st, PeerTable *T) {
  using Cfg = Matvec2Cfg<NTOT, EPT, WARPS, NBUF>;
  constexpr int KP = (EPT * NTOT + 31) / 32, RP = NTOT / 2;
  if (st && *(volatile const int *)&st->done) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *tiles = reinterpret_cast<double *>(smem_raw);
  double *pm_all = tiles + (size_t)WARPS * NBUF * Cfg::kTileDoubles;
  uint64_t *bars = reinterpret_cast<uint64_t *>(pm_all + WARPS * Cfg::kPmDoubles);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ntiles = (nels + EPT - 1) / EPT;
  const long long t0 = ntiles * blockIdx.x / gridDim.x, t1 = ntiles * (blockIdx.x + 1) / gridDim.x;
  double *mytiles = tiles + (size_t)w * NBUF * Cfg::kTileDoubles;
  double *pm = pm_all + w * Cfg::kPmDoubles;
  uint64_t policy = 0;
  if (lane == 0) {
    for (int b = 0; b < NBUF; ++b) mbar_init(smem_u32(&bars[w * NBUF + b]), 1);
    fence_mbar_init();
    policy = policy_evict_first();
  }
  __syncwarp();
  auto issue = [&](int b, long long t) {                       // lane 0 only
    const long long e0 = t * EPT;
    const int ne = (int)((nels - e0) < EPT ? (nels - e0) : EPT);
    const uint32_t bytes = (uint32_t)ne * NTOT * NTOT * 8, bar = smem_u32(&bars[w * NBUF + b]);
    mbar_expect_tx(bar, bytes);
    bulk_g2s(smem_u32(mytiles + (size_t)b * Cfg::kTileDoubles), km + e0 * (long long)(NTOT * NTOT), bytes, bar, policy);
  };
  long long t = t0 + w;
  if (lane == 0)
    for (int b = 0; b < NBUF; ++b)
      if (t + (long long)b * WARPS < t1) issue(b, t + (long long)b * WARPS);
  // N ranks, peer transport: the owners' values must have landed before the first gather; the first tiles are on their way
  if (GATHER && T) warp_wait_fwd(T, st);
  int b = 0;
  uint32_t phase = 0;
  for (; t < t1; t += WARPS) {
    const long long e0 = t * EPT;
    const int ne = (int)((nels - e0) < EPT ? (nels - e0) : EPT);
    const double *tile = mytiles + (size_t)b * Cfg::kTileDoubles;
    // the tile's right-hand sides: every index first, then every value (two dependent loads per word, all in flight)
    if (ne == EPT) {
      int idx[KP];
      double val[KP];
#pragma unroll
      for (int kp = 0; kp < KP; ++kp) {
        const int s = lane + 32 * kp;
        idx[kp] = (GATHER && s < EPT * NTOT) ? ggl[e0 * NTOT + s] : 0;
      }
#pragma unroll
      for (int kp = 0; kp < KP; ++kp) {
        const int s = lane + 32 * kp;
        if (GATHER) val[kp] = T ? __ldcg(pvec + idx[kp]) : pvec[idx[kp]];
        else val[kp] = (s < EPT * NTOT) ? pvec[e0 * NTOT + s] : 0.0;
      }
#pragma unroll
      for (int kp = 0; kp < KP; ++kp) {
        const int s = lane + 32 * kp;
        if (s < EPT * NTOT) pm[s] = val[kp];
      }
    } else {
      for (int s = lane; s < ne * NTOT; s += 32) {
        if (GATHER) pm[s] = T ? __ldcg(pvec + ggl[e0 * NTOT + s]) : pvec[ggl[e0 * NTOT + s]];
        else pm[s] = pvec[e0 * NTOT + s];
      }
    }
    __syncwarp();
    mbar_wait(smem_u32(&bars[w * NBUF + b]), phase);
    for (int s = lane; s < ne * RP; s += 32) {
      const int el = s / RP, rp = s - el * RP;
      const double *K = tile + el * (NTOT * NTOT) + 2 * rp;
      const double *pv = pm + el * NTOT;
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int j = 0; j < NTOT; j += 2) {
        const double2 pj = *reinterpret_cast<const double2 *>(pv + j);
        const double2 k0 = *reinterpret_cast<const double2 *>(K + j * NTOT);
        const double2 k1 = *reinterpret_cast<const double2 *>(K + (j + 1) * NTOT);
        a0 = a0 + k0.x * pj.x; a1 = a1 + k0.y * pj.x;
        a0 = a0 + k1.x * pj.y; a1 = a1 + k1.y * pj.y;
      }
      *reinterpret_cast<double2 *>(utemp + (e0 + el) * NTOT + 2 * rp) = make_double2(a0, a1);
    }
    __syncwarp();
    const long long tn = t + (long long)NBUF * WARPS;
    if (tn < t1 && lane == 0) {
      fence_proxy_async();  // order this warp's generic reads of the slot before the async overwrite
      issue(b, tn);
    }
    if (++b == NBUF) { b = 0; phase ^= 1; }
  }
}

// ----------------------------------------------------------------------------
// a8 on the symmetric-packed layout (pf_set_storkm_layout(h, 1)): half the storkm stream.
// ----------------------------------------------------------------------------
// Only the lower triangle of every element matrix is stored, packed by columns: column j holds
// rows j..NTOT-1 at offset coloff(j) = j*NTOT - j(j-1)/2 (1 830 doubles = 14 640 B per 20-node
// brick instead of 28 800 B).  The product keeps the reference's order, u_r = sum_j K(r,j) p_j with
// j ascending from 0.0 and separate multiply / add, with K(r,j) = L(max(r,j), min(r,j)): it equals
// MATMUL(storkm,pmul) of the symmetrised matrix bit for bit (and of storkm itself wherever storkm
// is bitwise symmetric).  Lane t of an element owns rows t and NTOT-1-t, so every lane does the
// same work: for j <= r the entry sits in column j (lanes read consecutive words), for j > r in
// the lane's own column r.  Same persistent ring of bulk-copied tiles as k_matvec.
template <int NTOT>
struct SymCfg {
  static constexpr int kPacked = NTOT * (NTOT + 1) / 2;
  __host__ __device__ static constexpr int coloff(int j) { return j * NTOT - j * (j - 1) / 2; }
};
template <int NTOT, int EPT, int STAGES>
struct MatvecSymCfg {
  static constexpr int kTileDoubles = EPT * SymCfg<NTOT>::kPacked;
  static constexpr int kTileBytes = kTileDoubles * 8;
  static constexpr int kPmDoubles = EPT * NTOT;
  static constexpr int kThreads = STAGES * 32;
  static constexpr size_t kSmem = (size_t)STAGES * kTileBytes + (size_t)STAGES * kPmDoubles * 8 + STAGES * 8;
};

template <int NTOT, int EPT, int STAGES, bool GATHER>
__global__ void __launch_bounds__(STAGES * 32, 1)
k_matvec_sym(const double *__restrict__ kp, const int *__restrict__ ggl, const double *__restrict__ pvec,
             double *__restrict__ utemp, long long nels, const State *st, PeerTable *T) {
  using Cfg = MatvecSymCfg<NTOT, EPT, STAGES>;
  constexpr int P = SymCfg<NTOT>::kPacked, LPE = NTOT / 2, EPW = 32 / LPE;   // lanes per element, elements per warp pass
  static_assert(NTOT % 2 == 0 && LPE <= 32 && (EPT % EPW == 0 || EPW == 1), "tile shape");
  if (st && *(volatile const int *)&st->done) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *tiles = reinterpret_cast<double *>(smem_raw);
  double *pm_all = tiles + (size_t)STAGES * Cfg::kTileDoubles;
  uint64_t *bars = reinterpret_cast<uint64_t *>(pm_all + STAGES * Cfg::kPmDoubles);

  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ntiles = (nels + EPT - 1) / EPT;
  const long long t0 = ntiles * blockIdx.x / gridDim.x;
  const long long t1 = ntiles * (blockIdx.x + 1) / gridDim.x;
  double *tile = tiles + (size_t)w * Cfg::kTileDoubles;
  double *pm = pm_all + w * Cfg::kPmDoubles;
  const uint32_t bar = smem_u32(&bars[w]);
  const uint32_t tile_s = smem_u32(tile);
  uint64_t policy = 0;
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    policy = policy_evict_first();
  }
  __syncwarp();

  auto issue = [&](long long t) {
    const long long e0 = t * EPT;
    const int ne = (int)((nels - e0) < EPT ? (nels - e0) : EPT);
    const uint32_t bytes = (uint32_t)ne * P * 8;
    mbar_expect_tx(bar, bytes);
    bulk_g2s(tile_s, kp + e0 * (long long)P, bytes, bar, policy);
  };

  // this lane's two rows and where their own columns start (minus the row number)
  const int sub = lane / LPE, tl = lane - sub * LPE;
  const int r0 = tl, r1 = NTOT - 1 - tl;
  const int c0 = SymCfg<NTOT>::coloff(r0) - r0, c1 = SymCfg<NTOT>::coloff(r1) - r1;
  const bool lane_on = sub < EPW;

  long long t = t0 + w;
  if (t < t1 && lane == 0) issue(t);
  if (GATHER && T) warp_wait_fwd(T, st);
  uint32_t phase = 0;
  for (; t < t1; t += STAGES) {
    const long long e0 = t * EPT;
    const int ne = (int)((nels - e0) < EPT ? (nels - e0) : EPT);
    for (int s = lane; s < ne * NTOT; s += 32) {
      if (GATHER) pm[s] = T ? __ldcg(pvec + ggl[e0 * NTOT + s]) : pvec[ggl[e0 * NTOT + s]];
      else pm[s] = pvec[e0 * NTOT + s];
    }
    __syncwarp();
    mbar_wait(bar, phase);
    phase ^= 1;
    for (int eb = 0; eb < ne; eb += EPW) {
      const int el = eb + sub;
      const bool on = lane_on && el < ne;
      const double *L = tile + (on ? el : 0) * P;
      const double *pv = pm + (on ? el : 0) * NTOT;
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int j = 0; j < NTOT; ++j) {
        const double pj = pv[j];
        const int o0 = (r0 >= j) ? (SymCfg<NTOT>::coloff(j) - j) + r0 : c0 + j;
        const int o1 = (r1 >= j) ? (SymCfg<NTOT>::coloff(j) - j) + r1 : c1 + j;
        a0 = a0 + L[o0] * pj;
        a1 = a1 + L[o1] * pj;
      }
      if (on) {
        utemp[(e0 + el) * NTOT + r0] = a0;
        utemp[(e0 + el) * NTOT + r1] = a1;
      }
    }
    __syncwarp();
    const long long tn = t + STAGES;
    if (tn < t1 && lane == 0) {
      fence_proxy_async();
      issue(tn);
    }
  }
}

// ----------------------------------------------------------------------------
// a14: pcg_km (maths.f90:1152-1323) -- ONE element matrix km(ntot,ntot) shared by every element
// (utemp_pp = MATMUL(km,pmul_pp)).  Nothing streams from HBM but the gather indices and the vectors: thread
// (element slot, row) keeps its row of km in registers, the right-hand sides of the block's elements are staged in
// shared memory, every product is the j-ascending sum from 0.0 with separate multiply and add -- the same bits as
// MATMUL on a replicated storkm_pp.  FP64-pipe bound.
// ----------------------------------------------------------------------------
template <int NTOT, bool GATHER>
__global__ void __launch_bounds__(64)
k_matvec_km(const double *__restrict__ km1, const int *__restrict__ ggl, const double *__restrict__ pvec,
            double *__restrict__ utemp, long long nels, const State *st, PeerTable *T) {
  constexpr int THREADS = 64, EPB = THREADS / NTOT;        // elements per block pass
  static_assert(NTOT <= THREADS, "one thread per row");
  if (st && *(volatile const int *)&st->done) return;
  __shared__ double pm[EPB * NTOT];
  const int t = threadIdx.x, el = t / NTOT, row = t - el * NTOT;
  const bool on = el < EPB;
  double K[NTOT];
#pragma unroll
  for (int j = 0; j < NTOT; ++j) K[j] = on ? km1[j * NTOT + row] : 0.0;
  if (GATHER && T) { if (t < 32) warp_wait_fwd(T, st); __syncthreads(); }
  const long long npass = (nels + EPB - 1) / EPB;
  for (long long b = blockIdx.x; b < npass; b += gridDim.x) {
    const long long e0 = b * EPB;
    const int ne = (int)((nels - e0) < EPB ? (nels - e0) : EPB);
    __syncthreads();
    for (int q = t; q < ne * NTOT; q += THREADS) {
      if (GATHER) pm[q] = T ? __ldcg(pvec + ggl[e0 * NTOT + q]) : pvec[ggl[e0 * NTOT + q]];
      else pm[q] = pvec[e0 * NTOT + q];
    }
    __syncthreads();
    if (on && el < ne) {
      const double *pv = pm + el * NTOT;
      double a = 0.0;
#pragma unroll
      for (int j = 0; j < NTOT; ++j) a = a + K[j] * pv[j];
      utemp[(e0 + el) * NTOT + row] = a;
    }
  }
}

// pmul materialised (pf_gather only; the solver never does this)
__global__ void k_gather(const int *__restrict__ ggl, const double *__restrict__ p_ext,
                         double *__restrict__ pmul, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) pmul[i] = p_ext[ggl[i]];
}

// owner side of the reverse halo exchange: add received partial sums, sources in
// ascending rank order (acc_pos is sorted that way per equation)
__global__ void k_halo_accumulate(const int *__restrict__ acc_slot, const unsigned int *__restrict__ acc_ptr,
                                  const unsigned int *__restrict__ acc_pos, const double *__restrict__ recv,
                                  double *__restrict__ u_ext, int n, const State *st) {
  if (st && *(volatile const int *)&st->done) return;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc = u_ext[acc_slot[i]];
  for (unsigned int k = acc_ptr[i]; k < acc_ptr[i + 1]; ++k) acc = acc + recv[acc_pos[k]];
  u_ext[acc_slot[i]] = acc;
}

// pack owned values wanted by other ranks (forward halo exchange)
__global__ void k_halo_pack(const int *__restrict__ put_slot, const double *__restrict__ p_ext,
                            double *__restrict__ sendbuf, int n, const State *st) {
  if (st && *(volatile const int *)&st->done) return;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sendbuf[i] = p_ext[put_slot[i]];
}


// ----------------------------------------------------------------------------
// peer-memory collectives over NVLink/NVSwitch (one process per GPU, buffers mapped
// with CUDA IPC).  Inside the PCG loop the halo exchanges and the scalar reductions are
// plain st.global / ld.global on peer pointers plus release/acquire flags, issued from
// the same kernels that produce / consume the data -- no NCCL launch, no host round trip.
// Safety of buffer reuse: every PCG iteration passes two all-rank reductions, so a rank
// cannot overwrite a peer's halo segment or receive buffer of iteration k+1 before that
// peer has consumed iteration k.
// ----------------------------------------------------------------------------
constexpr int kMaxRanks = 16;
struct PeerSync {                       // lives in every rank's exported buffer; written by peers
  unsigned long long fwd_flag[kMaxRanks];
  unsigned long long rev_flag[kMaxRanks];
  unsigned long long red_flag[3][kMaxRanks];
  double red_val[3][2][kMaxRanks][4];   // [reduction][seq parity][source rank][dot,max,max,-]
};
struct PeerTable {                      // local; pointers into the peers' address ranges
  int rank, nranks;
  PeerSync *sync[kMaxRanks];
  double *p_ext[kMaxRanks];
  double *recv[kMaxRanks];
  long long fwd_dst_off[kMaxRanks];     // where my owned values land in peer r's p_ext
  long long rev_dst_off[kMaxRanks];     // where my partial sums land in owner r's receive buffer
  long long put_off[kMaxRanks + 1];     // my put list, segmented by destination rank
  long long get_off[kMaxRanks + 1];     // my halo slots, segmented by owner rank
  unsigned long long seq_fwd, seq_rev, seq_red[3];
  unsigned int ticket[2];
};
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Every wait on a peer is bounded: after kSpinTimeoutNs without the flag the rank raises State::fault and `done`
// (every kernel of the solve then returns at once) and pf_pcg_run reports the failure -- a rank that died or left a
// collective section early can no longer hang its peers for ever.
constexpr unsigned long long kSpinTimeoutNs = 30ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool spin_until(const unsigned long long *flag, unsigned long long seq, const State *st) {
  if (ld_acquire_sys(flag) >= seq) return true;
  const unsigned long long t0 = globaltimer_ns();
  unsigned int n = 0;
  while (ld_acquire_sys(flag) < seq) {
    if ((++n & 255u) == 0 && globaltimer_ns() - t0 > kSpinTimeoutNs) {
      if (st) { State *w = const_cast<State *>(st); w->fault = 1; w->done = 1; __threadfence(); }
      return false;
    }
  }
  return true;
}
// called by a whole warp: wait until every owner I gather from has stored this iteration's values into my halo
// segment (its k_pupdate / k_halo_put_peer released fwd_flag[owner] = seq_fwd)
__device__ __forceinline__ void warp_wait_fwd(PeerTable *T, const State *st) {
  const int lane = threadIdx.x & 31, me = T->rank;
  if (lane < T->nranks && lane != me && T->get_off[lane + 1] > T->get_off[lane])
    spin_until(&T->sync[me]->fwd_flag[lane], T->seq_fwd, st);
  __syncwarp();
}
// called by a whole block (>= nranks threads): the same for the ranks that send me partial sums
__device__ __forceinline__ void block_wait_rev(PeerTable *T, const State *st) {
  const int t = threadIdx.x, me = T->rank;
  if (t < T->nranks && t != me && T->put_off[t + 1] > T->put_off[t])
    spin_until(&T->sync[me]->rev_flag[t], T->seq_rev, st);
  __syncthreads();
}
// every block of a kernel that stored into peer memory calls this once, after its stores: the last block to arrive
// releases the flags (dir 0: forward, my owned values are in the peers' halo segments; dir 1: reverse, my partial
// sums are in the owners' receive buffers) and advances the sequence number
__device__ __forceinline__ void peer_release_flags(PeerTable *T, int dir) {
  __shared__ int flag;
  const unsigned long long seq = (dir == 0 ? T->seq_fwd : T->seq_rev) + 1;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(&T->ticket[dir], 1u);
    flag = (prev == gridDim.x - 1);
    if (flag) T->ticket[dir] = 0;
  }
  __syncthreads();
  if (flag) {
    __threadfence_system();
    const int t = threadIdx.x, me = T->rank;
    if (t < T->nranks && t != me) {
      if (dir == 0 && T->put_off[t + 1] > T->put_off[t]) st_release_sys(&T->sync[t]->fwd_flag[me], seq);
      if (dir == 1 && T->get_off[t + 1] > T->get_off[t]) st_release_sys(&T->sync[t]->rev_flag[me], seq);
    }
    __syncthreads();
    if (t == 0) { if (dir == 0) T->seq_fwd = seq; else T->seq_rev = seq; }
  }
}
// all threads of ONE block call; loc[4] in shared memory; result (rank-ordered sum, maxima)
// valid in thread 0.  Writes my partials into every peer, waits for theirs.
__device__ __forceinline__ void peer_allreduce(PeerTable *T, int which, const double *loc, double *out,
                                               const State *st) {
  const unsigned long long seq = T->seq_red[which] + 1;
  const int t = threadIdx.x, me = T->rank, n = T->nranks;
  const int par = (int)(seq & 1);
  if (t < n && t != me) {
    volatile double *dst = T->sync[t]->red_val[which][par][me];
    dst[0] = loc[0]; dst[1] = loc[1]; dst[2] = loc[2]; dst[3] = loc[3];
    __threadfence_system();
    st_release_sys(&T->sync[t]->red_flag[which][me], seq);
    spin_until(&T->sync[me]->red_flag[which][t], seq, st);
  }
  __syncthreads();
  if (t == 0) {
    double s = 0.0, m1 = 0.0, m2 = 0.0;
    for (int r = 0; r < n; ++r) {
      const volatile double *v = T->sync[me]->red_val[which][par][r];
      const double a = (r == me) ? loc[0] : v[0];
      const double b = (r == me) ? loc[1] : v[1];
      const double c = (r == me) ? loc[2] : v[2];
      s = (r == 0) ? a : s + a;      // ranks ascending, as k_scalars and the oracle
      m1 = fmax(m1, b); m2 = fmax(m2, c);
    }
    out[0] = s; out[1] = m1; out[2] = m2;
    T->seq_red[which] = seq;
  }
}

// forward halo exchange by peer stores: owned values straight into the peers' p_ext halo segments
__global__ void k_halo_put_peer(PeerTable *T, const int *__restrict__ put_slot, const double *__restrict__ p_ext,
                                long long nput, const State *st) {
  if (st && *(volatile const int *)&st->done) return;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nput; k += (long long)gridDim.x * blockDim.x) {
    int r = 0;
    while (k >= T->put_off[r + 1]) ++r;
    T->p_ext[r][T->fwd_dst_off[r] + (k - T->put_off[r])] = p_ext[put_slot[k]];
  }
  peer_release_flags(T, 0);
}
// reverse: partial sums of remote equations straight into their owners' receive buffers
__global__ void k_halo_rev_peer(PeerTable *T, const double *__restrict__ vec_ext, long long neq_pp, long long nhalo,
                                const State *st) {
  if (st && *(volatile const int *)&st->done) return;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nhalo; k += (long long)gridDim.x * blockDim.x) {
    int r = 0;
    while (k >= T->get_off[r + 1]) ++r;
    T->recv[r][T->rev_dst_off[r] + (k - T->get_off[r])] = vec_ext[1 + neq_pp + k];
  }
  peer_release_flags(T, 1);
}
// dir 0: wait for the owners I gather from; dir 1: wait for the ranks that send me partial sums
__global__ void k_halo_wait(PeerTable *T, int dir, const State *st) {
  if (st && *(volatile const int *)&st->done) return;
  const int t = threadIdx.x, me = T->rank;
  if (t < T->nranks && t != me) {
    const bool need = dir == 0 ? (T->get_off[t + 1] > T->get_off[t]) : (T->put_off[t + 1] > T->put_off[t]);
    const unsigned long long seq = dir == 0 ? T->seq_fwd : T->seq_rev;
    const unsigned long long *f = dir == 0 ? &T->sync[me]->fwd_flag[t] : &T->sync[me]->rev_flag[t];
    if (need) spin_until(f, seq, st);
  }
}

// ----------------------------------------------------------------------------
// a9: deterministic scatter as a slot-centric gather.  csr_ptr/csr_pos list, for
// every slot >= 1 of the local gather buffer, the positions e*ntot+k in utemp of
// its contributions in ascending element order.  DIAG: read K_e(k,k) instead.
// N ranks, peer transport (T != nullptr): the partial sum of a halo slot (an equation another rank owns) is also
// stored straight into its owner's receive buffer, and the last block to finish releases the reverse-halo flags --
// the reverse exchange of gather_scatter.f90:694-850 without a kernel of its own.
// ----------------------------------------------------------------------------
template <bool DIAG>
__global__ void k_scatter(const unsigned int *__restrict__ csr_ptr, const unsigned int *__restrict__ csr_pos,
                          const double *__restrict__ src, double *__restrict__ u_ext, long long nslots,
                          int ntot, const State *st, int packed = 0, PeerTable *T = nullptr, long long neq_pp = 0) {
  if (st && *(volatile const int *)&st->done) return;
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x + 1;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; s < nslots; s += stride) {
    const unsigned int a = csr_ptr[s], b = csr_ptr[s + 1];
    double acc = 0.0;
    if (DIAG) {
      for (unsigned int k = a; k < b; ++k) {
        const unsigned int pos = csr_pos[k];
        const unsigned int e = pos / ntot, d = pos - e * ntot;
        if (packed) acc = acc + src[(size_t)e * (ntot * (ntot + 1) / 2) + (size_t)d * ntot - d * (d - 1) / 2];
        else acc = acc + src[(size_t)e * ntot * ntot + (size_t)d * ntot + d];
      }
    } else {
      // kBatch contributions in flight at a time (index loads, then value loads), added in list order
      constexpr int kBatch = 4;   // measured: 4 -> hex8 200^3 0.77 -> 0.58 ms, hex20 0.34 -> 0.32; 8 -> no gain on hex8
      for (unsigned int k = a; k < b; k += kBatch) {
        unsigned int pos[kBatch];
        double v[kBatch];
#pragma unroll
        for (int i = 0; i < kBatch; ++i) pos[i] = (k + i < b) ? __ldg(csr_pos + k + i) : 0xffffffffu;
#pragma unroll
        for (int i = 0; i < kBatch; ++i) v[i] = (pos[i] != 0xffffffffu) ? src[pos[i]] : 0.0;
#pragma unroll
        for (int i = 0; i < kBatch; ++i)
          if (pos[i] != 0xffffffffu) acc = acc + v[i];
      }
    }
    u_ext[s] = acc;
    if (!DIAG && T && s > neq_pp) {
      const long long k = s - 1 - neq_pp;      // halo slots are grouped by owner rank, ascending
      int r = 0;
      while (k >= T->get_off[r + 1]) ++r;
      T->recv[r][T->rev_dst_off[r] + (k - T->get_off[r])] = acc;
    }
  }
  if (!DIAG && T) peer_release_flags(T, 1);
}

// ----------------------------------------------------------------------------
// the fixed blocked reduction tree (mirrored by orc_dot_blocked in the oracle)
//   chunk c = entries [2048c, 2048c+2048); thread t adds the entries
//   2048c+512k+2t and +1 for k = 0..3 in order from 0.0; warp xor-tree
//   off = 16,8,4,2,1; then warps 0..7 in order.  Chunk sums are combined by the
//   last block to finish: thread t adds chunks t, t+256, ... in order, same tree.
// ----------------------------------------------------------------------------
constexpr int kChunk = 2048;
constexpr int kRedThreads = 256;

__device__ __forceinline__ double warp_tree_sum(double v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
__device__ __forceinline__ double warp_tree_max(double v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}
// all 256 threads call; result valid in thread 0
__device__ __forceinline__ double block_tree(double v, double *sh /*8*/) {
  v = warp_tree_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
    s = sh[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) s = s + sh[w];
  }
  return s;
}
__device__ __forceinline__ double block_max(double v, double *sh /*8*/) {
  v = warp_tree_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
    s = sh[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) s = fmax(s, sh[w]);
  }
  return s;
}
// true in every thread of exactly one block: the last to arrive
__device__ __forceinline__ bool last_block(unsigned int *ticket, int *sh_flag) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(ticket, 1u);
    *sh_flag = (prev == gridDim.x - 1);
    if (*sh_flag) *ticket = 0;
  }
  __syncthreads();
  const bool last = *sh_flag != 0;
  if (last) __threadfence();
  return last;
}
__device__ __forceinline__ double final_sum(const double *part, long long nchunks, double *sh) {
  double acc = 0.0;
  for (long long c = threadIdx.x; c < nchunks; c += kRedThreads) acc = acc + __ldcg(part + c);
  return block_tree(acc, sh);
}
__device__ __forceinline__ double final_max(const double *part, long long nchunks, double *sh) {
  double acc = 0.0;
  for (long long c = threadIdx.x; c < nchunks; c += kRedThreads) acc = fmax(acc, __ldcg(part + c));
  return block_max(acc, sh);
}

// scalar epilogues; run by one thread, either inside the last block (1 rank) or
// in k_scalars after the all-gather of the ranks' partials
__device__ __forceinline__ void finish_init(State *st, double rd) { st->up = rd; }
__device__ __forceinline__ void finish_pu(State *st, double pu) {
  st->pu = pu;
  st->alpha = st->up / pu;
}
__device__ __forceinline__ void finish_update(State *st, double rd, double maxloads, double maxdiff,
                                              double *ratio_hist) {
  st->rd_new = rd;
  st->beta = rd / st->up;
  st->up = rd;  // p121.f90:99 recomputes r.d next iteration: same operands, same order, same bits
  st->maxloads = maxloads; st->maxdiff = maxdiff;
  const double ratio = maxdiff / maxloads;  // checon_par, maths.f90:1060
  st->ratio = ratio;
  st->iters = st->iters + 1;
  if (ratio_hist) ratio_hist[st->iters - 1] = ratio;
  const int conv = ratio <= st->tol;
  st->converged = conv;
  // p = d + p*beta of the converging iteration still runs (p121.f90:102-103); `done` is
  // raised by k_pupdate after it
}

// mode: 0 init (up), 1 p.u, 2 update
__global__ void k_scalars(State *st, const double *all /*nranks x 4*/, int nranks, int mode,
                          double *ratio_hist) {
  if (*(volatile const int *)&st->done) return;
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s = all[0], m1 = all[1], m2 = all[2];
  for (int r = 1; r < nranks; ++r) {
    s = s + all[4 * r];
    m1 = fmax(m1, all[4 * r + 1]);
    m2 = fmax(m2, all[4 * r + 2]);
  }
  if (mode == 0) finish_init(st, s);
  else if (mode == 1) finish_pu(st, s);
  else finish_update(st, s, m1, m2, ratio_hist);
}

// d = diag*r, p = d, x = 0, up = r.d     (p121.f90:87; p_ext/u_ext slot 0 stays 0)
__global__ void __launch_bounds__(kRedThreads)
k_pcg_init(const double *__restrict__ diag, const double *__restrict__ r, double *__restrict__ d,
           double *__restrict__ p, double *__restrict__ x, long long n, double *part, State *st,
           int single_rank, PeerTable *T, int keep_x /* p122: the solve starts from the current x (p122.f90:139-146) */) {
  __shared__ double sh[8];
  __shared__ double sh_loc[4], sh_out[3];
  __shared__ int flag;
  const long long nchunks = (n + kChunk - 1) / kChunk;
  for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long i = c * kChunk + 512 * k + 2 * threadIdx.x;
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (i + h < n) {
          const double rr = r[i + h], dd = diag[i + h] * rr;
          d[i + h] = dd; p[i + h] = dd;
          if (!keep_x) x[i + h] = 0.0;
          acc = acc + rr * dd;
        }
    }
    const double s = block_tree(acc, sh);
    if (threadIdx.x == 0) part[c] = s;
  }
  if (last_block(&st->ticket[0], &flag)) {
    const double s = final_sum(part, nchunks, sh);
    if (threadIdx.x == 0) {
      st->loc[0] = s; st->loc[1] = 0.0; st->loc[2] = 0.0; st->loc[3] = 0.0;
      sh_loc[0] = s; sh_loc[1] = 0.0; sh_loc[2] = 0.0; sh_loc[3] = 0.0;
      if (single_rank) finish_init(st, s);
    }
    if (T) {
      __syncthreads();
      peer_allreduce(T, 0, sh_loc, sh_out, st);
      if (threadIdx.x == 0) finish_init(st, sh_out[0]);
    }
  }
}

// p123 fixed freedoms: u(j) = p(j)*store(i)   (p123.f90:141-145); mode 1: u(j) = 0, what p122 does on every plastic
// iteration after the first (p122.f90:154-160)
__global__ void k_fixed_u(const int *__restrict__ fix_slot, const double *__restrict__ store,
                          const double *__restrict__ p_ext, double *__restrict__ u_ext, int n, const State *st, int mode) {
  if (st && *(volatile const int *)&st->done) return;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) u_ext[fix_slot[i]] = mode == 0 ? p_ext[fix_slot[i]] * store[i] : 0.0;
}

// owner side of the reverse halo exchange folded into the p.u reduction (N ranks, peer transport): the accumulate
// entries whose equation lies in chunk c are acc_chunk_ptr[c] .. acc_chunk_ptr[c+1] (acc_slot is ascending)
struct AccTables {
  const int *slot;                 // owned slots that receive partial sums, ascending
  const unsigned int *ptr, *pos;   // per entry: positions in the receive buffer, source ranks ascending
  const unsigned int *chunk_ptr;   // per reduction chunk: first entry
  const double *recv;
  double *u_ext;                   // slot-indexed
};

// pu = p.u over the owned equations.  ACC: the block first waits for the partial sums of the other ranks and adds
// them to its chunk's equations (own partial sum first, then the sources in ascending rank order -- what
// k_halo_accumulate does in the NCCL transport), then reduces.
template <bool ACC>
__global__ void __launch_bounds__(kRedThreads)
k_dot(const double *__restrict__ a, const double *b_, long long n, double *part, State *st,
      int single_rank, int mode /*1: p.u epilogue, -1: plain dot into loc[0]*/, PeerTable *T, AccTables A) {
  if (mode == 1 && *(volatile const int *)&st->done) return;
  __shared__ double sh[8];
  __shared__ double sh_loc[4], sh_out[3];
  __shared__ int flag;
  const long long nchunks = (n + kChunk - 1) / kChunk;
  if (ACC) block_wait_rev(T, st);
  for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
    double acc = 0.0;
    if (ACC) {
      for (unsigned int k = A.chunk_ptr[c] + threadIdx.x; k < A.chunk_ptr[c + 1]; k += kRedThreads) {
        const int slot = A.slot[k];
        double v = A.u_ext[slot];
        for (unsigned int q = A.ptr[k]; q < A.ptr[k + 1]; ++q) v = v + __ldcg(A.recv + A.pos[q]);   // peer-written: L2
        A.u_ext[slot] = v;
      }
      __syncthreads();
      const double *b = b_;          // same memory as A.u_ext + 1: plain (coherent) loads
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long long i = c * kChunk + 512 * k + 2 * threadIdx.x;
#pragma unroll
        for (int h = 0; h < 2; ++h)
          if (i + h < n) acc = acc + a[i + h] * b[i + h];
      }
    } else {
      const double *__restrict__ b = b_;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long long i = c * kChunk + 512 * k + 2 * threadIdx.x;
#pragma unroll
        for (int h = 0; h < 2; ++h)
          if (i + h < n) acc = acc + a[i + h] * b[i + h];
      }
    }
    const double s = block_tree(acc, sh);
    if (threadIdx.x == 0) part[c] = s;
  }
  if (last_block(&st->ticket[1], &flag)) {
    const double s = final_sum(part, nchunks, sh);
    if (threadIdx.x == 0) {
      st->loc[0] = s; st->loc[1] = 0.0; st->loc[2] = 0.0; st->loc[3] = 0.0;
      sh_loc[0] = s; sh_loc[1] = 0.0; sh_loc[2] = 0.0; sh_loc[3] = 0.0;
      if (single_rank && mode == 1) finish_pu(st, s);
    }
    if (T && mode == 1) {
      __syncthreads();
      peer_allreduce(T, 1, sh_loc, sh_out, st);
      if (threadIdx.x == 0) finish_pu(st, sh_out[0]);
    }
  }
}

// xnew = x + p*alpha; r = r - u*alpha; d = diag*r; partial r.d, max|xnew|, max|xnew-x|; x = xnew
// (p121.f90:100-102 and checon_par maths.f90:1048-1061)
__global__ void __launch_bounds__(kRedThreads)
k_pcg_update(const double *__restrict__ diag, const double *__restrict__ p, const double *__restrict__ u,
             double *__restrict__ x, double *__restrict__ r, double *__restrict__ d, long long n,
             double *part /*3*nchunks*/, State *st, int single_rank, double *ratio_hist, PeerTable *T) {
  if (*(volatile const int *)&st->done) return;
  __shared__ double sh[8];
  __shared__ double sh3[2][24];
  __shared__ double sh_loc[4], sh_out[3];
  __shared__ int flag;
  int nloc = 0;
  const double alpha = st->alpha;
  const long long nchunks = (n + kChunk - 1) / kChunk;
  for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
    double acc = 0.0, ml = 0.0, md = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long i = c * kChunk + 512 * k + 2 * threadIdx.x;
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (i + h < n) {
          const double xo = x[i + h];
          const double xn = xo + p[i + h] * alpha;
          const double rr = r[i + h] - u[i + h] * alpha;
          const double dd = diag[i + h] * rr;
          x[i + h] = xn; r[i + h] = rr; d[i + h] = dd;
          acc = acc + rr * dd;
          ml = fmax(ml, fabs(xn));
          md = fmax(md, fabs(xn - xo));
        }
    }
    // the three block reductions of a chunk share one barrier (double-buffered by chunk parity);
    // same trees as block_tree / block_max: warp xor-tree, then warps 0..7 in order
    const double ws = warp_tree_sum(acc), wm1 = warp_tree_max(ml), wm2 = warp_tree_max(md);
    double *buf = sh3[nloc & 1];
    if ((threadIdx.x & 31) == 0) {
      const int w = threadIdx.x >> 5;
      buf[w] = ws; buf[8 + w] = wm1; buf[16 + w] = wm2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = buf[0], m1 = buf[8], m2 = buf[16];
#pragma unroll
      for (int w = 1; w < 8; ++w) { s = s + buf[w]; m1 = fmax(m1, buf[8 + w]); m2 = fmax(m2, buf[16 + w]); }
      part[c] = s; part[nchunks + c] = m1; part[2 * nchunks + c] = m2;
    }
    ++nloc;
  }
  if (last_block(&st->ticket[2], &flag)) {
    const double s = final_sum(part, nchunks, sh);
    const double m1 = final_max(part + nchunks, nchunks, sh);
    const double m2 = final_max(part + 2 * nchunks, nchunks, sh);
    if (threadIdx.x == 0) {
      st->loc[0] = s; st->loc[1] = m1; st->loc[2] = m2; st->loc[3] = 0.0;
      sh_loc[0] = s; sh_loc[1] = m1; sh_loc[2] = m2; sh_loc[3] = 0.0;
      if (single_rank) finish_update(st, s, m1, m2, ratio_hist);
    }
    if (T) {
      __syncthreads();
      peer_allreduce(T, 2, sh_loc, sh_out, st);
      if (threadIdx.x == 0) finish_update(st, sh_out[0], sh_out[1], sh_out[2], ratio_hist);
    }
  }
}

// forward halo exchange folded into the p update (N ranks, peer transport): which owned equations are wanted by
// peers (one bit each), and where each goes
struct PutTables {
  const unsigned int *bits;        // bit i of word i/32: owned equation i (0-based) is wanted by at least one peer
  const int *slot0;                // those equations, ascending (0-based)
  const unsigned int *ptr;         // per such equation: its destinations
  const int *rank;                 // destination rank
  const long long *dst;            // index in that rank's p_ext
  int n;                           // number of such equations
};
// p = d + p*beta (p121.f90:102), then the exit test of p121.f90:103.  `done` is raised by the last
// block to finish, so no block of this kernel can observe it early and every later kernel sees it.
// T != nullptr: every new p value a peer's elements need is stored into that peer's halo segment as it is formed,
// and the last block releases the forward-halo flags: the gather's exchange of the NEXT iteration
// (gather_scatter.f90:547-688) costs no kernel of its own.
__global__ void k_pupdate(const double *__restrict__ d, double *__restrict__ p, long long n, State *st,
                          PeerTable *T, PutTables P) {
  if (*(volatile const int *)&st->done) return;
  __shared__ int flag;
  const double beta = st->beta;
  if (T) {
    // first the equations peers gather, each exactly once (the sweep below skips them): the stores cross NVLink
    // while the rest of the vector is updated
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < P.n; j += gridDim.x * blockDim.x) {
      const int i = P.slot0[j];
      const double v = d[i] + p[i] * beta;
      p[i] = v;
      for (unsigned int q = P.ptr[j]; q < P.ptr[j + 1]; ++q) T->p_ext[P.rank[q]][P.dst[q]] = v;
    }
  }
  long long i = 2 * ((long long)blockIdx.x * blockDim.x + threadIdx.x);
  const long long stride = 2 * (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const unsigned int w = T ? P.bits[i >> 5] >> (i & 31) : 0u;      // i is even: both bits sit in the same word
    if (!(w & 1u)) p[i] = d[i] + p[i] * beta;
    if (i + 1 < n && !(w & 2u)) p[i + 1] = d[i + 1] + p[i + 1] * beta;
  }
  if (T) peer_release_flags(T, 0);
  if (last_block(&st->ticket[3], &flag) && threadIdx.x == 0) {
    if (st->converged || st->iters == st->limit) st->done = 1;
  }
}

// diag = 1/diag (+ penalty on fixed equations first; p123.f90:120-125, p121.f90:86)
__global__ void k_fixed_penalty(const int *__restrict__ fix_slot, double *__restrict__ diag_ext,
                                double *__restrict__ store, double penalty, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double v = diag_ext[fix_slot[i]] + penalty;
    diag_ext[fix_slot[i]] = v;
    store[i] = v;
  }
}
__global__ void k_invert(double *__restrict__ v, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) v[i] = 1.0 / v[i];
}

// ----------------------------------------------------------------------------
// a3-a5: element matrices.  One CTA per element, Gauss points in sequence.
// Tables (shape-function derivatives at the Gauss points, weights, dee) are
// computed on the host with the reference's formulas and held in constant memory.
// ----------------------------------------------------------------------------
struct ElemTables {
  double der[27 * 3 * 20]; // [ig][a][m]  = der(a,m) at Gauss point ig   (shape_der); up to the 27-point rule (p129)
  double weights[27];      // sample
  double dee[36];          // dee(l,k) at [k*6+l]                        (deemat)
  double kxyz[3];          // p123 / p124 conductivities (diagonal of kay)
  double fun[8 * 8];       // [ig][m] = fun(m) at Gauss point ig, 8-node brick   (shape_fun)
  double fun20[27 * 20];   // the same for the 20-node brick (new_library.f90:449-468), p129's consistent mass
  double trans[4];         // p124: rho, cp, theta, dtim
  int nip;
};
__constant__ ElemTables c_tab;  // single translation unit (device.cu)

// 3x3 determinant and adjugate inverse with the reference's operation order
// (maths.f90:617-619, 526-547)
__device__ __forceinline__ double det3(const double *m /*col-major*/) {
#define A(r, c) m[(c - 1) * 3 + (r - 1)]
  double det = A(1, 1) * (A(2, 2) * A(3, 3) - A(3, 2) * A(2, 3));
  det = det - A(1, 2) * (A(2, 1) * A(3, 3) - A(3, 1) * A(2, 3));
  det = det + A(1, 3) * (A(2, 1) * A(3, 2) - A(3, 1) * A(2, 2));
  return det;
}
__device__ __forceinline__ void inv3(const double *m, double det, double *o) {
  o[0] = (A(2, 2) * A(3, 3) - A(3, 2) * A(2, 3)) / det;      // (1,1)
  o[1] = (-(A(2, 1) * A(3, 3)) + A(3, 1) * A(2, 3)) / det;   // (2,1)
  o[2] = (A(2, 1) * A(3, 2) - A(3, 1) * A(2, 2)) / det;      // (3,1)
  o[3] = (-(A(1, 2) * A(3, 3)) + A(3, 2) * A(1, 3)) / det;   // (1,2)
  o[4] = (A(1, 1) * A(3, 3) - A(3, 1) * A(1, 3)) / det;      // (2,2)
  o[5] = (-(A(1, 1) * A(3, 2)) + A(3, 1) * A(1, 2)) / det;   // (3,2)
  o[6] = (A(1, 2) * A(2, 3) - A(2, 2) * A(1, 3)) / det;      // (1,3)
  o[7] = (-(A(1, 1) * A(2, 3)) + A(2, 1) * A(1, 3)) / det;   // (2,3)
  o[8] = (A(1, 1) * A(2, 2) - A(2, 1) * A(1, 2)) / det;      // (3,3)
}
// the tensor-core matrix-free kernels (order 2 of the oracle): ONE IEEE division per Gauss point instead of nine -- the
// adjugate entries with the operation order above, times 1/det
__device__ __forceinline__ void inv3_recip(const double *m, double det, double *o) {
  const double rd = 1.0 / det;
  o[0] = (A(2, 2) * A(3, 3) - A(3, 2) * A(2, 3)) * rd;
  o[1] = (-(A(2, 1) * A(3, 3)) + A(3, 1) * A(2, 3)) * rd;
  o[2] = (A(2, 1) * A(3, 2) - A(3, 1) * A(2, 2)) * rd;
  o[3] = (-(A(1, 2) * A(3, 3)) + A(3, 2) * A(1, 3)) * rd;
  o[4] = (A(1, 1) * A(3, 3) - A(3, 1) * A(1, 3)) * rd;
  o[5] = (-(A(1, 1) * A(3, 2)) + A(3, 1) * A(1, 2)) * rd;
  o[6] = (A(1, 2) * A(2, 3) - A(2, 2) * A(1, 3)) * rd;
  o[7] = (-(A(1, 1) * A(2, 3)) + A(2, 1) * A(1, 3)) * rd;
  o[8] = (A(1, 1) * A(2, 2) - A(2, 1) * A(1, 2)) * rd;
#undef A
}

// jac = der*coord, det, deriv = jac^-1 * der for Gauss point ig; coord(nod,3) in smem.
// Leaves deriv(a,m) at s_deriv[m*3+a]; returns det in every thread.
template <int NOD>
__device__ __forceinline__ double gauss_point(int ig, const double *s_coord, double *s_jac, double *s_deriv) {
  const double *der = c_tab.der + ig * 3 * 20;
  if (threadIdx.x < 9) {
    const int a = threadIdx.x % 3, b = threadIdx.x / 3;
    double s = 0.0;
#pragma unroll
    for (int m = 0; m < NOD; ++m) s = s + der[a * 20 + m] * s_coord[b * NOD + m];
    s_jac[b * 3 + a] = s;
  }
  __syncthreads();
  double jac[9], inv[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) jac[q] = s_jac[q];
  const double det = det3(jac);
  inv3(jac, det, inv);
  if (threadIdx.x < 3 * NOD) {
    const int m = threadIdx.x / 3, a = threadIdx.x - 3 * m;
    double s = 0.0;
#pragma unroll
    for (int b = 0; b < 3; ++b) s = s + inv[b * 3 + a] * der[b * 20 + m];
    s_deriv[m * 3 + a] = s;
  }
  __syncthreads();
  return det;
}

// elements_1 of p121.f90:56-64.  km(i,j) += (sum_k btd(i,k)*bee(k,j)) * det * w
// MAT: one dee per material (xx2.f90:176-180: e, v = prop(:,etype_pp(iel))), dee_tab (36, np_types) in global
// memory and etype 1-based; otherwise the single dee of p121 in constant memory.
template <int NOD, int THREADS, bool MAT = false>
__global__ void __launch_bounds__(THREADS)
k_form_km_elastic(const double *__restrict__ g_coord, double *__restrict__ km, long long nels,
                  double *__restrict__ diag_only /* matrix-free: (ntot,nels) diagonal instead of km */,
                  int packed /* 1: lower triangle only, packed by columns (SymCfg) */,
                  const double *__restrict__ dee_tab = nullptr, const int *__restrict__ etype = nullptr) {
  constexpr int NTOT = 3 * NOD, NENT = NTOT * NTOT, PER = (NENT + THREADS - 1) / THREADS;
  __shared__ double s_coord[NOD * 3], s_jac[9], s_deriv[NOD * 3];
  __shared__ double s_bee[6 * NTOT];  // bee(l,c) at [c*6+l]
  __shared__ double s_btd[6 * NTOT];  // btd(i,k) at [k*NTOT+i]
  __shared__ double s_dee[MAT ? 36 : 1];
  for (long long e = blockIdx.x; e < nels; e += gridDim.x) {
    __syncthreads();
    for (int q = threadIdx.x; q < NOD * 3; q += THREADS) s_coord[q] = g_coord[e * NOD * 3 + q];
    if constexpr (MAT) {
      if (threadIdx.x < 36) s_dee[threadIdx.x] = dee_tab[(long long)(etype[e] - 1) * 36 + threadIdx.x];
    }
    double acc[PER];
#pragma unroll
    for (int n = 0; n < PER; ++n) acc[n] = 0.0;
    __syncthreads();
    for (int ig = 0; ig < c_tab.nip; ++ig) {
      const double det = gauss_point<NOD>(ig, s_coord, s_jac, s_deriv);
      const double wt = c_tab.weights[ig];
      // beemat (new_library.f90:976-993)
      for (int c = threadIdx.x; c < NTOT; c += THREADS) {
        const int m = c / 3, comp = c - 3 * m;
        const double x = s_deriv[m * 3 + 0], y = s_deriv[m * 3 + 1], z = s_deriv[m * 3 + 2];
        double b0 = 0, b1 = 0, b2 = 0, b3 = 0, b4 = 0, b5 = 0;
        if (comp == 0) { b0 = x; b3 = y; b5 = z; }
        else if (comp == 1) { b1 = y; b3 = x; b4 = z; }
        else { b2 = z; b4 = y; b5 = x; }
        double *bc = s_bee + c * 6;
        bc[0] = b0; bc[1] = b1; bc[2] = b2; bc[3] = b3; bc[4] = b4; bc[5] = b5;
      }
      __syncthreads();
      // btd = MATMUL(TRANSPOSE(bee),dee)
      for (int q = threadIdx.x; q < 6 * NTOT; q += THREADS) {
        const int k = q / NTOT, i = q - k * NTOT;
        double s = 0.0;
#pragma unroll
        for (int l = 0; l < 6; ++l) {
          if constexpr (MAT) s = s + s_bee[i * 6 + l] * s_dee[k * 6 + l];
          else s = s + s_bee[i * 6 + l] * c_tab.dee[k * 6 + l];
        }
        s_btd[k * NTOT + i] = s;
      }
      __syncthreads();
      if (diag_only) {
        // matrix-free setup: only km(i,i) is needed -- same expression, 1/NTOT of the work
#pragma unroll
        for (int m = 0; m < (NTOT + THREADS - 1) / THREADS; ++m) {
          const int i = threadIdx.x + m * THREADS;
          if (i < NTOT) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) s = s + s_btd[k * NTOT + i] * s_bee[i * 6 + k];
            acc[m] = acc[m] + s * det * wt;
          }
        }
      } else {
#pragma unroll
        for (int n = 0; n < PER; ++n) {
          const int idx = threadIdx.x + n * THREADS;
          if (idx < NENT) {
            const int j = idx / NTOT, i = idx - j * NTOT;
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) s = s + s_btd[k * NTOT + i] * s_bee[j * 6 + k];
            acc[n] = acc[n] + s * det * wt;
          }
        }
      }
      __syncthreads();
    }
    if (diag_only) {
#pragma unroll
      for (int m = 0; m < (NTOT + THREADS - 1) / THREADS; ++m) {
        const int i = threadIdx.x + m * THREADS;
        if (i < NTOT) diag_only[e * (long long)NTOT + i] = acc[m];
      }
      continue;
    }
#pragma unroll
    for (int n = 0; n < PER; ++n) {
      const int idx = threadIdx.x + n * THREADS;
      if (idx < NENT) {
        if (packed) {
          const int j = idx / NTOT, i = idx - j * NTOT;
          if (i >= j) km[e * (long long)SymCfg<NTOT>::kPacked + SymCfg<NTOT>::coloff(j) + (i - j)] = acc[n];
        } else {
          km[e * (long long)NENT + idx] = acc[n];
        }
      }
    }
  }
}


// elements_1 of p121.f90:56-64 again, register-tiled: the version that forms the whole storkm_pp.
// One CTA per element.  (1) jac / det / inverse / deriv of ALL Gauss points at once (no barrier per point);
// (2) thread (ta,tb) owns the TA x TB node block of km = 3TA x 3TB entries in registers and walks the
// Gauss points in order: btd = MATMUL(TRANSPOSE(bee),dee) and km += (btd*bee)*det*w are evaluated from the
// node derivatives directly, with exactly the reference's products in the reference's order (l, k ascending)
// MINUS the products that have a structural zero of bee or dee as a factor -- those contribute +-0.0 and
// can only change the sign of a zero; (3) the element matrix is staged in shared memory and written out in
// 128-bit coalesced stores (or packed).  21 products per 3x3 node block and point instead of 54, and 0.1
// shared loads per product instead of 2: the first build (k_form_km_elastic above, kept for the matrix-free
// diagonal) was bound by its shared-memory loads at ~19 % of the FP64 pipe.
// DIAG (matrix-free setup): only the diagonal of km is wanted, km receives (ntot,nels); thread a < NOD evaluates
// the diagonal entries of node block (a,a) with the same products in the same order.
template <int NOD, int TA, int TB, int THREADS, bool MAT, bool DIAG = false, int MINB = 1>
__global__ void __launch_bounds__(THREADS, MINB)
k_form_km_tiled(const double *__restrict__ g_coord, double *__restrict__ km, long long nels, int packed,
                const double *__restrict__ dee_tab, const int *__restrict__ etype) {
  constexpr int NTOT = 3 * NOD, NENT = NTOT * NTOT, NA = NOD / TA, NB = NOD / TB, MAXIP = 8;
  static_assert(NOD % TA == 0 && NOD % TB == 0 && NA * NB <= THREADS && NENT % 2 == 0, "tile shape");
  __shared__ double s_coord[NOD * 3], s_jac[MAXIP * 9], s_det[MAXIP], s_dee[36];
  __shared__ double s_deriv[MAXIP * NOD * 3];            // [ig][m*3+a] = deriv(a,m) at Gauss point ig
  __shared__ __align__(16) double s_km[NENT];
  const int nip = c_tab.nip, t = threadIdx.x;
  const int ta = t % NA, tb = t / NA;
  for (long long e = blockIdx.x; e < nels; e += gridDim.x) {
    __syncthreads();                                      // the previous element's s_km has been written out
    for (int q = t; q < NOD * 3; q += THREADS) s_coord[q] = g_coord[e * NOD * 3 + q];
    if (t < 36) s_dee[t] = MAT ? dee_tab[(long long)(etype[e] - 1) * 36 + t] : c_tab.dee[t];
    __syncthreads();
    for (int q = t; q < nip * 9; q += THREADS) {          // jac = MATMUL(der,coord), every point
      const int ig = q / 9, r = q - 9 * ig, a = r % 3, b = r / 3;
      const double *der = c_tab.der + ig * 60;
      double sum = 0.0;
#pragma unroll
      for (int m = 0; m < NOD; ++m) sum = sum + der[a * 20 + m] * s_coord[b * NOD + m];
      s_jac[ig * 9 + b * 3 + a] = sum;
    }
    __syncthreads();
    if (t < nip) {                                        // determinant, invert
      double jac[9], inv[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) jac[q] = s_jac[t * 9 + q];
      const double det = det3(jac);
      inv3(jac, det, inv);
#pragma unroll
      for (int q = 0; q < 9; ++q) s_jac[t * 9 + q] = inv[q];
      s_det[t] = det;
    }
    __syncthreads();
    for (int q = t; q < nip * NOD * 3; q += THREADS) {    // deriv = MATMUL(jac^-1,der)
      const int ig = q / (NOD * 3), r = q - ig * (NOD * 3), m = r / 3, a = r - 3 * m;
      const double *der = c_tab.der + ig * 60, *inv = s_jac + ig * 9;
      double sum = 0.0;
#pragma unroll
      for (int b = 0; b < 3; ++b) sum = sum + inv[b * 3 + a] * der[b * 20 + m];
      s_deriv[q] = sum;
    }
    __syncthreads();
    if constexpr (DIAG) {
      if (t < NOD) {
        const double d00 = s_dee[0], d11 = s_dee[7], d22 = s_dee[14], d33 = s_dee[21], d44 = s_dee[28], d55 = s_dee[35];
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        for (int ig = 0; ig < nip; ++ig) {
          const double det = s_det[ig], wt = c_tab.weights[ig];
          const double *dv = s_deriv + ig * (NOD * 3);
          const double x = dv[t * 3], y = dv[t * 3 + 1], z = dv[t * 3 + 2];
          double s;
          s = (x * d00) * x; s = s + (y * d33) * y; s = s + (z * d55) * z; a0 = a0 + s * det * wt;   // (0,0): k = 0,3,5
          s = (y * d11) * y; s = s + (x * d33) * x; s = s + (z * d44) * z; a1 = a1 + s * det * wt;   // (1,1): k = 1,3,4
          s = (z * d22) * z; s = s + (y * d44) * y; s = s + (x * d55) * x; a2 = a2 + s * det * wt;   // (2,2): k = 2,4,5
        }
        km[e * (long long)NTOT + 3 * t] = a0;
        km[e * (long long)NTOT + 3 * t + 1] = a1;
        km[e * (long long)NTOT + 3 * t + 2] = a2;
      }
      continue;
    }
    if (tb < NB) {
      // D(l,k) = dee(l,k); only the entries an isotropic dee holds
      const double d00 = s_dee[0], d10 = s_dee[1], d20 = s_dee[2];        // dee(:,0) at [0*6+l]
      const double d01 = s_dee[6], d11 = s_dee[7], d21 = s_dee[8];
      const double d02 = s_dee[12], d12 = s_dee[13], d22 = s_dee[14];
      const double d33 = s_dee[21], d44 = s_dee[28], d55 = s_dee[35];
      double acc[TA][TB][9];
#pragma unroll
      for (int ia = 0; ia < TA; ++ia)
#pragma unroll
        for (int ib = 0; ib < TB; ++ib)
#pragma unroll
          for (int q = 0; q < 9; ++q) acc[ia][ib][q] = 0.0;
      for (int ig = 0; ig < nip; ++ig) {
        const double det = s_det[ig], wt = c_tab.weights[ig];
        const double *dv = s_deriv + ig * (NOD * 3);
        double xb[TB], yb[TB], zb[TB];
#pragma unroll
        for (int ib = 0; ib < TB; ++ib) {
          const int b = tb * TB + ib;
          xb[ib] = dv[b * 3]; yb[ib] = dv[b * 3 + 1]; zb[ib] = dv[b * 3 + 2];
        }
#pragma unroll
        for (int ia = 0; ia < TA; ++ia) {
          const int a = ta * TA + ia;
          const double x = dv[a * 3], y = dv[a * 3 + 1], z = dv[a * 3 + 2];
          // btd(3a+p,k) = sum_l bee(l,3a+p)*dee(l,k): one surviving product each
          const double b00 = x * d00, b01 = x * d01, b02 = x * d02, b03 = y * d33, b05 = z * d55;
          const double b10 = y * d10, b11 = y * d11, b12 = y * d12, b13 = x * d33, b14 = z * d44;
          const double b20 = z * d20, b21 = z * d21, b22 = z * d22, b24 = y * d44, b25 = x * d55;
#pragma unroll
          for (int ib = 0; ib < TB; ++ib) {
            const double X = xb[ib], Y = yb[ib], Z = zb[ib];
            double *A = acc[ia][ib];                     // A[q*3+p] = km(3a+p, 3b+q)
            double s;
            s = b00 * X; s = s + b03 * Y; s = s + b05 * Z; A[0] = A[0] + s * det * wt;   // (0,0): k = 0,3,5
            s = b10 * X; s = s + b13 * Y;                  A[1] = A[1] + s * det * wt;   // (1,0): k = 0,3
            s = b20 * X; s = s + b25 * Z;                  A[2] = A[2] + s * det * wt;   // (2,0): k = 0,5
            s = b01 * Y; s = s + b03 * X;                  A[3] = A[3] + s * det * wt;   // (0,1): k = 1,3
            s = b11 * Y; s = s + b13 * X; s = s + b14 * Z; A[4] = A[4] + s * det * wt;   // (1,1): k = 1,3,4
            s = b21 * Y; s = s + b24 * Z;                  A[5] = A[5] + s * det * wt;   // (2,1): k = 1,4
            s = b02 * Z; s = s + b05 * X;                  A[6] = A[6] + s * det * wt;   // (0,2): k = 2,5
            s = b12 * Z; s = s + b14 * Y;                  A[7] = A[7] + s * det * wt;   // (1,2): k = 2,4
            s = b22 * Z; s = s + b24 * Y; s = s + b25 * X; A[8] = A[8] + s * det * wt;   // (2,2): k = 2,4,5
          }
        }
      }
#pragma unroll
      for (int ia = 0; ia < TA; ++ia)
#pragma unroll
        for (int ib = 0; ib < TB; ++ib)
#pragma unroll
          for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int pp = 0; pp < 3; ++pp)
              s_km[(3 * (tb * TB + ib) + q) * NTOT + 3 * (ta * TA + ia) + pp] = acc[ia][ib][q * 3 + pp];
    }
    __syncthreads();
    if (!packed) {
      double2 *dst = reinterpret_cast<double2 *>(km + e * (long long)NENT);
      const double2 *src = reinterpret_cast<const double2 *>(s_km);
      for (int q = t; q < NENT / 2; q += THREADS) dst[q] = src[q];
    } else {
      double *dst = km + e * (long long)SymCfg<NTOT>::kPacked;
      for (int q = t; q < NENT; q += THREADS) {
        const int j = q / NTOT, i = q - j * NTOT;
        if (i >= j) dst[SymCfg<NTOT>::coloff(j) + (i - j)] = s_km[q];
      }
    }
  }
}


// ----------------------------------------------------------------------------
// BASELINE config E: matrix-free element products, utemp(:,e) = sum_gp B^T (D (B p)) det w
// ----------------------------------------------------------------------------
// One THREAD per element, a warp owns 32 consecutive elements.  With the Cartesian derivatives
// written as deriv = jac^-1 * der, the operator factors through two 3x3 matrices per Gauss point:
//   H(b,c)  = sum_m der(b,m) p_c(m)                    (phase 1, 9 fma per node and point)
//   G = jac^-1 H  ->  eps  ->  sigma = D eps det w  ->  T = jac^-T S(sigma)      (~100 flop per point)
//   u_c(m)  = sum_gp sum_b der(b,m) T_gp(b,c)          (phase 2, 9 fma per node and point)
// i.e. 18 fma per node and point instead of the 36 of the B-matrix form, and the sum over the Gauss
// points happens inside the phase-2 chains.  der is the same for every element, so with the loops
// fully unrolled every der value is an immediate constant-bank operand of its DFMA: the FP64 pipe
// sees almost nothing but DFMAs, shared memory only carries the thread's own row of right-hand
// sides / results (staged by the warp with coalesced global accesses), and the 72 T values live in
// registers.  The Gauss points are processed two at a time so that at most 2x9 H accumulators and
// the two points' geometric factors are live next to T.  Same operation order as orc_apply_mf.
// FP64-pipe bound (7 020 flop per hex20 element with stored factors), not HBM bound.
template <int NOD>
struct MfCfg {
  static constexpr int NTOT = 3 * NOD;
  // per-lane rows are read/written with 128-bit accesses: a stride of NTOT+2 doubles (496 / 208 B)
  // keeps them 16-byte aligned and puts the 8 lanes of a quarter-warp on distinct 16-byte bank groups
  static constexpr int kRow = NTOT + 2;
  static constexpr int kGeom = 80;                        // doubles per element: 8 x (jac^-1 (9), det*w)
  static constexpr int kIdxBytes = 32 * NTOT * 4;         // one group's gather indices (bulk-copied)
  static constexpr int kDerBytes = NOD * 24 * 8;          // der(b,m) of the 8 points as [m][g][b]
  static constexpr size_t smem(int warps) { return (size_t)kDerBytes + (size_t)warps * (32 * kRow * 8 + kIdxBytes + 16); }
  // geometric factors of a group of 32 elements: [g][j][lane] double2, j = 0..4 -- lanes read
  // consecutive 16-byte words (coalesced); 2560 doubles per group
  static constexpr long long kGroupGeom = 32LL * kGeom;
};

__device__ __forceinline__ void lds128(uint32_t addr, double &x, double &y) {
  asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
}
__device__ __forceinline__ void sts128(uint32_t addr, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}

// H(b,c) at [b*3+c], inv = jac^-1 with (a,b) at [b*3+a], f = det*w  ->  T(b,c) at [b*3+c]
__device__ __forceinline__ void mf_point_mid(const double *H, const double *inv, double f, double *T) {
  double G[9];  // G(a,c) at [a*3+c]: displacement gradient d u_c / d x_a
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double s = inv[a] * H[c];
      s = fma(inv[3 + a], H[3 + c], s);
      s = fma(inv[6 + a], H[6 + c], s);
      G[a * 3 + c] = s;
    }
  // strains in beemat's row order (new_library.f90:976-993)
  const double eps[6] = {G[0], G[4], G[8], G[3] + G[1], G[7] + G[5], G[6] + G[2]};
  double sig[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    double s = c_tab.dee[r] * eps[0];
#pragma unroll
    for (int c = 1; c < 6; ++c) s = fma(c_tab.dee[c * 6 + r], eps[c], s);
    sig[r] = s * f;
  }
  const double S[9] = {sig[0], sig[3], sig[5], sig[3], sig[1], sig[4], sig[5], sig[4], sig[2]};  // S(a,c) at [a*3+c]
#pragma unroll
  for (int b = 0; b < 3; ++b)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double s = inv[b * 3] * S[c];
      s = fma(inv[b * 3 + 1], S[3 + c], s);
      s = fma(inv[b * 3 + 2], S[6 + c], s);
      T[b * 3 + c] = s;
    }
}

// The same with deemat's structural zeros left out (isotropic dee, new_library.f90:905-930: a 3x3 block and a diagonal):
// the normal stresses are 3-term chains, the shear stresses single products -- 18 instead of 42 FP64 instructions.
// The left-out terms are products with an exact 0.0.  k_apply_mf3 and the oracle's order 2 use this form.
__device__ __forceinline__ void mf_point_mid_iso(const double *H, const double *inv, double f, double *T) {
  double G[9];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double s = inv[a] * H[c];
      s = fma(inv[3 + a], H[3 + c], s);
      s = fma(inv[6 + a], H[6 + c], s);
      G[a * 3 + c] = s;
    }
  const double eps[6] = {G[0], G[4], G[8], G[3] + G[1], G[7] + G[5], G[6] + G[2]};
  double sig[6];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double s = c_tab.dee[r] * eps[0];
    s = fma(c_tab.dee[6 + r], eps[1], s);
    s = fma(c_tab.dee[12 + r], eps[2], s);
    sig[r] = s * f;
  }
#pragma unroll
  for (int r = 3; r < 6; ++r) sig[r] = (c_tab.dee[r * 6 + r] * eps[r]) * f;
  const double S[9] = {sig[0], sig[3], sig[5], sig[3], sig[1], sig[4], sig[5], sig[4], sig[2]};
#pragma unroll
  for (int b = 0; b < 3; ++b)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double s = inv[b * 3] * S[c];
      s = fma(inv[b * 3 + 1], S[3 + c], s);
      s = fma(inv[b * 3 + 2], S[6 + c], s);
      T[b * 3 + c] = s;
    }
}

// GEOM 0: rebuild jac / inverse / det at every point from the coordinates every call (config E as
//         named); the 80 factors of an element pass through a per-thread scratch line in `geom`
//         (L2-resident, never re-read by another thread);
// GEOM 1: only write the factors of every element to `geom` (setup of mode 2);
// GEOM 2: read them (640 B per hex element instead of 28 800 B of storkm): no Jacobian pass, no
//         FP64 divisions.  Modes 0 and 2 produce identical bits.
// Data movement per group of 32 elements: the NEXT group's gather indices arrive by one 1-D bulk
// async copy (mbarrier) and its geometric factors are prefetched into L2 while the current group is
// computed; the gather itself keeps one load per dof of the whole group in flight per lane; der comes
// from a 3.8 KB shared-memory table as warp-uniform (broadcast) 128-bit loads; the factors of the
// next Gauss point are loaded while the current one is finished.  The Gauss points are accumulated
// four at a time, so a row of right-hand sides is read from shared memory only twice.
template <int NOD, bool GATHER, int GEOM, int WARPS, int UNR = 2>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_apply_mf(const double *__restrict__ g_coord, const int *__restrict__ ggl, const double *__restrict__ pvec,
           double *__restrict__ utemp, long long nels, const State *st, double *geom, PeerTable *T) {
  using Cfg = MfCfg<NOD>;
  constexpr int NTOT = Cfg::NTOT, ROW = Cfg::kRow, KP = (NTOT + 31) / 32;
  static_assert(NOD % 2 == 0, "nodes are processed in pairs");
  if (st && *(volatile const int *)&st->done) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *s_der = reinterpret_cast<double *>(smem_raw);                                   // [m][g][b]
  double *rows = reinterpret_cast<double *>(smem_raw + Cfg::kDerBytes) + (size_t)w * 32 * ROW;
  int *idxbuf = reinterpret_cast<int *>(smem_raw + Cfg::kDerBytes + (size_t)WARPS * 32 * ROW * 8) + (size_t)w * 32 * NTOT;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + Cfg::kDerBytes + (size_t)WARPS * (32 * ROW * 8 + Cfg::kIdxBytes));
  const uint32_t bar = smem_u32(&bars[w]);
  const uint32_t row_s = smem_u32(rows + lane * ROW);       // this lane's element row
  for (int q = threadIdx.x; q < NOD * 24; q += blockDim.x) {
    const int m = q / 24, r = q - m * 24, g = r / 3, b = r - g * 3;
    s_der[q] = c_tab.der[g * 60 + b * 20 + m];
  }
  __syncthreads();
  const long long ngroups = (nels + 31) / 32;
  const long long gstride = (long long)gridDim.x * WARPS;
  uint64_t policy = 0;
  uint32_t phase = 0;
  auto issue_idx = [&](long long g) {                      // lane 0 only
    const long long e0 = g * 32;
    const int ne = (int)((nels - e0) < 32 ? (nels - e0) : 32);
    const uint32_t bytes = (uint32_t)ne * NTOT * 4;
    mbar_expect_tx(bar, bytes);
    bulk_g2s(smem_u32(idxbuf), ggl + e0 * NTOT, bytes, bar, policy);
  };
  long long grp = (long long)blockIdx.x * WARPS + w;
  if (GATHER && GEOM != 1) {
    if (lane == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
      policy = policy_evict_first();
      if (grp < ngroups) issue_idx(grp);
    }
    __syncwarp();
    // N ranks, peer transport: the owners' values are in my halo segment of pvec.  The gather below uses ld.global.ca
    // (not the non-coherent path): nothing of pvec was loaded by this kernel before the flags were seen.
    if (T) warp_wait_fwd(T, st);
  }
  for (; grp < ngroups; grp += gstride) {
    const long long e0 = grp * 32;
    const int ne = (int)((nels - e0) < 32 ? (nels - e0) : 32);
    // this lane's 16-byte words of the group's factors: word (g*5+j) at gfl[(g*5+j)*32]
    double2 *gfl = reinterpret_cast<double2 *>(geom + (GEOM == 0 ? (long long)blockIdx.x * WARPS + w : grp) * Cfg::kGroupGeom) + lane;
    if (GEOM == 2 && lane == 0 && grp + gstride < ngroups) {
      // next group's factors (20 KB) towards L2 while this group is computed
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(geom + (grp + gstride) * Cfg::kGroupGeom),
                   "r"((uint32_t)(Cfg::kGroupGeom * 8)) : "memory");
    }
    if (GEOM != 2) {
      // coordinates g_coord_pp(nod,3,nels) -> the element's row, [b*NOD+m]
      if (ne == 32) {
        double val[32][KP];
#pragma unroll
        for (int el = 0; el < 32; ++el)
#pragma unroll
          for (int kp = 0; kp < KP; ++kp) {
            const int k = lane + 32 * kp;
            val[el][kp] = (k < NTOT) ? g_coord[(e0 + el) * NTOT + k] : 0.0;
          }
#pragma unroll
        for (int el = 0; el < 32; ++el)
#pragma unroll
          for (int kp = 0; kp < KP; ++kp) {
            const int k = lane + 32 * kp;
            if (k < NTOT) rows[el * ROW + k] = val[el][kp];
          }
      } else {
        for (int el = 0; el < ne; ++el)
          for (int k = lane; k < NTOT; k += 32) rows[el * ROW + k] = g_coord[(e0 + el) * NTOT + k];
      }
      __syncwarp();
      if (lane < ne) {
#pragma unroll
        for (int Q = 0; Q < 2; ++Q) {
          double J[4][9];
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int z = 0; z < 9; ++z) J[g][z] = 0.0;
#pragma unroll UNR
          for (int t = 0; t < NOD / 2; ++t) {
            // [b*NOD+m]: x of nodes 2t,2t+1 / y / z
            double cx[2], cy[2], cz[2];
            lds128(row_s + (2 * t) * 8, cx[0], cx[1]);
            lds128(row_s + (NOD + 2 * t) * 8, cy[0], cy[1]);
            lds128(row_s + (2 * NOD + 2 * t) * 8, cz[0], cz[1]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const double *der = s_der + (2 * t + h) * 24 + Q * 12;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const double d0 = der[3 * g], d1 = der[3 * g + 1], d2 = der[3 * g + 2];
                // jac(a,b) at [b*3+a], node-ascending fma chains
                J[g][0] = fma(d0, cx[h], J[g][0]); J[g][1] = fma(d1, cx[h], J[g][1]); J[g][2] = fma(d2, cx[h], J[g][2]);
                J[g][3] = fma(d0, cy[h], J[g][3]); J[g][4] = fma(d1, cy[h], J[g][4]); J[g][5] = fma(d2, cy[h], J[g][5]);
                J[g][6] = fma(d0, cz[h], J[g][6]); J[g][7] = fma(d1, cz[h], J[g][7]); J[g][8] = fma(d2, cz[h], J[g][8]);
              }
            }
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            double inv[9];
            const double det = det3(J[g]);
            inv3(J[g], det, inv);
            const double f = det * c_tab.weights[4 * Q + g];
            double2 *o = gfl + (4 * Q + g) * 5 * 32;
            o[0] = make_double2(inv[0], inv[1]); o[32] = make_double2(inv[2], inv[3]);
            o[64] = make_double2(inv[4], inv[5]); o[96] = make_double2(inv[6], inv[7]);
            o[128] = make_double2(inv[8], f);
          }
        }
      }
      __syncwarp();
      asm volatile("" ::: "memory");  // GEOM 0 re-reads its scratch line below
    }
    if (GEOM != 1) {
      // right-hand sides into the elements' rows: one load per dof of the group in flight per lane
      if (GATHER) {
        mbar_wait(bar, phase);
        phase ^= 1;
      }
      if (ne == 32) {
        double val[32][KP];
#pragma unroll
        for (int el = 0; el < 32; ++el)
#pragma unroll
          for (int kp = 0; kp < KP; ++kp) {
            const int k = lane + 32 * kp;
            if (GATHER) val[el][kp] = T ? __ldca(pvec + ((k < NTOT) ? idxbuf[el * NTOT + k] : 0)) : pvec[(k < NTOT) ? idxbuf[el * NTOT + k] : 0];
            else val[el][kp] = (k < NTOT) ? pvec[(e0 + el) * NTOT + k] : 0.0;
          }
#pragma unroll
        for (int el = 0; el < 32; ++el)
#pragma unroll
          for (int kp = 0; kp < KP; ++kp) {
            const int k = lane + 32 * kp;
            if (k < NTOT) rows[el * ROW + k] = val[el][kp];
          }
      } else {
        for (int el = 0; el < ne; ++el)
          for (int k = lane; k < NTOT; k += 32)
            rows[el * ROW + k] = GATHER ? (T ? __ldca(pvec + idxbuf[el * NTOT + k]) : pvec[idxbuf[el * NTOT + k]]) : pvec[(e0 + el) * NTOT + k];
      }
      __syncwarp();
      if (GATHER && lane == 0 && grp + gstride < ngroups) {
        fence_proxy_async();  // the warp's generic reads of idxbuf are ordered before the async overwrite
        issue_idx(grp + gstride);
      }
      if (lane < ne) {
        double T[8][9];
        double gq[4][10];                                   // factors of up to three Gauss points in flight
        auto load_gq = [&](int g, double *dst) {
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            const double2 *src = gfl + (g * 5 + j) * 32;
            const double2 v = (GEOM == 2) ? __ldg(src) : *src;
            dst[2 * j] = v.x; dst[2 * j + 1] = v.y;
          }
        };
#pragma unroll
        for (int Q = 0; Q < 2; ++Q) {
          // the first two points of this half: in flight during the 4 x 9 x NOD fma below
          load_gq(4 * Q, gq[0]);
          load_gq(4 * Q + 1, gq[1]);
          double H[4][9];
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int z = 0; z < 9; ++z) H[g][z] = 0.0;
#pragma unroll UNR
          for (int t = 0; t < NOD / 2; ++t) {
            // dofs 6t..6t+5 = (x,y,z) of nodes 2t and 2t+1
            double v[6];
            lds128(row_s + (6 * t) * 8, v[0], v[1]);
            lds128(row_s + (6 * t + 2) * 8, v[2], v[3]);
            lds128(row_s + (6 * t + 4) * 8, v[4], v[5]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const double px = v[3 * h], py = v[3 * h + 1], pz = v[3 * h + 2];
              const double *der = s_der + (2 * t + h) * 24 + Q * 12;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const double d0 = der[3 * g], d1 = der[3 * g + 1], d2 = der[3 * g + 2];
                H[g][0] = fma(d0, px, H[g][0]); H[g][1] = fma(d0, py, H[g][1]); H[g][2] = fma(d0, pz, H[g][2]);
                H[g][3] = fma(d1, px, H[g][3]); H[g][4] = fma(d1, py, H[g][4]); H[g][5] = fma(d1, pz, H[g][5]);
                H[g][6] = fma(d2, px, H[g][6]); H[g][7] = fma(d2, py, H[g][7]); H[g][8] = fma(d2, pz, H[g][8]);
              }
            }
          }
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            if (g4 + 2 < 4) load_gq(4 * Q + g4 + 2, gq[g4 + 2]);     // two points ahead
            mf_point_mid(H[g4], gq[g4], gq[g4][9], T[4 * Q + g4]);
          }
        }
        // phase 2: one 24-term chain per dof, Gauss points ascending, b ascending inside
        // (the node loops are NOT fully unrolled: the whole kernel stays within the instruction cache)
#pragma unroll UNR
        for (int t = 0; t < NOD / 2; ++t) {
          double o[6];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const double *der = s_der + (2 * t + h) * 24;
            double ox = der[0] * T[0][0], oy = der[0] * T[0][1], oz = der[0] * T[0][2];
#pragma unroll
            for (int g = 0; g < 8; ++g)
#pragma unroll
              for (int b = 0; b < 3; ++b) {
                if (g == 0 && b == 0) continue;
                const double d = der[g * 3 + b];
                ox = fma(d, T[g][b * 3], ox); oy = fma(d, T[g][b * 3 + 1], oy); oz = fma(d, T[g][b * 3 + 2], oz);
              }
            o[3 * h] = ox; o[3 * h + 1] = oy; o[3 * h + 2] = oz;
          }
          sts128(row_s + (6 * t) * 8, o[0], o[1]);
          sts128(row_s + (6 * t + 2) * 8, o[2], o[3]);
          sts128(row_s + (6 * t + 4) * 8, o[4], o[5]);
        }
      }
      __syncwarp();
#pragma unroll 8
      for (int el = 0; el < ne; ++el)
#pragma unroll
        for (int kp = 0; kp < KP; ++kp) {
          const int k = lane + 32 * kp;
          if (k < NTOT) utemp[(e0 + el) * NTOT + k] = rows[el * ROW + k];
        }
      __syncwarp();
    }
  }
}

// ----------------------------------------------------------------------------
// k_apply_mf2: the same operator with TWO LANES PER ELEMENT (round 2).  k_apply_mf above keeps 72 T values, 36 H
// accumulators and the factors of three points in one thread: ~254 registers, 8 warps per SM, and ncu showed the
// FP64 pipe 49 % active with nothing saturated -- a latency problem.  Here lane `half` of an element's lane pair owns
// the Gauss points 4*half .. 4*half+3: 36 H accumulators -> 36 T values per lane, <= 128 registers, 16 warps per SM
// with the same number of elements in flight and the same FP64 work per element.  Phase 2 forms, per freedom, one chain
// over the lane's four points (first term a product, then 11 fma) and the two chains are added with one shuffle:
//   u_c(m) = [sum_{g<4} sum_b der_g(b,m) T_g(b,c)] + [sum_{g>=4} sum_b der_g(b,m) T_g(b,c)]
// (orc_apply_mf mirrors exactly this).  A warp owns 16 consecutive elements; shared-memory rows, the bulk-copied
// gather indices and the factor layout [group of 32][point][word][lane] are those of k_apply_mf.
// ----------------------------------------------------------------------------
template <int NOD>
struct Mf2Cfg {
  static constexpr int NTOT = 3 * NOD, kRow = NTOT + 2, EPW = 16;   // elements per warp
  static constexpr int kIdxBytes = EPW * NTOT * 4;
  static constexpr int kDerBytes = NOD * 24 * 8;
  static constexpr size_t smem(int warps) { return (size_t)kDerBytes + (size_t)warps * (EPW * kRow * 8 + kIdxBytes + 16); }
};

template <int NOD, bool GATHER, int GEOM, int WARPS, int UNR = 2, int PAIR = 1>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_apply_mf2(const double *__restrict__ g_coord, const int *__restrict__ ggl, const double *__restrict__ pvec,
            double *__restrict__ utemp, long long nels, const State *st, double *geom, PeerTable *T) {
  using Cfg = Mf2Cfg<NOD>;
  constexpr int NTOT = Cfg::NTOT, ROW = Cfg::kRow, KP = (NTOT + 31) / 32, EPW = Cfg::EPW;
  constexpr long long kGroupGeom = 32LL * 80;               // doubles per group of 32 elements (MfCfg::kGroupGeom)
  static_assert(NOD % 2 == 0, "nodes are processed in pairs");
  if (st && *(volatile const int *)&st->done) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // this lane's element of the warp's 16 and its half of the points: partner lane = lane ^ PAIR (PAIR 1: adjacent lanes,
  // PAIR 16: the two half-warps -- every quarter-warp then reads ONE der address)
  const int el = PAIR == 1 ? lane >> 1 : lane & 15, half = PAIR == 1 ? lane & 1 : lane >> 4;
  double *s_der = reinterpret_cast<double *>(smem_raw);                                   // [m][g][b]
  double *rows = reinterpret_cast<double *>(smem_raw + Cfg::kDerBytes) + (size_t)w * EPW * ROW;
  int *idxbuf = reinterpret_cast<int *>(smem_raw + Cfg::kDerBytes + (size_t)WARPS * EPW * ROW * 8) + (size_t)w * EPW * NTOT;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + Cfg::kDerBytes + (size_t)WARPS * (EPW * ROW * 8 + Cfg::kIdxBytes));
  const uint32_t bar = smem_u32(&bars[w]);
  const uint32_t row_s = smem_u32(rows + el * ROW);          // this lane pair's element row
  for (int q = threadIdx.x; q < NOD * 24; q += blockDim.x) {
    const int m = q / 24, r = q - m * 24, g = r / 3, b = r - g * 3;
    s_der[q] = c_tab.der[g * 60 + b * 20 + m];
  }
  __syncthreads();
  const long long nhg = (nels + EPW - 1) / EPW;              // half-groups of 16 elements
  const long long hstride = (long long)gridDim.x * WARPS;
  uint64_t policy = 0;
  uint32_t phase = 0;
  auto issue_idx = [&](long long hg) {                       // lane 0 only
    const long long e0 = hg * EPW;
    const int ne = (int)((nels - e0) < EPW ? (nels - e0) : EPW);
    const uint32_t bytes = (uint32_t)ne * NTOT * 4;
    mbar_expect_tx(bar, bytes);
    bulk_g2s(smem_u32(idxbuf), ggl + e0 * NTOT, bytes, bar, policy);
  };
  long long hg = (long long)blockIdx.x * WARPS + w;
  if (GATHER && GEOM != 1) {
    if (lane == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
      policy = policy_evict_first();
      if (hg < nhg) issue_idx(hg);
    }
    __syncwarp();
    if (T) warp_wait_fwd(T, st);
  }
  for (; hg < nhg; hg += hstride) {
    const long long e0 = hg * EPW;
    const int ne = (int)((nels - e0) < EPW ? (nels - e0) : EPW);
    const unsigned int pairmask = ne >= EPW ? 0xffffffffu : (PAIR == 1 ? ((1u << (2 * ne)) - 1u) : (((1u << ne) - 1u) * 0x10001u));
    // this lane pair's 16-byte words of the factors: word (g*5+j) of the element at gfl[(g*5+j)*32]
    double2 *gfl = reinterpret_cast<double2 *>(geom + (GEOM == 0 ? (long long)blockIdx.x * WARPS + w : (hg >> 1)) * kGroupGeom) +
                   (GEOM == 0 ? el : (int)(hg & 1) * 16 + el);
    if (GEOM == 2 && lane == 0 && hg + hstride < nhg) {
      const long long nh = hg + hstride;                     // next half-group's factors towards L2
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(geom + (nh >> 1) * kGroupGeom),
                   "r"((uint32_t)(kGroupGeom * 8)) : "memory");
    }
    if (GEOM != 2) {
      // coordinates g_coord_pp(nod,3,nels) -> the elements' rows, [b*NOD+m]
      for (int e2 = 0; e2 < ne; ++e2)
#pragma unroll
        for (int kp = 0; kp < KP; ++kp) {
          const int k = lane + 32 * kp;
          if (k < NTOT) rows[e2 * ROW + k] = g_coord[(e0 + e2) * NTOT + k];
        }
      __syncwarp();
      if (el < ne) {
        double J[4][9];
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int z = 0; z < 9; ++z) J[g][z] = 0.0;
#pragma unroll UNR
        for (int t = 0; t < NOD / 2; ++t) {
          double cx[2], cy[2], cz[2];
          lds128(row_s + (2 * t) * 8, cx[0], cx[1]);
          lds128(row_s + (NOD + 2 * t) * 8, cy[0], cy[1]);
          lds128(row_s + (2 * NOD + 2 * t) * 8, cz[0], cz[1]);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const double *der = s_der + (2 * t + h) * 24 + half * 12;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const double d0 = der[3 * g], d1 = der[3 * g + 1], d2 = der[3 * g + 2];
              J[g][0] = fma(d0, cx[h], J[g][0]); J[g][1] = fma(d1, cx[h], J[g][1]); J[g][2] = fma(d2, cx[h], J[g][2]);
              J[g][3] = fma(d0, cy[h], J[g][3]); J[g][4] = fma(d1, cy[h], J[g][4]); J[g][5] = fma(d2, cy[h], J[g][5]);
              J[g][6] = fma(d0, cz[h], J[g][6]); J[g][7] = fma(d1, cz[h], J[g][7]); J[g][8] = fma(d2, cz[h], J[g][8]);
            }
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          double inv[9];
          const double det = det3(J[g]);
          inv3(J[g], det, inv);
          const double f = det * c_tab.weights[4 * half + g];
          double2 *o = gfl + (4 * half + g) * 5 * 32;
          o[0] = make_double2(inv[0], inv[1]); o[32] = make_double2(inv[2], inv[3]);
          o[64] = make_double2(inv[4], inv[5]); o[96] = make_double2(inv[6], inv[7]);
          o[128] = make_double2(inv[8], f);
        }
      }
      __syncwarp();
      asm volatile("" ::: "memory");  // GEOM 0 re-reads its scratch line below
    }
    if (GEOM != 1) {
      if (GATHER) {
        mbar_wait(bar, phase);
        phase ^= 1;
      }
      if (ne == EPW) {
        double val[EPW][KP];
#pragma unroll
        for (int e2 = 0; e2 < EPW; ++e2)
#pragma unroll
          for (int kp = 0; kp < KP; ++kp) {
            const int k = lane + 32 * kp;
            if (GATHER) {
              const int idx = (k < NTOT) ? idxbuf[e2 * NTOT + k] : 0;
              val[e2][kp] = T ? __ldca(pvec + idx) : pvec[idx];
            } else val[e2][kp] = (k < NTOT) ? pvec[(e0 + e2) * NTOT + k] : 0.0;
          }
#pragma unroll
        for (int e2 = 0; e2 < EPW; ++e2)
#pragma unroll
          for (int kp = 0; kp < KP; ++kp) {
            const int k = lane + 32 * kp;
            if (k < NTOT) rows[e2 * ROW + k] = val[e2][kp];
          }
      } else {
        for (int e2 = 0; e2 < ne; ++e2)
          for (int k = lane; k < NTOT; k += 32)
            rows[e2 * ROW + k] = GATHER ? (T ? __ldca(pvec + idxbuf[e2 * NTOT + k]) : pvec[idxbuf[e2 * NTOT + k]]) : pvec[(e0 + e2) * NTOT + k];
      }
      __syncwarp();
      if (GATHER && lane == 0 && hg + hstride < nhg) {
        fence_proxy_async();  // the warp's generic reads of idxbuf are ordered before the async overwrite
        issue_idx(hg + hstride);
      }
      if (el < ne) {
        double Tm[4][9];
        double gq[2][10];                                     // factors of the current point and the next one
        auto load_gq = [&](int g4, double *dst) {
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            const double2 *src = gfl + ((4 * half + g4) * 5 + j) * 32;
            const double2 v = (GEOM == 2) ? __ldg(src) : *src;
            dst[2 * j] = v.x; dst[2 * j + 1] = v.y;
          }
        };
        load_gq(0, gq[0]);                                    // in flight during the 4 x 9 x NOD fma below
        double H[4][9];
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int z = 0; z < 9; ++z) H[g][z] = 0.0;
#pragma unroll UNR
        for (int t = 0; t < NOD / 2; ++t) {
          double v[6];
          lds128(row_s + (6 * t) * 8, v[0], v[1]);
          lds128(row_s + (6 * t + 2) * 8, v[2], v[3]);
          lds128(row_s + (6 * t + 4) * 8, v[4], v[5]);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const double px = v[3 * h], py = v[3 * h + 1], pz = v[3 * h + 2];
            const double *der = s_der + (2 * t + h) * 24 + half * 12;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const double d0 = der[3 * g], d1 = der[3 * g + 1], d2 = der[3 * g + 2];
              H[g][0] = fma(d0, px, H[g][0]); H[g][1] = fma(d0, py, H[g][1]); H[g][2] = fma(d0, pz, H[g][2]);
              H[g][3] = fma(d1, px, H[g][3]); H[g][4] = fma(d1, py, H[g][4]); H[g][5] = fma(d1, pz, H[g][5]);
              H[g][6] = fma(d2, px, H[g][6]); H[g][7] = fma(d2, py, H[g][7]); H[g][8] = fma(d2, pz, H[g][8]);
            }
          }
        }
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
          if (g4 + 1 < 4) load_gq(g4 + 1, gq[(g4 + 1) & 1]);  // one point ahead
          mf_point_mid(H[g4], gq[g4 & 1], gq[g4 & 1][9], Tm[g4]);
        }
        // every lane has read its row of right-hand sides: the pair may now overwrite it with the products
        __syncwarp(pairmask);
        // phase 2: per freedom one 12-term chain over this lane's four points, then the partner's chain is added
        // (exchanging only the six sums each lane needs -- two node pairs per shuffle round -- was measured SLOWER:
        // 0.804 against 0.754 ms, the longer loop body loses more overlap than the 60 shuffles cost)
#pragma unroll UNR
        for (int t = 0; t < NOD / 2; ++t) {
          double o[6];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const double *der = s_der + (2 * t + h) * 24 + half * 12;
            double ox = der[0] * Tm[0][0], oy = der[0] * Tm[0][1], oz = der[0] * Tm[0][2];
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
              for (int b = 0; b < 3; ++b) {
                if (g == 0 && b == 0) continue;
                const double d = der[g * 3 + b];
                ox = fma(d, Tm[g][b * 3], ox); oy = fma(d, Tm[g][b * 3 + 1], oy); oz = fma(d, Tm[g][b * 3 + 2], oz);
              }
            o[3 * h] = ox; o[3 * h + 1] = oy; o[3 * h + 2] = oz;
          }
#pragma unroll
          for (int q = 0; q < 6; ++q) o[q] = o[q] + __shfl_xor_sync(pairmask, o[q], PAIR);   // (points 0-3) + (points 4-7)
          if ((t & 1) == half) {
            sts128(row_s + (6 * t) * 8, o[0], o[1]);
            sts128(row_s + (6 * t + 2) * 8, o[2], o[3]);
            sts128(row_s + (6 * t + 4) * 8, o[4], o[5]);
          }
        }
      }
      __syncwarp();
      for (int e2 = 0; e2 < ne; ++e2)
#pragma unroll
        for (int kp = 0; kp < KP; ++kp) {
          const int k = lane + 32 * kp;
          if (k < NTOT) utemp[(e0 + e2) * NTOT + k] = rows[e2 * ROW + k];
        }
      __syncwarp();
    }
  }
}

// ----------------------------------------------------------------------------
// k_apply_mf3: the same operator on the FP64 TENSOR pipe (round 2, second build).  The two node sums of the matrix-free
// product are small GEMMs with a CONSTANT operand (der is the same for every element):
//   phase 1   H_e,g(b,c) = sum_m der_g(b,m) p_e,c(m)       [8 elements x 3] x [NOD] x [8 points x 3]
//   phase 3   u_e,c(m)   = sum_(g,b) T_e,g(b,c) der_g(b,m)  [8 elements x 3] x [8 points x 3] x [NOD]
// issued as mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4; measured on the B200, scripts/probe/dmma_probe.cu: 37.2 TFLOP/s
// against 33.9 of the DFMA loop, and D = fma(a3,b3,fma(a2,b2,fma(a1,b1,fma(a0,b0,c)))) bit for bit, i.e. a k-ascending
// fma chain -- which is how the oracle states it).  A warp pass is 8 elements: lane (r = lane/4, q = lane%4) feeds
// element r.  The fragment layouts are chosen so that NOTHING is transposed between the phases:
//   phase 1: rows (M) = (component c | element r), k = node, columns (N) = (b | point g): the accumulator fragment of
//            lane (r,q) is the complete 3x3 H of element r at the points 2q and 2q+1 -> mf_point_mid in registers;
//   phase 3: rows = (c | element r), k = (h, b | q) <-> point 2q+h, columns = node: the A fragments are exactly the
//            T values the lane has just formed.
// der lives in registers as the B fragments of both phases (33 doubles per lane for 20-node bricks), so the shared-memory
// der table and its 30 LDS wavefronts per element of k_apply_mf / k_apply_mf2 are gone; shared memory only stages the
// gathered right-hand sides (coalesced gather -> rows -> A fragments) and the products (fragments -> rows -> coalesced
// store).  Chain orders (orc_apply_mf, order 2): H and jac node-ascending from 0.0 as before; u_c(m) ONE 24-term fma
// chain from 0.0 over (h = 0,1 | b = 0,1,2 | q = 0..3), point g = 2q+h.
// GEOM 0: jac comes from the same phase-1 product on the coordinates, det / inverse (adjugate times ONE reciprocal per
//         point, inv3_recip) stay in registers (no scratch);
// GEOM 1: only writes those factors ([group of 32][point][word][lane]: the layout of k_apply_mf / k_apply_mf2) -- the
//         setup of mode 2 for this kernel family;
// GEOM 2: reads them.  Same bits in modes 0 and 2.
// ----------------------------------------------------------------------------
template <int NOD>
struct Mf3Cfg {
  static constexpr int NTOT = 3 * NOD, EPP = 8;                       // elements per warp pass
  static constexpr int KS1 = NOD / 4, NT3 = (NOD + 7) / 8;            // k-steps of phase 1, node tiles of phase 3
  static constexpr int NF = KS1 * 3 + 6 * NT3;                        // der fragments per lane (33 / 12)
  // row stride (doubles) == 4 or 12 mod 16: the A-fragment reads rows[r][3(4s+q)+c] of a half-warp hit 16 distinct
  // 8-byte bank pairs (20-node: 60, 8-node: 28)
  static constexpr int kRow = (NTOT % 16 == 4 || NTOT % 16 == 12) ? NTOT : NTOT + ((12 - NTOT % 16) + 16) % 16;
  static constexpr int kIdxBytes = EPP * NTOT * 4;
  static constexpr int kFragBytes = NF * 32 * 8;
  // per warp: two row buffers (pass n is computed while pass n+1 lands), with GEOM 0 / 1 one coordinate buffer, the indices
  __host__ __device__ static constexpr size_t per_warp(int geom) { return (size_t)(geom != 2 ? 3 : 2) * EPP * kRow * 8 + kIdxBytes; }
  static constexpr size_t smem(int warps, int geom) { return (size_t)kFragBytes + (size_t)warps * (per_warp(geom) + 16); }
};

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// BREG 2: all der fragments live in registers for the whole kernel; 1: those of phase 1 (and the Jacobian product);
// 0: every fragment is read from a conflict-free shared-memory table [fragment][lane] at its use.
// The products leave through ONE bulk async store per pass (shared -> global, no LSU wavefronts).
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int NOD, bool GATHER, int GEOM, int WARPS, int BREG>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_apply_mf3(const double *__restrict__ g_coord, const int *__restrict__ ggl, const double *__restrict__ pvec,
            double *__restrict__ utemp, long long nels, const State *st, double *geom, PeerTable *T) {
  using Cfg = Mf3Cfg<NOD>;
  constexpr int NTOT = Cfg::NTOT, ROW = Cfg::kRow, EPP = Cfg::EPP, KS1 = Cfg::KS1, NT3 = Cfg::NT3, NF = Cfg::NF;
  constexpr int KP = (EPP * NTOT + 31) / 32;                 // flat (element, freedom) words of a pass per lane
  constexpr long long kGroupGeom = 32LL * 80;                // doubles per group of 32 elements (MfCfg::kGroupGeom)
  static_assert(NOD % 4 == 0, "phase 1 takes the nodes four at a time");
  static_assert(GEOM >= 0 && GEOM <= 2, "0: rebuild the factors every call, 1: only write them (setup of mode 2), 2: read them");
  if (st && *(volatile const int *)&st->done) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, r = lane >> 2, q = lane & 3;
  double *s_frag = reinterpret_cast<double *>(smem_raw);
  unsigned char *mine = smem_raw + Cfg::kFragBytes + (size_t)w * Cfg::per_warp(GEOM);
  double *rows2 = reinterpret_cast<double *>(mine);                                        // [2][EPP*ROW]
  double *cbuf = rows2 + 2 * EPP * ROW;                                                    // GEOM 0 / 1 only
  int *idxbuf = reinterpret_cast<int *>(mine + (size_t)(GEOM != 2 ? 3 : 2) * EPP * ROW * 8);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + Cfg::kFragBytes + (size_t)WARPS * Cfg::per_warp(GEOM));
  const uint32_t bar = smem_u32(&bars[w]);
  // der as B fragments (lane holds B[k = q][n = r]):
  //   fragment s*3+b            phase 1, k-step s, column tile b:   der_{g=r}(b, m = 4s+q)
  //   fragment 3 KS1 + (3h+b) NT3 + nt   phase 3, k-step 3h+b, node tile nt: der_{g=2q+h}(b, m = 8nt+r), 0 beyond the last node
  for (int i = threadIdx.x; i < NF * 32; i += blockDim.x) {
    const int f = i >> 5, l = i & 31, rr = l >> 2, qq = l & 3;
    double v;
    if (f < 3 * KS1) {
      const int s = f / 3, b = f - 3 * s;
      v = c_tab.der[rr * 60 + b * 20 + 4 * s + qq];
    } else {
      const int f3 = f - 3 * KS1, sb = f3 / NT3, nt = f3 - sb * NT3, h = sb / 3, b = sb - 3 * h;
      v = (8 * nt + rr < NOD) ? c_tab.der[(2 * qq + h) * 60 + b * 20 + 8 * nt + rr] : 0.0;
    }
    s_frag[i] = v;
  }
  __syncthreads();
  constexpr int NREG = BREG == 2 ? NF : (BREG == 1 ? 3 * KS1 : 0);
  double Breg[NREG > 0 ? NREG : 1];
#pragma unroll
  for (int f = 0; f < NREG; ++f) {
    Breg[f] = s_frag[f * 32 + lane];
    asm volatile("" : "+d"(Breg[f]));                        // opaque: kept in a register, not re-read
  }
  auto fragB = [&](int f) -> double { return f < NREG ? Breg[f] : s_frag[f * 32 + lane]; };
  const long long npass = (nels + EPP - 1) / EPP;
  const long long pstride = (long long)gridDim.x * WARPS;
  uint64_t policy = 0;
  uint32_t phase = 0;
  auto issue_idx = [&](long long ps) {                       // lane 0 only
    const long long e0 = ps * EPP;
    const int ne = (int)((nels - e0) < EPP ? (nels - e0) : EPP);
    const uint32_t bytes = (uint32_t)ne * NTOT * 4;
    mbar_expect_tx(bar, bytes);
    bulk_g2s(smem_u32(idxbuf), ggl + e0 * NTOT, bytes, bar, policy);
  };
  // right-hand sides (and with GEOM 0 the coordinates) of pass ps: asynchronous 8-byte copies straight into shared
  // memory, coalesced over the pass's (element, freedom) words; nothing is held in registers while they fly
  auto issue_data = [&](long long ps, double *dst) {
    const long long e0 = ps * EPP;
    const int ne = (int)((nels - e0) < EPP ? (nels - e0) : EPP);
    const int nw = ne * NTOT;
    if (GEOM != 1) {
#pragma unroll
      for (int kp = 0; kp < KP; ++kp) {
        const int f = lane + 32 * kp;
        if (f < EPP * NTOT) {
          double *d = dst + (f / NTOT) * ROW + (f % NTOT);
          if (f < nw) cp_async8(smem_u32(d), GATHER ? pvec + idxbuf[f] : pvec + e0 * NTOT + f);
          else *d = 0.0;
        }
      }
    }
    if (GEOM != 2) {
#pragma unroll
      for (int kp = 0; kp < KP; ++kp) {
        const int f = lane + 32 * kp;
        if (f < EPP * NTOT) {
          double *d = cbuf + (f / NTOT) * ROW + (f % NTOT);
          if (f < nw) cp_async8(smem_u32(d), g_coord + e0 * NTOT + f);
          else *d = 0.0;
        }
      }
    }
    cp_async_commit();
  };
  long long ps = (long long)blockIdx.x * WARPS + w;
  if (ps >= npass) return;
  if (GATHER) {
    if (lane == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
      policy = policy_evict_first();
      issue_idx(ps);
    }
    __syncwarp();
    // N ranks, peer transport: the owners' values are in my halo segment of pvec; nothing of pvec is read before this
    if (T) warp_wait_fwd(T, st);
    mbar_wait(bar, phase);
    phase ^= 1;
  }
  issue_data(ps, rows2);
  __syncwarp();
  if (GATHER && lane == 0 && ps + pstride < npass) {
    fence_proxy_async();                                     // generic reads of idxbuf before the async overwrite
    issue_idx(ps + pstride);
  }
  int buf = 0;
  for (; ps < npass; ps += pstride, buf ^= 1) {
    const long long e0 = ps * EPP;
    const int ne = (int)((nels - e0) < EPP ? (nels - e0) : EPP);
    const int nw = ne * NTOT;                                // words of this pass
    double *rows = rows2 + buf * EPP * ROW;
    double gq[2][10];                                        // jac^-1 (9) and det*w of the lane's two points
    if (GEOM == 2) {
      // word (g*5+j) of element i of a group at [(g*5+j)*32 + i] (16-byte words); in flight during phase 1
      const double2 *gfl = reinterpret_cast<const double2 *>(geom + (e0 >> 5) * kGroupGeom) + ((int)(e0 & 31) + r);
      if (r < ne) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            const double2 v = __ldg(gfl + ((2 * q + h) * 5 + j) * 32);
            gq[h][2 * j] = v.x; gq[h][2 * j + 1] = v.y;
          }
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j = 0; j < 10; ++j) gq[h][j] = 0.0;
      }
      const long long en = (ps + pstride) * EPP;             // next pass: its 40 lines of factors towards L2
      if (r == 0 && en < nels) {
        const double2 *gn = reinterpret_cast<const double2 *>(geom + (en >> 5) * kGroupGeom) + (int)(en & 31);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j = 0; j < 5; ++j) asm volatile("prefetch.global.L2 [%0];" ::"l"(gn + ((2 * q + h) * 5 + j) * 32));
      }
    }
    cp_async_wait_all();
    __syncwarp();                                            // this pass's rows (and coordinates) have landed
    if (GEOM != 2) {
      double J[3][3][2];                                     // [b: x,y,z][a][h] = jac(a,b) at point 2q+h
#pragma unroll
      for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int a = 0; a < 3; ++a) J[b][a][0] = J[b][a][1] = 0.0;
#pragma unroll
      for (int s = 0; s < KS1; ++s) {
        double a3[3];
#pragma unroll
        for (int b = 0; b < 3; ++b) a3[b] = cbuf[r * ROW + b * NOD + 4 * s + q];
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int a = 0; a < 3; ++a) dmma884(J[b][a][0], J[b][a][1], a3[b], fragB(s * 3 + a));
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double Jm[9];
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int a = 0; a < 3; ++a) Jm[b * 3 + a] = J[b][a][h];
        const double det = det3(Jm);
        inv3_recip(Jm, det, gq[h]);
        gq[h][9] = det * c_tab.weights[2 * q + h];
      }
      __syncwarp();                                          // the coordinate buffer is free for the next pass
    }
    if (GEOM == 1) {
      // setup of mode 2: word (g*5+j) of element i of a group of 32 at [(g*5+j)*32 + i] (16-byte words)
      double2 *gfl = reinterpret_cast<double2 *>(geom + (e0 >> 5) * kGroupGeom) + ((int)(e0 & 31) + r);
      if (r < ne) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j = 0; j < 5; ++j) gfl[((2 * q + h) * 5 + j) * 32] = make_double2(gq[h][2 * j], gq[h][2 * j + 1]);
      }
      if (ps + pstride < npass) issue_data(ps + pstride, rows2);
      continue;
    }
    // the next pass's data starts to move now and lands during this pass's arithmetic
    if (ps + pstride < npass) {
      if (GATHER) {
        mbar_wait(bar, phase);
        phase ^= 1;
      }
      bulk_wait_read_all();                                  // the previous pass's products have left the other buffer
      __syncwarp();
      issue_data(ps + pstride, rows2 + (buf ^ 1) * EPP * ROW);
      __syncwarp();
      if (GATHER && lane == 0 && ps + 2 * pstride < npass) {
        fence_proxy_async();
        issue_idx(ps + 2 * pstride);
      }
    }
    // phase 1: H[c][b][h] = H_{element r, point 2q+h}(b,c)
    double H[3][3][2];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int b = 0; b < 3; ++b) H[c][b][0] = H[c][b][1] = 0.0;
#pragma unroll
    for (int s = 0; s < KS1; ++s) {
      double a3[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) a3[c] = rows[r * ROW + 3 * (4 * s + q) + c];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int b = 0; b < 3; ++b) dmma884(H[c][b][0], H[c][b][1], a3[c], fragB(s * 3 + b));
    }
    double Tm[2][9];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double Hm[9];
#pragma unroll
      for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int c = 0; c < 3; ++c) Hm[b * 3 + c] = H[c][b][h];
      mf_point_mid_iso(Hm, gq[h], gq[h][9], Tm[h]);
    }
    // phase 3: U[c][nt][h'] = u_c(m = 8nt + 2q + h') of element r
    double U[3][NT3][2];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int nt = 0; nt < NT3; ++nt) U[c][nt][0] = U[c][nt][1] = 0.0;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int nt = 0; nt < NT3; ++nt) dmma884(U[c][nt][0], U[c][nt][1], Tm[h][b * 3 + c], fragB(3 * KS1 + (3 * h + b) * NT3 + nt));
    // every lane's A-fragment reads of the rows fed the warp-wide mma above, so they are complete; the barrier states
    // that order for the memory model (and for racecheck): the rows may take the products
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < NT3; ++nt)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int m = 8 * nt + 2 * q + hh;
        if (m < NOD) {
#pragma unroll
          for (int c = 0; c < 3; ++c) rows[r * ROW + 3 * m + c] = U[c][nt][hh];
        }
      }
    fence_proxy_async();                                     // my generic stores before the async proxy reads them
    __syncwarp();
    if (ROW == NTOT) {
      if (lane == 0) { bulk_s2g(utemp + e0 * NTOT, smem_u32(rows), (uint32_t)nw * 8); bulk_commit(); }
    } else if (lane < ne) {
      bulk_s2g(utemp + (e0 + lane) * NTOT, smem_u32(rows + lane * ROW), NTOT * 8);
      bulk_commit();
    }
  }
  bulk_wait_all();                                           // shared memory must outlive the stores
}

// p1210's Gauss-point update in operator form (orc_p1210_elements_mf): H(b,c) at [b*3+c], inv = jac^-1, f = det*w, the
// point's etensor / tensor (6 doubles each, updated in place) -> T(b,c).  Strain from the displacement gradient minus
// etensor, elastic trial stress (deemat's zeros left out), on yield the scaled-back stress -> vmpl -> dee - fac*pl formed
// entry by entry inside the stress sum, stress * det*w -> T = jac^-T S.
struct VmParams {
  double e, v, sbary;
};
// second invariant as invar forms it (new_library.f90:1891-1897): dsbar = sqrt(3)*sqrt(d2)
__device__ __forceinline__ double vm_dsbar(const double *s) {
  const double d2 = ((s[0] - s[1]) * (s[0] - s[1]) + (s[1] - s[2]) * (s[1] - s[2]) + (s[2] - s[0]) * (s[2] - s[0])) / 6.0 +
                    s[3] * s[3] + s[4] * s[4] + s[5] * s[5];
  return sqrt(3.0) * sqrt(d2);
}
__device__ __forceinline__ void vm_point_mid(const double *H, const double *inv, double f, double *et, double *te,
                                             const VmParams &P, double *T) {
  double G[9];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double s = inv[a] * H[c];
      s = fma(inv[3 + a], H[3 + c], s);
      s = fma(inv[6 + a], H[6 + c], s);
      G[a * 3 + c] = s;
    }
  double eps[6] = {G[0], G[4], G[8], G[3] + G[1], G[7] + G[5], G[6] + G[2]};
  double sigma[6], stressv[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) eps[r] = eps[r] - et[r];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double s = c_tab.dee[r] * eps[0];
    s = fma(c_tab.dee[6 + r], eps[1], s);
    s = fma(c_tab.dee[12 + r], eps[2], s);
    sigma[r] = s;
  }
#pragma unroll
  for (int r = 3; r < 6; ++r) sigma[r] = c_tab.dee[r * 6 + r] * eps[r];
#pragma unroll
  for (int r = 0; r < 6; ++r) stressv[r] = sigma[r] + te[r];
  const double fnew = vm_dsbar(stressv) - P.sbary;
  if (fnew >= 0.0) {                                              // yield is violated
    const double fy = vm_dsbar(te) - P.sbary, fac = fnew / (fnew - fy);
#pragma unroll
    for (int r = 0; r < 6; ++r) stressv[r] = te[r] + (1.0 - fac) * sigma[r];
    const double sx = stressv[0], sy = stressv[1], sz = stressv[2], txy = stressv[3], tyz = stressv[4], tzx = stressv[5];
    const double dsb = sqrt((sx - sy) * (sx - sy) + (sy - sz) * (sy - sz) + (sz - sx) * (sz - sx) + 6.0 * (txy * txy) +
                            6.0 * (tyz * tyz) + 6.0 * (tzx * tzx)) / sqrt(2.0);
    const double ee = 1.5 * P.e / ((1.0 + P.v) * dsb * dsb);
    const double term[6] = {(2.0 * sx - sy - sz) / 3.0, (2.0 * sy - sz - sx) / 3.0, (2.0 * sz - sx - sy) / 3.0, txy, tyz, tzx};
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      double s = (c_tab.dee[r] - fac * (term[r] * term[0] * ee)) * eps[0];
#pragma unroll
      for (int q = 1; q < 6; ++q) s = fma(c_tab.dee[q * 6 + r] - fac * (term[r] * term[q] * ee), eps[q], s);
      sigma[r] = s + te[r];
    }
  } else {
#pragma unroll
    for (int r = 0; r < 6; ++r) sigma[r] = stressv[r];
  }
#pragma unroll
  for (int r = 0; r < 6; ++r) { te[r] = sigma[r]; et[r] = et[r] + eps[r]; }
  double sg[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) sg[r] = sigma[r] * f;
  const double S[9] = {sg[0], sg[3], sg[5], sg[3], sg[1], sg[4], sg[5], sg[4], sg[2]};
#pragma unroll
  for (int b = 0; b < 3; ++b)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double s = inv[b * 3] * S[c];
      s = fma(inv[b * 3 + 1], S[3 + c], s);
      s = fma(inv[b * 3 + 2], S[6 + c], s);
      T[b * 3 + c] = s;
    }
}

// ----------------------------------------------------------------------------
// k_apply_mf4: k_apply_mf3's arithmetic, WARP-SPECIALISED.  ncu of k_apply_mf3 (profiles/r02_mf_tensor_kernel.md): every
// warp alternates between moving data (index loads -> gathers -> staging, LSU-bound, no FP64) and arithmetic, the DMMA
// pipe is busy 48 % of the time and the der-fragment table costs as many shared-memory wavefronts as the gather.  Here
// a CTA of 384 threads is one producer warpgroup (4 warps, 56 registers) and two consumer warpgroups (8 warps, 224
// registers, setmaxnreg):
//   producer warp j (scheduler j) feeds consumer warps j and j+4 (same scheduler): per pass of 8 elements it waits for a
//     free ring stage, reads the pass's gather indices (bulk-copied two passes ahead), issues the asynchronous 8-byte
//     gathers straight into the stage (cp.async, completion counted on the stage's `full` mbarrier), with GEOM 0 one bulk
//     copy of the coordinates, and prefetches the pass's geometric factors into L2;
//   consumer warp: waits for `full`, runs the two tensor-pipe products and the Gauss-point arithmetic with ALL der
//     fragments in registers, writes the products over the stage, sends them off with one bulk async store and hands
//     the previous stage back (`empty`) once its store has been read.
// Three ring stages per consumer.  Same bits as k_apply_mf3 (order 2 of the oracle).
// ----------------------------------------------------------------------------
template <int NOD, int NCW>
struct Mf4Cfg {
  static constexpr int NTOT = 3 * NOD, EPP = 8, KS1 = NOD / 4, NT3 = (NOD + 7) / 8, NF = KS1 * 3 + 6 * NT3;
  static constexpr int kRow = Mf3Cfg<NOD>::kRow;
  // NCW consumer warps (12: three per scheduler, 152 registers, the phase-3 fragments from a shared-memory table -- the
  // default, measured 0.505 against 0.602 ms at config C; 8: two per scheduler, 224 registers, every der fragment in
  // registers) + 4 producer warps.
  // Ring stages per consumer (a stage is handed back one pass after its store was issued, so the producer runs S-2
  // passes ahead of the arithmetic) and index buffers per consumer (bulk-copied NIDX passes ahead).
  static constexpr int S = NCW == 8 ? (NOD == 20 ? 4 : 6) : (NOD == 20 ? 3 : 4), NIDX = NCW == 8 ? 4 : 2;
  static constexpr int NCONS = NCW, NPROD = 4, KPW = NCW / NPROD, kThreads = 32 * (NCONS + NPROD);
  static constexpr int kConsRegs = NCW == 8 ? 224 : 152, kProdRegs = 56;
  static constexpr bool kAllFragsInRegs = NCW == 8;
  static constexpr int kStageBytes = EPP * kRow * 8, kIdxBytes = EPP * NTOT * 4;
  static constexpr int kBarsPer = 2 * S + NIDX;
  static constexpr int kFragBytes = kAllFragsInRegs ? 0 : 6 * NT3 * 32 * 8;
  static constexpr size_t kSmem = (size_t)NCONS * S * kStageBytes + (size_t)NCONS * NIDX * kIdxBytes + kFragBytes + (size_t)NCONS * kBarsPer * 8;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// MID 1 (k_p1210_mf): p1210's elasto-plastic Gauss-point update (vm_point_mid) instead of the elastic one, the points'
// etensor_pp / tensor_pp (6, nip, nels) read and written in place, and utemp = 0 - products (elements_2 of p1210.f90).
template <int NOD, bool GATHER, int GEOM, int NCW, int MID = 0>
__global__ void __launch_bounds__(Mf4Cfg<NOD, NCW>::kThreads, 1)
k_apply_mf4(const double *__restrict__ g_coord, const int *__restrict__ ggl, const double *__restrict__ pvec,
            double *__restrict__ utemp, long long nels, const State *st, const double *__restrict__ geom, PeerTable *T,
            double *etensor = nullptr, double *tensor = nullptr, VmParams vmp = VmParams{0.0, 0.0, 0.0}) {
  using Cfg = Mf4Cfg<NOD, NCW>;
  constexpr int NTOT = Cfg::NTOT, ROW = Cfg::kRow, EPP = Cfg::EPP, KS1 = Cfg::KS1, NT3 = Cfg::NT3, NF = Cfg::NF, S = Cfg::S;
  constexpr int NCONS = Cfg::NCONS, NPROD = Cfg::NPROD, NIDX = Cfg::NIDX, BP = Cfg::kBarsPer, KPW = Cfg::KPW;
  constexpr bool ALLREG = Cfg::kAllFragsInRegs;
  constexpr int KP = (EPP * NTOT + 31) / 32;
  constexpr long long kGroupGeom = 32LL * 80;
  static_assert(NOD % 4 == 0 && (EPP * NTOT) % 32 == 0, "a pass is a whole number of warp-wide words");
  static_assert(GEOM == 0 || GEOM == 2, "the factors of mode 2 are written by k_apply_mf3<GEOM 1>");
  static_assert((NIDX & (NIDX - 1)) == 0, "index ring: power of two");
  if (st && *(volatile const int *)&st->done) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char *stage0 = smem_raw;                                                         // [NCONS][S] rows
  unsigned char *idx0 = smem_raw + (size_t)NCONS * S * Cfg::kStageBytes;                    // [NCONS][NIDX]
  double *s_frag = reinterpret_cast<double *>(idx0 + (size_t)NCONS * NIDX * Cfg::kIdxBytes);   // phase-3 fragments [f][lane]
  uint64_t *bars = reinterpret_cast<uint64_t *>(idx0 + (size_t)NCONS * NIDX * Cfg::kIdxBytes + Cfg::kFragBytes);
  auto full_bar = [&](int c, int s) { return smem_u32(&bars[c * BP + s]); };
  auto empty_bar = [&](int c, int s) { return smem_u32(&bars[c * BP + S + s]); };
  auto idx_bar = [&](int c, int b) { return smem_u32(&bars[c * BP + 2 * S + b]); };
  if (threadIdx.x < NCONS) {
    const int c = threadIdx.x;
    for (int s2 = 0; s2 < S; ++s2) {
      mbar_init(full_bar(c, s2), 32);                          // the producer's 32 lanes, each when its copies have landed
      mbar_init(empty_bar(c, s2), 1);
    }
    for (int b = 0; b < NIDX; ++b) mbar_init(idx_bar(c, b), 1);
    fence_mbar_init();
  }
  if (!ALLREG) {
    for (int i = threadIdx.x; i < 6 * NT3 * 32; i += blockDim.x) {
      const int f3 = i >> 5, l = i & 31, rr = l >> 2, qq = l & 3, sb = f3 / NT3, nt = f3 - sb * NT3, h = sb / 3, b = sb - 3 * h;
      s_frag[i] = (8 * nt + rr < NOD) ? c_tab.der[(2 * qq + h) * 60 + b * 20 + 8 * nt + rr] : 0.0;
    }
  }
  __syncthreads();
  const long long npass = (nels + EPP - 1) / EPP;
  const long long stride = (long long)gridDim.x * NCONS;

  if (w < NPROD) {
    // ------------------------------------------------ producer ------------------------------------------------
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::kProdRegs));
    const uint64_t policy = policy_evict_first();
    auto issue_idx = [&](int c, int b, long long ps) {         // lane 0 only
      const long long e0 = ps * EPP;
      const int ne = (int)((nels - e0) < EPP ? (nels - e0) : EPP);
      const uint32_t bytes = (uint32_t)ne * NTOT * 4;
      mbar_expect_tx(idx_bar(c, b), bytes);
      bulk_g2s(smem_u32(idx0 + (size_t)(c * NIDX + b) * Cfg::kIdxBytes), ggl + e0 * NTOT, bytes, idx_bar(c, b), policy);
    };
    if (GATHER && lane == 0) {
      for (int it0 = 0; it0 < NIDX; ++it0)
#pragma unroll
        for (int k = 0; k < KPW; ++k) {
          const int c = w + NPROD * k;
          const long long ps = (long long)blockIdx.x * NCONS + c + it0 * stride;
          if (ps < npass) issue_idx(c, it0, ps);
        }
    }
    __syncwarp();
    // N ranks, peer transport: the owners' values are in my halo segment of pvec; nothing of pvec is read before this
    if (GATHER && T) warp_wait_fwd(T, st);
    int s2 = 0, sph = 1;                                       // ring stage and the parity of its `empty` barrier to wait for
    for (int it = 0;; ++it) {
      bool any = false;
#pragma unroll
      for (int k = 0; k < KPW; ++k) {
        const int c = w + NPROD * k;
        const long long ps = (long long)blockIdx.x * NCONS + c + (long long)it * stride;
        if (ps >= npass) continue;
        any = true;
        const int b = it & (NIDX - 1);
        const long long e0 = ps * EPP;
        const int ne = (int)((nels - e0) < EPP ? (nels - e0) : EPP);
        const int nw = ne * NTOT;
        mbar_wait(empty_bar(c, s2), sph);                      // the consumer's store out of this stage has been read
        double *rows = reinterpret_cast<double *>(stage0 + (size_t)(c * S + s2) * Cfg::kStageBytes);
        const int *idxbuf = reinterpret_cast<const int *>(idx0 + (size_t)(c * NIDX + b) * Cfg::kIdxBytes);
        if (GATHER) mbar_wait(idx_bar(c, b), (it / NIDX) & 1);
        if (ne == EPP) {
          int idx[KP];
          if (GATHER) {
#pragma unroll
            for (int kp = 0; kp < KP; ++kp) idx[kp] = idxbuf[lane + 32 * kp];
          }
#pragma unroll
          for (int kp = 0; kp < KP; ++kp) {
            const int f = lane + 32 * kp;
            const int d = (ROW == NTOT) ? f : (f / NTOT) * ROW + (f % NTOT);
            cp_async8(smem_u32(rows + d), GATHER ? pvec + idx[kp] : pvec + e0 * NTOT + f);
          }
        } else {
          for (int f = lane; f < nw; f += 32)
            cp_async8(smem_u32(rows + (f / NTOT) * ROW + (f % NTOT)), GATHER ? pvec + idxbuf[f] : pvec + e0 * NTOT + f);
        }
        cp_async_arrive_noinc(full_bar(c, s2));                // fires when this lane's copies have landed
        __syncwarp();
        if (GATHER && lane == 0 && ps + NIDX * stride < npass) {
          fence_proxy_async();                                 // the warp's reads of this index buffer before the async overwrite
          issue_idx(c, b, ps + NIDX * stride);
        }
        if (MID == 1 && lane < 2) {
          // p1210: the pass's Gauss-point state (8 elements x 8 points x 48 B, contiguous in each array) towards L2
          const double *sp = (lane == 0 ? etensor : tensor) + e0 * 48;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(sp), "r"((uint32_t)ne * 384u) : "memory");
        }
        if (GEOM == 2) {
          // this pass's 40 lines of geometric factors towards L2 (the consumer loads them one pass ahead of their use)
          const double2 *gn = reinterpret_cast<const double2 *>(geom + (e0 >> 5) * kGroupGeom) + (int)(e0 & 31);
          asm volatile("prefetch.global.L2 [%0];" ::"l"(gn + lane * 32));
          if (lane < 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(gn + (lane + 32) * 32));
        }
      }
      if (!any) break;
      if (++s2 == S) { s2 = 0; sph ^= 1; }
    }
    cp_async_wait_all();
  } else {
    // ------------------------------------------------ consumer ------------------------------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::kConsRegs));
    const int c = w - NPROD, r = lane >> 2, q = lane & 3;
    // der as B fragments (lane holds B[k = q][n = r]), in registers for the whole kernel:
    //   [s*3+b]                      phase 1, k-step s, column tile b:   der_{g=r}(b, m = 4s+q)
    //   [3 KS1 + (3h+b) NT3 + nt]    phase 3, k-step 3h+b, node tile nt: der_{g=2q+h}(b, m = 8nt+r), 0 beyond the last node
    constexpr int NREG = ALLREG ? NF : 3 * KS1;
    double Bf[NREG];
#pragma unroll
    for (int s3 = 0; s3 < KS1; ++s3)
#pragma unroll
      for (int b = 0; b < 3; ++b) Bf[s3 * 3 + b] = c_tab.der[r * 60 + b * 20 + 4 * s3 + q];
    if (ALLREG) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int nt = 0; nt < NT3; ++nt)
            Bf[(3 * KS1 + (3 * h + b) * NT3 + nt) % NREG] = (8 * nt + r < NOD) ? c_tab.der[(2 * q + h) * 60 + b * 20 + 8 * nt + r] : 0.0;
    }
#pragma unroll
    for (int f = 0; f < NREG; ++f) asm volatile("" : "+d"(Bf[f]));   // opaque: kept in registers, never re-read
    auto frag3 = [&](int f3) -> double { return ALLREG ? Bf[(3 * KS1 + f3) % NREG] : s_frag[f3 * 32 + lane]; };
    // loaded straight into registers as soon as the previous pass no longer needs them (after ITS Gauss-point
    // arithmetic: no extra registers, and phase 3 + the store + the next phase 1 cover the latency): GEOM 2 the lane's
    // two points' factors (jac^-1, det*w), GEOM 0 the lane's coordinates as the A fragments of the Jacobian product
    // (x,y,z of nodes 4s+q of element r)
    constexpr int NPRE = GEOM == 2 ? 20 : 3 * KS1;
    double pre[NPRE];
    auto load_pre = [&](long long ps) {
      const long long e0 = ps * EPP;
      const int ne = (int)((nels - e0) < EPP ? (nels - e0) : EPP);
      if (r < ne) {
        if (GEOM == 2) {
          // word (g*5+j) of element i of a group at [(g*5+j)*32 + i] (16-byte words)
          const double2 *gfl = reinterpret_cast<const double2 *>(geom + (e0 >> 5) * kGroupGeom) + ((int)(e0 & 31) + r);
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < 5; ++j) {
              const double2 v = __ldg(gfl + ((2 * q + h) * 5 + j) * 32);
              pre[h * 10 + 2 * j] = v.x; pre[h * 10 + 2 * j + 1] = v.y;
            }
        } else {
          const double *cr = g_coord + (e0 + r) * NTOT;        // g_coord_pp(nod,3,iel): [b*NOD+m]
#pragma unroll
          for (int s3 = 0; s3 < KS1; ++s3)
#pragma unroll
            for (int b = 0; b < 3; ++b) pre[s3 * 3 + b] = __ldg(cr + b * NOD + 4 * s3 + q);
        }
      } else {
#pragma unroll
        for (int j = 0; j < NPRE; ++j) pre[j] = 0.0;
      }
    };
    long long ps = (long long)blockIdx.x * NCONS + c;
    if (ps < npass) load_pre(ps);
    int s2 = 0, sph = 0, it = 0;
    for (; ps < npass; ps += stride, ++it) {
      const long long e0 = ps * EPP;
      const int ne = (int)((nels - e0) < EPP ? (nels - e0) : EPP);
      double *rows = reinterpret_cast<double *>(stage0 + (size_t)(c * S + s2) * Cfg::kStageBytes);
      double gq0[GEOM == 0 ? 20 : 1];                          // GEOM 0: jac^-1 (9) and det*w of the lane's two points
      if (GEOM == 0) {
        double J[3][3][2];                                     // [b: x,y,z][a][h] = jac(a,b) at point 2q+h
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int a = 0; a < 3; ++a) J[b][a][0] = J[b][a][1] = 0.0;
#pragma unroll
        for (int s3 = 0; s3 < KS1; ++s3)
#pragma unroll
          for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int a = 0; a < 3; ++a) dmma884(J[b][a][0], J[b][a][1], pre[s3 * 3 + b], Bf[s3 * 3 + a]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          double Jm[9];
#pragma unroll
          for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int a = 0; a < 3; ++a) Jm[b * 3 + a] = J[b][a][h];
          const double det = det3(Jm);
          inv3_recip(Jm, det, gq0 + (GEOM == 0 ? 10 * h : 0));
          gq0[GEOM == 0 ? 10 * h + 9 : 0] = det * c_tab.weights[2 * q + h];
        }
      }
      const double *gq = GEOM == 0 ? gq0 : pre;                // [h*10 + j]
      mbar_wait(full_bar(c, s2), sph);                         // the pass's right-hand sides have landed
      // phase 1: H[c2][b][h] = H_{element r, point 2q+h}(b,c2)
      double H[3][3][2];
#pragma unroll
      for (int c2 = 0; c2 < 3; ++c2)
#pragma unroll
        for (int b = 0; b < 3; ++b) H[c2][b][0] = H[c2][b][1] = 0.0;
#pragma unroll
      for (int s3 = 0; s3 < KS1; ++s3) {
        double a3[3];
#pragma unroll
        for (int c2 = 0; c2 < 3; ++c2) a3[c2] = rows[r * ROW + 3 * (4 * s3 + q) + c2];
#pragma unroll
        for (int c2 = 0; c2 < 3; ++c2)
#pragma unroll
          for (int b = 0; b < 3; ++b) dmma884(H[c2][b][0], H[c2][b][1], a3[c2], Bf[s3 * 3 + b]);
      }
      double Tm[2][9];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double Hm[9];
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int c2 = 0; c2 < 3; ++c2) Hm[b * 3 + c2] = H[c2][b][h];
        if (MID == 1) {
          if (r < ne) {
            // the point's 2 x 48 contiguous bytes of state: three 16-byte words each way
            double2 *ep = reinterpret_cast<double2 *>(etensor + ((e0 + r) * 8 + (2 * q + h)) * 6);
            double2 *tp = reinterpret_cast<double2 *>(tensor + ((e0 + r) * 8 + (2 * q + h)) * 6);
            double et[6], te[6];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const double2 a = ep[j], b2 = tp[j];
              et[2 * j] = a.x; et[2 * j + 1] = a.y; te[2 * j] = b2.x; te[2 * j + 1] = b2.y;
            }
            vm_point_mid(Hm, gq + 10 * h, gq[10 * h + 9], et, te, vmp, Tm[h]);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              ep[j] = make_double2(et[2 * j], et[2 * j + 1]);
              tp[j] = make_double2(te[2 * j], te[2 * j + 1]);
            }
          } else {
#pragma unroll
            for (int z = 0; z < 9; ++z) Tm[h][z] = 0.0;
          }
        } else {
          mf_point_mid_iso(Hm, gq + 10 * h, gq[10 * h + 9], Tm[h]);
        }
      }
      if (ps + stride < npass) load_pre(ps + stride);          // the next pass's factors / coordinates
      // phase 3: U[c2][nt][h'] = u_c2(m = 8nt + 2q + h') of element r
      double U[3][NT3][2];
#pragma unroll
      for (int c2 = 0; c2 < 3; ++c2)
#pragma unroll
        for (int nt = 0; nt < NT3; ++nt) U[c2][nt][0] = U[c2][nt][1] = 0.0;
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int c2 = 0; c2 < 3; ++c2)
#pragma unroll
            for (int nt = 0; nt < NT3; ++nt) dmma884(U[c2][nt][0], U[c2][nt][1], Tm[h][b * 3 + c2], frag3((3 * h + b) * NT3 + nt));
      // every lane's A-fragment reads of the stage fed the warp-wide mma above, so they are complete; the barrier
      // states that order for the memory model (and for racecheck): the stage may take the products
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < NT3; ++nt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int m = 8 * nt + 2 * q + hh;
          if (m < NOD) {
#pragma unroll
            for (int c2 = 0; c2 < 3; ++c2) rows[r * ROW + 3 * m + c2] = MID == 1 ? 0.0 - U[c2][nt][hh] : U[c2][nt][hh];
          }
        }
      fence_proxy_async();                                     // my generic stores before the async proxy reads them
      __syncwarp();
      if (lane == 0) {
        if (ROW == NTOT) bulk_s2g(utemp + e0 * NTOT, smem_u32(rows), (uint32_t)(ne * NTOT) * 8);
        else
          for (int e2 = 0; e2 < ne; ++e2) bulk_s2g(utemp + (e0 + e2) * NTOT, smem_u32(rows + e2 * ROW), NTOT * 8);
        bulk_commit();
        if (it > 0) {
          bulk_wait_read_1();                                  // the PREVIOUS pass's store has read its stage
          mbar_arrive(empty_bar(c, s2 == 0 ? S - 1 : s2 - 1));
        }
      }
      if (++s2 == S) { s2 = 0; sph ^= 1; }
    }
    if (lane == 0) bulk_wait_all();                            // shared memory must outlive the stores
  }
}

// ----------------------------------------------------------------------------
// p122 (programs/5th_ed/p122/p122.f90): elasto-plasticity, Mohr-Coulomb, viscoplastic strain method
// ----------------------------------------------------------------------------
// The Gauss-point update of elements_4 (p122.f90:197-231): strain increment from the displacement increment,
// stress = dee*(eps - evpt) + tensor, invar / mocouf; on yield mocouq + formm -> flow -> evp = flow*stress*dt,
// evpt += evp, devp = dee*evp (on the last plastic iteration devp = stress and tensor = stress); body loads
// bload = sum_gp bee^T devp det w.  Same summation orders as orc_p122_elements (pf_oracle.c): q / r / k ascending,
// Gauss points ascending, separate multiply and add; products with a structural zero of bee are left out (they add
// +-0.0).  The angle constants (sin / cos of phi and psi in radians) come from the host's libm, so that the only
// transcendental functions evaluated here are asin (invar) and sin / cos / tan of the Lode angle.
struct PlasticParams {
  double snph, csph, cohesion, snps, dt;
};
// invar, nst = 6 (new_library.f90:1893-1912)
__device__ __forceinline__ void invar6(const double *s, double &sigm, double &dsbar, double &theta) {
  const double sq3 = sqrt(3.0);
  sigm = (s[0] + s[1] + s[2]) / 3.0;
  const double d2 = ((s[0] - s[1]) * (s[0] - s[1]) + (s[1] - s[2]) * (s[1] - s[2]) + (s[2] - s[0]) * (s[2] - s[0])) / 6.0 +
                    s[3] * s[3] + s[4] * s[4] + s[5] * s[5];
  const double ds1 = s[0] - sigm, ds2 = s[1] - sigm, ds3 = s[2] - sigm;
  const double d3 = ds1 * ds2 * ds3 - ds1 * s[4] * s[4] - ds2 * s[5] * s[5] - ds3 * s[3] * s[3] + 2.0 * s[3] * s[4] * s[5];
  dsbar = sq3 * sqrt(d2);
  if (dsbar < 1e-10) theta = 0.0;
  else {
    const double r = sqrt(d2);
    double sine = -3.0 * sq3 * d3 / (2.0 * r * r * r);
    if (sine > 1.0) sine = 1.0;
    if (sine < -1.0) sine = -1.0;
    theta = asin(sine) / 3.0;
  }
}
// mocouf (new_library.f90:2364-2417)
__device__ __forceinline__ double mocouf(const PlasticParams &P, double sigm, double dsbar, double theta) {
  const double csth = cos(theta), snth = sin(theta);
  return P.snph * sigm + dsbar * (csth / sqrt(3.0) - snth * P.snph / 3.0) - P.cohesion * P.csph;
}
// mocouq (new_library.f90:2423-2490)
__device__ __forceinline__ void mocouq(const PlasticParams &P, double dsbar, double theta, double &dq1, double &dq2, double &dq3) {
  const double snth = sin(theta), snps = P.snps, sq3 = sqrt(3.0);
  dq1 = snps;
  if (fabs(snth) > 0.49) {
    const double c1 = snth < 0.0 ? -1.0 : 1.0;
    dq2 = (sq3 * 0.5 - c1 * snps * 0.5 / sq3) * sq3 * 0.5 / dsbar;
    dq3 = 0.0;
  } else {
    const double csth = cos(theta), cs3th = cos(3.0 * theta), tn3th = tan(3.0 * theta), tnth = snth / csth;
    dq2 = sq3 * csth / dsbar * ((1.0 + tnth * tn3th) + snps * (tn3th - tnth) / sq3) * 0.5;
    dq3 = 0.5 * 3.0 * (sq3 * snth + snps * csth) / (cs3th * dsbar * dsbar);
  }
}
// formm, nst = 6 (new_library.f90:145-194); m(i,j) at [j*6+i]
__device__ __forceinline__ void formm6(const double *st, double *m1, double *m2, double *m3) {
  const double sx = st[0], sy = st[1], sz = st[2], txy = st[3], tyz = st[4], tzx = st[5];
  const double sigm = (sx + sy + sz) / 3.0, dx = sx - sigm, dy = sy - sigm, dz = sz - sigm;
#pragma unroll
  for (int q = 0; q < 36; ++q) { m1[q] = 0.0; m2[q] = 0.0; m3[q] = 0.0; }
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
#undef M
  for (int q = 0; q < 36; ++q) { m1[q] = m1[q] / 3.0; m2[q] = m2[q] / 3.0; m3[q] = m3[q] / 3.0; }
}

// One CTA of 64 threads per element: (1) jac / det / inverse / deriv of all Gauss points (as k_form_km_tiled);
// (2) thread ig < nip updates Gauss point ig (strain, stress, yield function, flow, evpt / tensor) and leaves devp
// and the yield flag in shared memory; (3) thread q < ntot adds the points' contributions to bload(q) in point order.
// loads_ext: the displacement increment, slot-indexed (gather through ggl); evpt / tensor (6,nip,nels).
template <int NOD>
__global__ void __launch_bounds__(64)
k_p122_elements(const double *__restrict__ g_coord, const int *__restrict__ ggl, const double *__restrict__ loads_ext,
                double *__restrict__ evpt, double *__restrict__ tensor, double *__restrict__ utemp, long long nels,
                PlasticParams P, int last) {
  constexpr int NTOT = 3 * NOD, MAXIP = 8, THREADS = 64;
  static_assert(NTOT <= THREADS, "one thread per element freedom");
  __shared__ double s_coord[NOD * 3], s_jac[MAXIP * 9], s_det[MAXIP], s_deriv[MAXIP * NOD * 3], s_eld[NTOT];
  __shared__ double s_devp[MAXIP * 6];
  __shared__ int s_yield[MAXIP];
  const int nip = c_tab.nip, t = threadIdx.x;
  for (long long e = blockIdx.x; e < nels; e += gridDim.x) {
    __syncthreads();
    for (int q = t; q < NOD * 3; q += THREADS) s_coord[q] = g_coord[e * NOD * 3 + q];
    if (t < NTOT) s_eld[t] = loads_ext[ggl[e * NTOT + t]];           // gather(loads_pp,pmul_pp): slot 0 holds 0.0
    __syncthreads();
    for (int q = t; q < nip * 9; q += THREADS) {                      // jac = MATMUL(der,coord), every point
      const int ig = q / 9, r = q - 9 * ig, a = r % 3, b = r / 3;
      const double *der = c_tab.der + ig * 60;
      double sum = 0.0;
#pragma unroll
      for (int m = 0; m < NOD; ++m) sum = sum + der[a * 20 + m] * s_coord[b * NOD + m];
      s_jac[ig * 9 + b * 3 + a] = sum;
    }
    __syncthreads();
    if (t < nip) {
      double jac[9], inv[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) jac[q] = s_jac[t * 9 + q];
      const double det = det3(jac);
      inv3(jac, det, inv);
#pragma unroll
      for (int q = 0; q < 9; ++q) s_jac[t * 9 + q] = inv[q];
      s_det[t] = det;
    }
    __syncthreads();
    for (int q = t; q < nip * NOD * 3; q += THREADS) {                // deriv = MATMUL(jac^-1,der)
      const int ig = q / (NOD * 3), r = q - ig * (NOD * 3), m = r / 3, a = r - 3 * m;
      const double *der = c_tab.der + ig * 60, *inv = s_jac + ig * 9;
      double sum = 0.0;
#pragma unroll
      for (int b = 0; b < 3; ++b) sum = sum + inv[b * 3 + a] * der[b * 20 + m];
      s_deriv[q] = sum;
    }
    __syncthreads();
    if (t < nip) {
      const double *dv = s_deriv + t * (NOD * 3);
      double *ev = evpt + (e * nip + t) * 6, *te = tensor + (e * nip + t) * 6;
      double eps[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, stress[6], devp[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      for (int m = 0; m < NOD; ++m) {                                 // eps = MATMUL(bee,eld), q ascending
        const double x = dv[m * 3], y = dv[m * 3 + 1], z = dv[m * 3 + 2];
        const double e0 = s_eld[3 * m], e1 = s_eld[3 * m + 1], e2 = s_eld[3 * m + 2];
        eps[0] = eps[0] + x * e0; eps[3] = eps[3] + y * e0; eps[5] = eps[5] + z * e0;     // column 3m
        eps[1] = eps[1] + y * e1; eps[3] = eps[3] + x * e1; eps[4] = eps[4] + z * e1;     // column 3m+1
        eps[2] = eps[2] + z * e2; eps[4] = eps[4] + y * e2; eps[5] = eps[5] + x * e2;     // column 3m+2
      }
#pragma unroll
      for (int r = 0; r < 6; ++r) eps[r] = eps[r] - ev[r];
#pragma unroll
      for (int r = 0; r < 6; ++r) {                                   // sigma = MATMUL(dee,eps); stress = sigma + tensor
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < 6; ++q) sum = sum + c_tab.dee[q * 6 + r] * eps[q];
        stress[r] = sum + te[r];
      }
      double sigm, dsbar, theta;
      invar6(stress, sigm, dsbar, theta);
      const double f = mocouf(P, sigm, dsbar, theta);
      if (last) {
#pragma unroll
        for (int r = 0; r < 6; ++r) devp[r] = stress[r];
      } else if (f >= 0.0) {
        double dq1, dq2, dq3, m1[36], m2[36], m3[36], evp[6];
        mocouq(P, dsbar, theta, dq1, dq2, dq3);
        formm6(stress, m1, m2, m3);
        for (int r = 0; r < 6; ++r) {
          double sum = 0.0;
          for (int q = 0; q < 6; ++q) {
            const double flow = f * (m1[q * 6 + r] * dq1 + m2[q * 6 + r] * dq2 + m3[q * 6 + r] * dq3);
            sum = sum + flow * stress[q];
          }
          evp[r] = sum * P.dt;
          ev[r] = ev[r] + evp[r];
        }
        for (int r = 0; r < 6; ++r) {
          double sum = 0.0;
          for (int q = 0; q < 6; ++q) sum = sum + c_tab.dee[q * 6 + r] * evp[q];
          devp[r] = sum;
        }
      }
#pragma unroll
      for (int r = 0; r < 6; ++r) s_devp[t * 6 + r] = devp[r];
      s_yield[t] = f >= 0.0;
      if (last) {
#pragma unroll
        for (int r = 0; r < 6; ++r) te[r] = stress[r];
      }
    }
    __syncthreads();
    if (t < NTOT) {                                                   // bload = bload + MATMUL(TRANSPOSE(bee),devp)*det*w
      const int m = t / 3, comp = t - 3 * m;
      double bl = 0.0;
      for (int ig = 0; ig < nip; ++ig) {
        if (!s_yield[ig]) continue;
        const double *dv = s_deriv + ig * (NOD * 3), *dp = s_devp + ig * 6;
        const double x = dv[m * 3], y = dv[m * 3 + 1], z = dv[m * 3 + 2];
        double sum = 0.0;
        if (comp == 0) { sum = sum + x * dp[0]; sum = sum + y * dp[3]; sum = sum + z * dp[5]; }
        else if (comp == 1) { sum = sum + y * dp[1]; sum = sum + x * dp[3]; sum = sum + z * dp[4]; }
        else { sum = sum + z * dp[2]; sum = sum + y * dp[4]; sum = sum + x * dp[5]; }
        bl = bl + sum * s_det[ig] * c_tab.weights[ig];
      }
      utemp[e * NTOT + t] = bl;
    }
  }
}

// ----------------------------------------------------------------------------
// p1210 (programs/5th_ed/p1210/p1210.f90): forced vibration of an elastic-plastic (von Mises) solid, lumped mass,
// explicit integration.  No PCG: a time step is gather -> Gauss-point stress update -> scatter -> three vector updates.
// ----------------------------------------------------------------------------
// elements_1 (p1210.f90:93-104): one thread per element, emm(ntot) -> utemp (mm_tmp)
__global__ void k_p1210_mass(const double *__restrict__ g_coord, double *__restrict__ utemp, long long nels, double rho) {
  constexpr int NOD = 20;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nels; e += (long long)gridDim.x * blockDim.x) {
    const double *c = g_coord + e * NOD * 3;
    double volume = 0.0;
    for (int ig = 0; ig < 8; ++ig) {
      const double *der = c_tab.der + ig * 60;
      double jac[9];
      for (int b = 0; b < 3; ++b)
        for (int a = 0; a < 3; ++a) {
          double sum = 0.0;
          for (int m = 0; m < NOD; ++m) sum = sum + der[a * 20 + m] * c[b * NOD + m];
          jac[b * 3 + a] = sum;
        }
      volume = volume + det3(jac) * c_tab.weights[ig] * rho;
    }
    const double mid = volume / 13.0, corner = mid * .125;
    double *o = utemp + e * 60;
    for (int m = 0; m < NOD; ++m) {
      const bool is_corner = (m < 8 || m >= 12) && (m % 2 == 0);     // freedoms 1:19:6 .. 3:21:6 and 37:55:6 .. 39:57:6
      o[3 * m] = o[3 * m + 1] = o[3 * m + 2] = is_corner ? corner : mid;
    }
  }
}
// elements_2 (p1210.f90:120-147): geometry of every Gauss point as in k_p122_elements; thread ig updates point ig (strain
// from the gathered displacements minus etensor, elastic trial stress, on yield the scaled-back stress -> vmpl ->
// dee - fac*pl, stress, tensor / etensor update); thread q adds the points' sigma*bee*det*w to bload(q) in point order;
// utemp = 0 - bload.  Summation orders as orc_p1210_elements (products with a structural zero of bee left out).
template <int NOD>
__global__ void __launch_bounds__(64)
k_p1210_elements(const double *__restrict__ g_coord, const int *__restrict__ ggl, const double *__restrict__ x_ext,
                 double *__restrict__ etensor, double *__restrict__ tensor, double *__restrict__ utemp, long long nels, VmParams P) {
  constexpr int NTOT = 3 * NOD, MAXIP = 8, THREADS = 64;
  static_assert(NTOT <= THREADS, "one thread per element freedom");
  __shared__ double s_coord[NOD * 3], s_jac[MAXIP * 9], s_det[MAXIP], s_deriv[MAXIP * NOD * 3], s_eld[NTOT], s_sig[MAXIP * 6];
  const int nip = c_tab.nip, t = threadIdx.x;
  for (long long e = blockIdx.x; e < nels; e += gridDim.x) {
    __syncthreads();
    for (int q = t; q < NOD * 3; q += THREADS) s_coord[q] = g_coord[e * NOD * 3 + q];
    if (t < NTOT) s_eld[t] = x_ext[ggl[e * NTOT + t]];                // gather(x1_pp,pmul_pp): slot 0 holds 0.0
    __syncthreads();
    for (int q = t; q < nip * 9; q += THREADS) {                      // jac = MATMUL(der,coord), every point
      const int ig = q / 9, r = q - 9 * ig, a = r % 3, b = r / 3;
      const double *der = c_tab.der + ig * 60;
      double sum = 0.0;
#pragma unroll
      for (int m = 0; m < NOD; ++m) sum = sum + der[a * 20 + m] * s_coord[b * NOD + m];
      s_jac[ig * 9 + b * 3 + a] = sum;
    }
    __syncthreads();
    if (t < nip) {
      double jac[9], inv[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) jac[q] = s_jac[t * 9 + q];
      const double det = det3(jac);
      inv3(jac, det, inv);
#pragma unroll
      for (int q = 0; q < 9; ++q) s_jac[t * 9 + q] = inv[q];
      s_det[t] = det;
    }
    __syncthreads();
    for (int q = t; q < nip * NOD * 3; q += THREADS) {                // deriv = MATMUL(jac^-1,der)
      const int ig = q / (NOD * 3), r = q - ig * (NOD * 3), m = r / 3, a = r - 3 * m;
      const double *der = c_tab.der + ig * 60, *inv = s_jac + ig * 9;
      double sum = 0.0;
#pragma unroll
      for (int b = 0; b < 3; ++b) sum = sum + inv[b * 3 + a] * der[b * 20 + m];
      s_deriv[q] = sum;
    }
    __syncthreads();
    if (t < nip) {
      const double *dv = s_deriv + t * (NOD * 3);
      double *et = etensor + (e * nip + t) * 6, *te = tensor + (e * nip + t) * 6;
      double eps[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, sigma[6], stressv[6], ten[6], dee[36];
      for (int m = 0; m < NOD; ++m) {                                 // eps = MATMUL(bee,eld), q ascending
        const double x = dv[m * 3], y = dv[m * 3 + 1], z = dv[m * 3 + 2];
        const double e0 = s_eld[3 * m], e1 = s_eld[3 * m + 1], e2 = s_eld[3 * m + 2];
        eps[0] = eps[0] + x * e0; eps[3] = eps[3] + y * e0; eps[5] = eps[5] + z * e0;     // column 3m
        eps[1] = eps[1] + y * e1; eps[3] = eps[3] + x * e1; eps[4] = eps[4] + z * e1;     // column 3m+1
        eps[2] = eps[2] + z * e2; eps[4] = eps[4] + y * e2; eps[5] = eps[5] + x * e2;     // column 3m+2
      }
#pragma unroll
      for (int r = 0; r < 6; ++r) { eps[r] = eps[r] - et[r]; ten[r] = te[r]; }
#pragma unroll
      for (int q = 0; q < 36; ++q) dee[q] = c_tab.dee[q];
#pragma unroll
      for (int r = 0; r < 6; ++r) {                                   // sigma = MATMUL(dee,eps); stressv = sigma + tensor
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < 6; ++q) sum = sum + dee[q * 6 + r] * eps[q];
        sigma[r] = sum;
        stressv[r] = sum + ten[r];
      }
      const double fnew = vm_dsbar(stressv) - P.sbary;
      if (fnew >= 0.0) {                                              // yield is violated
        const double f = vm_dsbar(ten) - P.sbary, fac = fnew / (fnew - f);
#pragma unroll
        for (int r = 0; r < 6; ++r) stressv[r] = ten[r] + (1.0 - fac) * sigma[r];
        // vmpl (new_library.f90:2262-2284)
        const double sx = stressv[0], sy = stressv[1], sz = stressv[2], txy = stressv[3], tyz = stressv[4], tzx = stressv[5];
        const double dsb = sqrt((sx - sy) * (sx - sy) + (sy - sz) * (sy - sz) + (sz - sx) * (sz - sx) + 6.0 * (txy * txy) +
                                6.0 * (tyz * tyz) + 6.0 * (tzx * tzx)) / sqrt(2.0);
        const double ee = 1.5 * P.e / ((1.0 + P.v) * dsb * dsb);
        const double term[6] = {(2.0 * sx - sy - sz) / 3.0, (2.0 * sy - sz - sx) / 3.0, (2.0 * sz - sx - sy) / 3.0, txy, tyz, tzx};
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j < 6; ++j) dee[j * 6 + i] = dee[j * 6 + i] - fac * (term[i] * term[j] * ee);
      }
#pragma unroll
      for (int r = 0; r < 6; ++r) {                                   // sigma = MATMUL(dee,eps) + tensor
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < 6; ++q) sum = sum + dee[q * 6 + r] * eps[q];
        sigma[r] = sum + ten[r];
      }
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        s_sig[t * 6 + r] = sigma[r];
        te[r] = sigma[r];
        et[r] = et[r] + eps[r];
      }
    }
    __syncthreads();
    if (t < NTOT) {                                                   // bload = bload + MATMUL(sigma,bee)*det*w
      const int m = t / 3, comp = t - 3 * m;
      double bl = 0.0;
      for (int ig = 0; ig < nip; ++ig) {
        const double *dv = s_deriv + ig * (NOD * 3), *sg = s_sig + ig * 6;
        const double x = dv[m * 3], y = dv[m * 3 + 1], z = dv[m * 3 + 2];
        double sum = 0.0;
        if (comp == 0) { sum = sum + sg[0] * x; sum = sum + sg[3] * y; sum = sum + sg[5] * z; }
        else if (comp == 1) { sum = sum + sg[1] * y; sum = sum + sg[3] * x; sum = sum + sg[4] * z; }
        else { sum = sum + sg[2] * z; sum = sum + sg[4] * y; sum = sum + sg[5] * x; }
        bl = bl + sum * s_det[ig] * c_tab.weights[ig];
      }
      utemp[e * NTOT + t] = 0.0 - bl;
    }
  }
}
// x1 = x1 + dtim*d1x1 + 0.5*dtim**2*d2x1   (p1210.f90:117); x1 lives in the owned part of p_ext
__global__ void k_p1210_predict(double *__restrict__ x1, const double *__restrict__ d1, const double *__restrict__ d2, double dtim,
                                long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const double c = 0.5 * (dtim * dtim);
  for (; i < n; i += stride) x1[i] = x1[i] + (dtim * d1[i]) + (c * d2[i]);
}
// bdylds = (bdylds + fext*pload)/mm; d1x1 = d1x1 + (d2x1 + bdylds)*.5*dtim; d2x1 = bdylds   (p1210.f90:148-150)
__global__ void k_p1210_update(const double *__restrict__ bdy, const double *__restrict__ fext, const double *__restrict__ mm,
                               double *__restrict__ d1, double *__restrict__ d2, double pload, double dtim, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    double b = bdy[i] + fext[i] * pload;
    b = b / mm[i];
    d1[i] = d1[i] + (d2[i] + b) * .5 * dtim;
    d2[i] = b;
  }
}

// loads = ld0*q (or 0) [+ bdylds]   (p122.f90:127-137)
__global__ void k_plastic_loads(double *__restrict__ loads, const double *__restrict__ ld0, const double *__restrict__ bdylds,
                                double q, int add_bdy, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    double v = ld0 ? ld0[i] * q : 0.0;
    if (add_bdy) v = v + bdylds[i];
    loads[i] = v;
  }
}
// fixed freedoms of p122: mode 0  loads(j) = store*valf*qinc (p122.f90:121-126); mode 1  loads(j) = 0 (:133-137).
// dst is slot-indexed.
__global__ void k_plastic_fixed(const int *__restrict__ fix_slot, const double *__restrict__ store, const double *__restrict__ valf,
                                double *__restrict__ dst, double q, int n, int mode) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dst[fix_slot[i]] = mode == 0 ? store[i] * valf[i] * q : 0.0;
}
// c = a - b ; c = a + b
__global__ void k_vsub(double *__restrict__ c, const double *__restrict__ a, const double *__restrict__ b, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) c[i] = a[i] - b[i];
}
__global__ void k_vadd(double *__restrict__ c, const double *a, const double *__restrict__ b, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) c[i] = a[i] + b[i];
}
// checon_par (maths.f90:1048-1061) outside the solver: max|loads|, max|loads - oldlds| of this rank into st->loc[1],
// loc[2]; oldlds = loads
__global__ void __launch_bounds__(kRedThreads)
k_checon(const double *__restrict__ loads, double *__restrict__ oldlds, long long n, double *part, State *st) {
  __shared__ double sh[8];
  __shared__ int flag;
  const long long nchunks = (n + kChunk - 1) / kChunk;
  for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
    double ml = 0.0, md = 0.0;
    for (int k = 0; k < 8; ++k) {
      const long long i = c * kChunk + 256 * k + threadIdx.x;
      if (i < n) {
        const double v = loads[i], o = oldlds[i];
        ml = fmax(ml, fabs(v)); md = fmax(md, fabs(v - o));
        oldlds[i] = v;
      }
    }
    const double a = block_max(ml, sh);
    const double b = block_max(md, sh);
    if (threadIdx.x == 0) { part[c] = a; part[nchunks + c] = b; }
  }
  if (last_block(&st->ticket[0], &flag)) {
    const double m1 = final_max(part, nchunks, sh);
    const double m2 = final_max(part + nchunks, nchunks, sh);
    if (threadIdx.x == 0) { st->loc[0] = 0.0; st->loc[1] = m1; st->loc[2] = m2; st->loc[3] = 0.0; }
  }
}

// ----------------------------------------------------------------------------
// p129 (programs/5th_ed/p129/p129.f90): forced vibration, implicit theta method, consistent mass
// ----------------------------------------------------------------------------
// elements_2, mass part (p129.f90:90-93): emm += ecmat(fun)*det*w*rho per Gauss point, ecmat = MATMUL(nt,tn)
// (new_library.f90:1536-1563), i.e. emm(3i+j,3i'+j') += fun(i)*fun(i')*det*w*rho for j == j' and +0.0 otherwise.
// One CTA of 128 threads per element, every thread a strided share of the ntot x ntot entries.
template <int NOD>
__global__ void __launch_bounds__(128)
k_form_mass(const double *__restrict__ g_coord, double *__restrict__ mm, long long nels, double rho) {
  constexpr int NTOT = 3 * NOD, NENT = NTOT * NTOT, THREADS = 128, PER = (NENT + THREADS - 1) / THREADS;
  __shared__ double s_coord[NOD * 3], s_jac[9], s_deriv[NOD * 3];
  for (long long e = blockIdx.x; e < nels; e += gridDim.x) {
    __syncthreads();
    for (int q = threadIdx.x; q < NOD * 3; q += THREADS) s_coord[q] = g_coord[e * NOD * 3 + q];
    double acc[PER];
#pragma unroll
    for (int n = 0; n < PER; ++n) acc[n] = 0.0;
    __syncthreads();
    for (int ig = 0; ig < c_tab.nip; ++ig) {
      const double det = gauss_point<NOD>(ig, s_coord, s_jac, s_deriv);
      const double wt = c_tab.weights[ig];
      const double *fun = NOD == 20 ? c_tab.fun20 + ig * 20 : c_tab.fun + ig * 8;
#pragma unroll
      for (int n = 0; n < PER; ++n) {
        const int idx = threadIdx.x + n * THREADS;
        if (idx < NENT) {
          const int b = idx / NTOT, a = idx - b * NTOT;
          const double ecm = (a % 3 == b % 3) ? fun[a / 3] * fun[b / 3] : 0.0;
          acc[n] = acc[n] + ecm * det * wt * rho;
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int n = 0; n < PER; ++n) {
      const int idx = threadIdx.x + n * THREADS;
      if (idx < NENT) mm[e * (long long)NENT + idx] = acc[n];
    }
  }
}
// out = a*ca + b*cb, entry by entry (p129.f90:114 temp_pp = store_km_pp*c2 + store_mm_pp*c3, :128 the PCG matrix);
// out may be b
__global__ void k_lincomb(double *out, const double *__restrict__ a, double ca, const double *b, double cb, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = a[i] * ca + b[i] * cb;
}
__global__ void k_divide(double *__restrict__ out, const double *__restrict__ a, double s, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = a[i] / s;
}
// loads = u + vu + fext*factor   (p129.f90:124-126; factor = theta*dtim*cos(omega t) + c1*cos(omega (t - dtim)))
__global__ void k_dyn_rhs(double *__restrict__ r, const double *__restrict__ u, const double *__restrict__ vu,
                          const double *__restrict__ fext, double factor, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) r[i] = u[i] + vu[i] + fext[i] * factor;
}
// x1 = xnew; d1x1 = (x1-x0)/(theta*dtim) - d1x0*(1-theta)/theta; d2x1 = (d1x1-d1x0)/(theta*dtim) - d2x0*(1-theta)/theta;
// x0 = x1; d1x0 = d1x1; d2x0 = d2x1   (p129.f90:146-149)
__global__ void k_dyn_update(const double *__restrict__ x1, double *__restrict__ x0, double *__restrict__ d1x0,
                             double *__restrict__ d2x0, double theta, double dtim, long long n) {
  const double td = theta * dtim, omt = 1.0 - theta;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const double a = x1[i], v0 = d1x0[i];
    const double v1 = (a - x0[i]) / td - v0 * omt / theta;
    const double w1 = (v1 - v0) / td - d2x0[i] * omt / theta;
    x0[i] = a; d1x0[i] = v1; d2x0[i] = w1;
  }
}

// elements_1 of p123.f90:71-84; 8-node bricks (or 4-node tetrahedra), NOD*NOD threads = one per entry
template <int NOD>
__global__ void __launch_bounds__(NOD * NOD < 32 ? 32 : NOD * NOD)
k_form_kc_laplace(const double *__restrict__ g_coord, double *__restrict__ kc, long long nels, int packed) {
  constexpr int NENT = NOD * NOD;
  __shared__ double s_coord[NOD * 3], s_jac[9], s_deriv[NOD * 3];
  const int j = threadIdx.x / NOD, i = threadIdx.x % NOD;
  for (long long e = blockIdx.x; e < nels; e += gridDim.x) {
    __syncthreads();
    if (threadIdx.x < NOD * 3) s_coord[threadIdx.x] = g_coord[e * NOD * 3 + threadIdx.x];
    __syncthreads();
    double kx = 0.0, ky = 0.0, kz = 0.0;
    for (int ig = 0; ig < c_tab.nip; ++ig) {
      const double det = gauss_point<NOD>(ig, s_coord, s_jac, s_deriv);
      const double wt = c_tab.weights[ig];
      if (threadIdx.x < NENT) {
        kx = kx + s_deriv[i * 3 + 0] * s_deriv[j * 3 + 0] * det * wt;
        ky = ky + s_deriv[i * 3 + 1] * s_deriv[j * 3 + 1] * det * wt;
        kz = kz + s_deriv[i * 3 + 2] * s_deriv[j * 3 + 2] * det * wt;
      }
      __syncthreads();
    }
    if (threadIdx.x < NENT) {
      const double val = kx * c_tab.kxyz[0] + ky * c_tab.kxyz[1] + kz * c_tab.kxyz[2];
      if (!packed) kc[e * NENT + threadIdx.x] = val;
      else if (i >= j) kc[e * SymCfg<NOD>::kPacked + SymCfg<NOD>::coloff(j) + (i - j)] = val;
    }
  }
}

// elements_3 / gauss_pts of p124.f90:81-95; 8-node bricks, 64 threads = one per entry:
//   kc += MATMUL(MATMUL(TRANSPOSE(deriv),kay),deriv)*det*w ;  pm += fun fun^T *det*w*rho*cp
//   storka = pm + kc*theta*dtim ; storkb = pm - kc*(1-theta)*dtim
__global__ void __launch_bounds__(64)
k_form_k_transient(const double *__restrict__ g_coord, double *__restrict__ ka, double *__restrict__ kb, long long nels) {
  constexpr int NOD = 8;
  __shared__ double s_coord[NOD * 3], s_jac[9], s_deriv[NOD * 3];
  const int j = threadIdx.x / 8, i = threadIdx.x % 8;
  const double rho = c_tab.trans[0], cp = c_tab.trans[1], theta = c_tab.trans[2], dtim = c_tab.trans[3];
  const double omt = 1.0 - theta;
  for (long long e = blockIdx.x; e < nels; e += gridDim.x) {
    __syncthreads();
    if (threadIdx.x < NOD * 3) s_coord[threadIdx.x] = g_coord[e * NOD * 3 + threadIdx.x];
    __syncthreads();
    double kc = 0.0, pm = 0.0;
    for (int ig = 0; ig < c_tab.nip; ++ig) {
      const double det = gauss_point<NOD>(ig, s_coord, s_jac, s_deriv);
      const double wt = c_tab.weights[ig];
      double s = 0.0;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        // t1(i,b) = sum_a deriv(a,i)*kay(a,b), kay diagonal: the full a-ascending sum from 0.0
        double t1 = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) t1 = t1 + s_deriv[i * 3 + a] * (a == b ? c_tab.kxyz[b] : 0.0);
        s = s + t1 * s_deriv[j * 3 + b];
      }
      kc = kc + s * det * wt;
      double f = 0.0;
      f = f + c_tab.fun[ig * 8 + i] * c_tab.fun[ig * 8 + j];
      pm = pm + f * det * wt * rho * cp;
      __syncthreads();
    }
    ka[e * 64 + threadIdx.x] = pm + kc * theta * dtim;
    kb[e * 64 + threadIdx.x] = pm - kc * omt * dtim;
  }
}

// elements_1 of p125.f90:66-79 (explicit transient conduction); 8-node bricks, 64 threads = one per entry:
//   kc, pm as in p124 (no rho*cp); mass(i) = SUM(pm(i,:)), j ascending; store_pm = diag(mass) - kc*dtim;
//   mass_out(i,iel) = mass(i) feeds globma_pp (p125.f90:78,82)
__global__ void __launch_bounds__(64)
k_form_k_explicit(const double *__restrict__ g_coord, double *__restrict__ store_pm, double *__restrict__ mass_out, long long nels) {
  constexpr int NOD = 8;
  __shared__ double s_coord[NOD * 3], s_jac[9], s_deriv[NOD * 3], s_pm[64], s_mass[8];
  const int j = threadIdx.x / 8, i = threadIdx.x % 8;
  const double dtim = c_tab.trans[3];
  for (long long e = blockIdx.x; e < nels; e += gridDim.x) {
    __syncthreads();
    if (threadIdx.x < NOD * 3) s_coord[threadIdx.x] = g_coord[e * NOD * 3 + threadIdx.x];
    __syncthreads();
    double kc = 0.0, pm = 0.0;
    for (int ig = 0; ig < c_tab.nip; ++ig) {
      const double det = gauss_point<NOD>(ig, s_coord, s_jac, s_deriv);
      const double wt = c_tab.weights[ig];
      double s = 0.0;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        double t1 = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) t1 = t1 + s_deriv[i * 3 + a] * (a == b ? c_tab.kxyz[b] : 0.0);
        s = s + t1 * s_deriv[j * 3 + b];
      }
      kc = kc + s * det * wt;
      double f = 0.0;
      f = f + c_tab.fun[ig * 8 + i] * c_tab.fun[ig * 8 + j];
      pm = pm + f * det * wt;
      __syncthreads();
    }
    s_pm[threadIdx.x] = pm;                       // pm(i,j) at [j*8+i]
    __syncthreads();
    if (threadIdx.x < 8) {
      double m = 0.0;
#pragma unroll
      for (int c = 0; c < 8; ++c) m = m + s_pm[c * 8 + threadIdx.x];
      s_mass[threadIdx.x] = m;
      mass_out[e * 8 + threadIdx.x] = m;
    }
    __syncthreads();
    store_pm[e * 64 + threadIdx.x] = (i == j ? s_mass[i] : 0.0) - kc * dtim;
  }
}

// loads_pp = newlo_pp*globma_pp (p125.f90:99)
__global__ void k_scale(double *__restrict__ dst, const double *__restrict__ a, const double *__restrict__ b, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = a[i] * b[i];
}

// scatter tables on the device (pf_setup_mesh): positions 0..n-1, and csr_ptr from the sorted slot numbers:
// csr_ptr[s] = (number of sorted keys < s) - (number of keys == 0) for s >= 1, csr_ptr[0] = 0 -- slot 0, the
// restrained dump slot, keeps no contributions
__global__ void k_iota(unsigned int *__restrict__ v, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) v[i] = (unsigned int)i;
}
__device__ __forceinline__ long long lower_bound_i32(const int *__restrict__ keys, long long n, int s) {
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (keys[mid] < s) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__global__ void k_csr_ptr(const int *__restrict__ sorted_keys, long long n, unsigned int *__restrict__ ptr, long long nslots) {
  const long long n0 = lower_bound_i32(sorted_keys, n, 1);
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; s <= nslots; s += stride)
    ptr[s] = s == 0 ? 0u : (unsigned int)((s == nslots ? n : lower_bound_i32(sorted_keys, n, (int)s)) - n0);
}

// p124 time stepping, right-hand side of one step (p124.f90:143-200):
//   loads = (loaded freedoms, or 0) + u ;  r = loads - r0 with r0 = +0.0 off the fixed freedoms
__global__ void k_transient_rhs(double *__restrict__ r, const double *__restrict__ loads, const double *__restrict__ u, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) r[i] = ((loads ? loads[i] : 0.0) + u[i]) - 0.0;
}
__global__ void k_fill(double *__restrict__ v, double val, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) v[i] = val;
}
// fixed freedoms of p124: mode 0  x(l) = val_f (p124.f90:170-173); mode 1  u(l) = store*val_f (:156-159);
// mode 2  r(l) = loads(l) - store*val_f (:193-198, dst holds loads).  dst is slot-indexed (slot 0 = dump).
__global__ void k_fixed_transient(const int *__restrict__ fix_slot, const double *__restrict__ store,
                                  const double *__restrict__ val_f, double *__restrict__ dst, int n, int mode) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = fix_slot[i];
  if (mode == 0) dst[s] = val_f[i];
  else if (mode == 1) dst[s] = store[i] * val_f[i];
  else dst[s] = dst[s] - store[i] * val_f[i];
}

// full (ntot,ntot) column-major element matrices <-> packed lower triangles (pf_set_storkm / pf_get_storkm
// on the symmetric layout); unpacking mirrors the lower triangle
__global__ void k_pack_lower(const double *__restrict__ full, double *__restrict__ packed, long long nel, int ntot) {
  const int P = ntot * (ntot + 1) / 2;
  const long long total = nel * (long long)ntot * ntot;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const long long e = q / (ntot * ntot);
    const int r = (int)(q - e * ntot * ntot), j = r / ntot, i = r - j * ntot;
    if (i >= j) packed[e * P + j * ntot - j * (j - 1) / 2 + (i - j)] = full[q];
  }
}
__global__ void k_unpack_lower(const double *__restrict__ packed, double *__restrict__ full, long long nel, int ntot) {
  const int P = ntot * (ntot + 1) / 2;
  const long long total = nel * (long long)ntot * ntot;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const long long e = q / (ntot * ntot);
    const int r = (int)(q - e * ntot * ntot), j = r / ntot, i = r - j * ntot;
    const int hi = i >= j ? i : j, lo = i >= j ? j : i;
    full[q] = packed[e * P + lo * ntot - lo * (lo - 1) / 2 + (hi - lo)];
  }
}

// Read-only HBM stream probe: the k_matvec data path (persistent CTAs, ring of 1-D bulk copies, evict-first)
// without the arithmetic.  MEASURED_PEAKS.json's figure is a COPY (read + write); a pure read stream
// can run faster, so this is the ceiling the storkm stream should be compared with as well.
template <int TILE_BYTES, int STAGES>
__global__ void __launch_bounds__(STAGES * 32, 1)
k_stream_read(const double *__restrict__ src, long long ntiles, double *out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *tiles = reinterpret_cast<double *>(smem_raw);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * TILE_BYTES);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t0 = ntiles * blockIdx.x / gridDim.x, t1 = ntiles * (blockIdx.x + 1) / gridDim.x;
  double *tile = tiles + (size_t)w * (TILE_BYTES / 8);
  const uint32_t bar = smem_u32(&bars[w]), tile_s = smem_u32(tile);
  uint64_t policy = 0;
  if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); policy = policy_evict_first(); }
  __syncwarp();
  long long t = t0 + w;
  if (t < t1 && lane == 0) { mbar_expect_tx(bar, TILE_BYTES); bulk_g2s(tile_s, src + t * (TILE_BYTES / 8), TILE_BYTES, bar, policy); }
  uint32_t phase = 0;
  double acc = 0.0;
  for (; t < t1; t += STAGES) {
    mbar_wait(bar, phase);
    phase ^= 1;
    acc += tile[lane];
    __syncwarp();
    if (t + STAGES < t1 && lane == 0) {
      fence_proxy_async();
      mbar_expect_tx(bar, TILE_BYTES);
      bulk_g2s(tile_s, src + (t + STAGES) * (TILE_BYTES / 8), TILE_BYTES, bar, policy);
    }
  }
  if (acc == 12345.678) out[0] = acc;
}

// DFMA micro-benchmark: the FP64 denominator for the matrix-free variant ("of measured").
// 8 independent fma chains per thread, `iters` rounds; 2 flop per fma.
__global__ void k_fp64_peak(double *out, int iters, double a, double b) {
  double v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
  for (int i = 0; i < iters; ++i) {
    v0 = fma(v0, a, b); v1 = fma(v1, a, b); v2 = fma(v2, a, b); v3 = fma(v3, a, b);
    v4 = fma(v4, a, b); v5 = fma(v5, a, b); v6 = fma(v6, a, b); v7 = fma(v7, a, b);
  }
  const double s = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
  if (s == 12345.678) out[0] = s;  // keep the chains alive
}

// The same for the FP64 tensor pipe (mma.sync.m8n8k4.f64 = DMMA.8x8x4, 256 fma per warp instruction): the denominator
// of k_apply_mf3.  9 independent accumulator fragments per warp, as in the kernel's phases.
__global__ void k_fp64_tensor_peak(double *out, int iters, double a, double b) {
  double acc[9][2];
#pragma unroll
  for (int i = 0; i < 9; ++i) acc[i][0] = acc[i][1] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 9; ++i) dmma884(acc[i][0], acc[i][1], a, b);
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) s += acc[i][0] + acc[i][1];
  if (s == 12345.678) out[0] = s;
}

}  // namespace pf
