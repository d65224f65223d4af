// device.cu -- section A of include/parafem_b200.h: the device side of the
// p121 / p123 EBE-PCG path (context, setup, solver loop, halo exchange, C-ABI).
// Kernels live in kernels.cuh.  Built for sm_100a only; there is no CPU path:
// every entry point returns an error when no CUDA device is usable.
#include "kernels.cuh"
#include "parafem_b200.h"

#include <nccl.h>
#include <cub/device/device_radix_sort.cuh>   // setup only: the stable (slot, position) sort behind the scatter tables

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <set>
#include <string>
#include <vector>

using namespace pf;

namespace {

std::string g_last_error;

// ---- NCCL through dlopen: the library stays loadable on a CPU-only box and picks
// up whichever libnccl.so.2 the process already has (torch's bundled one under
// torchrun, the system one for the C++ driver) ----
struct Nccl {
  void *so = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string &err) {
    if (so) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) { so = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (so) break; }
    if (!so) { err = std::string("dlopen libnccl.so.2 failed: ") + dlerror(); return false; }
#define L(name) name = reinterpret_cast<decltype(name)>(dlsym(so, "nccl" #name)); if (!name) { err = "missing nccl" #name; return false; }
    L(GetUniqueId) L(CommInitRank) L(CommDestroy) L(GroupStart) L(GroupEnd) L(Send) L(Recv) L(AllGather) L(GetErrorString)
#undef L
    return true;
  }
} g_nccl;

template <class T>
struct DevBuf {
  T *p = nullptr; size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }          // temporaries are freed on every return path (CU / NC / NEED return early)
  cudaError_t alloc(size_t count) {
    release(); n = count;
    return cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  size_t bytes() const { return n * sizeof(T); }
};

enum { K_MATVEC = 0, K_SCATTER = 1, K_VECTOR = 2, K_HALO = 3, K_NKINDS = 4 };

}  // namespace

struct pf_ctx {
  int rank = 0, nranks = 1, device = 0, sm_count = 0;
  cudaStream_t stream = nullptr;
  ncclComm_t comm = nullptr;
  std::string err;
  int64_t launches = 0;

  // mesh
  int nod = 0, nodof = 0, nip = 0, ntot = 0;
  int64_t nels = 0, neq = 0, ieq_start = 0, neq_pp = 0, nhalo = 0, nslots = 0;
  bool have_mesh = false, have_km = false, have_precon = false, matrix_free = false;
  DevBuf<double> coord, km, utemp, diag_tmp, geom;
  // p129 (forced vibration, theta method): km holds store_mm*c3 + store_km*c4 (the PCG matrix), kb store_km*c2 +
  // store_mm*c3 and kc store_mm/theta (the two right-hand-side products); displacement / velocity / acceleration
  bool dynamic = false;
  double dyn_theta = 0.0, dyn_dtim = 0.0;
  DevBuf<double> kc, fext, x0, d1x0, d2x0;
  // pcg_km (maths.f90:1152-1323): one element matrix shared by every element
  DevBuf<double> km1;
  bool one_km = false;
  // p124 (transient conduction): km holds storka_pp (the PCG matrix), kb holds storkb_pp; mat_override
  // selects the matrix set of the next operator product (nullptr = km)
  DevBuf<double> kb, val_f;
  const double *mat_override = nullptr;
  bool transient = false, transient_first = false;
  bool explicit_ = false;   // p125: km holds store_pm_pp, diag_ext the inverted lumped mass globma_pp
  int mf_mode = 0;
  int km_layout = 0;   // 0: storkm_pp(ntot,ntot,nels_pp) as the reference; 1: packed lower triangles (SymCfg)
  DevBuf<int> ggl;
  DevBuf<unsigned int> csr_ptr, csr_pos;

  // halo tables (sizes per peer rank)
  std::vector<int64_t> get_cnt, get_off, put_cnt, put_off;
  int64_t nput = 0;
  DevBuf<int> put_slot;                 // owned slots wanted by peers, grouped by peer
  DevBuf<double> sendbuf, recvbuf;      // nput each
  int nacc = 0;
  DevBuf<int> acc_slot;
  DevBuf<unsigned int> acc_ptr, acc_pos;

  // fused exchanges of the peer transport: forward puts from k_pupdate (owned equations wanted by peers),
  // reverse accumulation inside k_dot (accumulate entries per reduction chunk)
  DevBuf<unsigned int> put_bits, pk_ptr, acc_chunk_ptr;
  DevBuf<int> pk_slot0, pk_rank;
  DevBuf<long long> pk_dst;
  int npk = 0;
  State *snap_pinned = nullptr;         // two pinned snapshots of the device state (double-buffered polling)
  cudaEvent_t snap_ev[2] = {nullptr, nullptr};

  // peer-memory collectives (CUDA IPC mappings of the peers' p_ext / receive buffer / sync block)
  bool peer_ok = false, use_peer = true, ptab_valid = false;
  DevBuf<PeerSync> sync;
  DevBuf<PeerTable> ptab;
  PeerTable host_tab;
  std::vector<void *> imports;

  // vectors (p_ext/u_ext/diag_ext are slot-indexed: [0] dump, 1..neq_pp owned, then halo)
  DevBuf<double> p_ext, u_ext, diag_ext, r, x, d, part, gath;
  DevBuf<State> state;
  DevBuf<double> ratio_hist;
  int ratio_cap = 0, last_iters = 0;
  double last_ms = 0.0;   // device time of the last pf_pcg_run loop (CUDA events on the solver stream)

  // p123 fixed freedoms; fixed_mode 1 (p122 after the first plastic iteration): u(j) = 0 instead of p(j)*store
  int fixed_mode = 0;
  int nfixed = 0;
  DevBuf<int> fix_slot;
  DevBuf<double> store;

  // p122 (elasto-plasticity): viscoplastic strains and stresses of every Gauss point, load-increment vectors
  bool plastic = false;
  PlasticParams pl_par;
  DevBuf<double> evpt, tensor, bdylds, oldis, totd, loads, ld0, valf;

  // p1210 (explicit elasto-plastic dynamics): strains / stresses of every Gauss point, lumped mass, external loads,
  // velocity and acceleration; the displacement lives in the owned part of p_ext
  bool vm_explicit = false;
  int vm_form = 0;          // 0: elements_2 as written (k_p1210_elements); 1: operator form on the tensor cores (k_p1210_mf)
  VmParams vm_par;
  double vm_dtim = 0.0, vm_pload = 0.0;
  DevBuf<double> vm_eten, vm_ten, vm_mm, vm_fext, vm_d1, vm_d2;

  // one PCG iteration captured as a CUDA graph (single rank; re-captured when the problem changes)
  cudaGraphExec_t graph_exec = nullptr;
  int64_t epoch = 0, graph_epoch = -1, graph_launches = 0;

  // per-handle (= per-device) launch state: kernels whose dynamic shared-memory opt-in was set on this device,
  // resident blocks per SM of the reduction kernels, and the element tables this handle last uploaded to c_tab
  std::set<const void *> smem_ok;
  std::map<const void *, int> occ;
  ElemTables tab;
  bool have_tab = false;

  // profiling
  bool profile = false;
  struct Span { cudaEvent_t a, b; int kind; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> pool;
  double kind_ms[K_NKINDS] = {0, 0, 0, 0};
  int64_t kind_n[K_NKINDS] = {0, 0, 0, 0};
};

namespace {

int fail(pf_handle h, int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  g_last_error = buf;
  if (h) h->err = buf;
  return code;
}

#define CU(call)                                                                           \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) return fail(h, 10, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define NC(call)                                                                           \
  do {                                                                                     \
    ncclResult_t r_ = (call);                                                              \
    if (r_ != ncclSuccess) return fail(h, 11, "%s: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
  } while (0)
#define NEED(cond, msg) do { if (!(cond)) return fail(h, 2, "%s: %s", __func__, msg); } while (0)

// c_tab is per DEVICE (one __constant__ copy each) while tables are per HANDLE: the handle that uploaded last is
// remembered per device, and a handle whose kernels read c_tab at apply time (k_apply_mf: dee, weights) re-uploads
// its own tables when another handle on the same device has formed matrices in between.
pf_ctx *g_ctab_owner[64] = {nullptr};
int upload_tables(pf_handle h, const ElemTables &T) {
  CU(cudaMemcpyToSymbol(c_tab, &T, sizeof T));
  h->tab = T; h->have_tab = true;
  if (h->device >= 0 && h->device < 64) g_ctab_owner[h->device] = h;
  return 0;
}
int ensure_tables(pf_handle h) {
  if (!h->have_tab || (h->device >= 0 && h->device < 64 && g_ctab_owner[h->device] == h)) return 0;
  CU(cudaStreamSynchronize(h->stream));
  return upload_tables(h, h->tab);
}
// dynamic shared memory above 48 KB needs an opt-in per kernel and per device
template <class K>
int ensure_smem(pf_handle h, K kern, size_t bytes) {
  const void *key = reinterpret_cast<const void *>(kern);
  if (h->smem_ok.count(key)) return 0;
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  h->smem_ok.insert(key);
  return 0;
}
// resident blocks per SM of a kernel at `threads` threads (cached per handle)
template <class K>
int resident_blocks(pf_handle h, K kern, int threads, int fallback) {
  const void *key = reinterpret_cast<const void *>(kern);
  auto it = h->occ.find(key);
  if (it != h->occ.end()) return it->second;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0) != cudaSuccess || per_sm < 1) per_sm = fallback;
  h->occ[key] = per_sm;
  return per_sm;
}

// ---- profiling spans: CUDA events on the solver stream around a group of launches ----
struct Scope {
  pf_handle h; int kind; cudaEvent_t a = nullptr, b = nullptr;
  Scope(pf_handle h_, int kind_) : h(h_), kind(kind_) {
    if (!h->profile) return;
    auto get = [&]() { cudaEvent_t e; if (h->pool.empty()) cudaEventCreate(&e); else { e = h->pool.back(); h->pool.pop_back(); } return e; };
    a = get(); b = get();
    cudaEventRecord(a, h->stream);
  }
  ~Scope() {
    if (!h->profile) return;
    cudaEventRecord(b, h->stream);
    h->spans.push_back({a, b, kind});
  }
};

void collect_spans(pf_handle h) {
  for (auto &s : h->spans) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) { h->kind_ms[s.kind] += ms; h->kind_n[s.kind] += 1; }
    h->pool.push_back(s.a); h->pool.push_back(s.b);
  }
  h->spans.clear();
}

// a pair of timing events released on every return path
struct EventPair {
  cudaEvent_t a = nullptr, b = nullptr;
  cudaError_t create() { cudaError_t e = cudaEventCreate(&a); return e != cudaSuccess ? e : cudaEventCreate(&b); }
  ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
};

int grid_for(pf_handle h, int64_t n, int threads, int per_sm = 8) {
  int64_t blocks = (n + threads - 1) / threads;
  int64_t cap = (int64_t)h->sm_count * per_sm;
  return (int)std::max<int64_t>(1, std::min(blocks, cap));
}

// shape_der at one local point (new_library.f90:745-794, :865-896); D[a*20+m] = der(a,m); fun (8-node brick only,
// shape_fun :397-422) may be null
int shape_der_at(int nod, double xi, double eta, double zeta, double *D, double *fun8) {
  for (int q = 0; q < 60; ++q) D[q] = 0.0;
  if (nod == 4) {        // shape_der, 3-D nod = 4 (new_library.f90:757-767): constant
    D[0] = 1.0; D[20 + 1] = 1.0; D[40 + 2] = 1.0;
    D[3] = -1.0; D[20 + 3] = -1.0; D[40 + 3] = -1.0;
  } else if (nod == 8) {
    const double em = 1.0 - eta, xm = 1.0 - xi, zm = 1.0 - zeta, ep = eta + 1.0, xp = xi + 1.0, zp = zeta + 1.0;
    const double dx[8] = {-0.125 * em * zm, -0.125 * em * zp, 0.125 * em * zp, 0.125 * em * zm,
                          -0.125 * ep * zm, -0.125 * ep * zp, 0.125 * ep * zp, 0.125 * ep * zm};
    const double dy[8] = {-0.125 * xm * zm, -0.125 * xm * zp, -0.125 * xp * zp, -0.125 * xp * zm,
                          0.125 * xm * zm, 0.125 * xm * zp, 0.125 * xp * zp, 0.125 * xp * zm};
    const double dz[8] = {-0.125 * xm * em, 0.125 * xm * em, 0.125 * xp * em, -0.125 * xp * em,
                          -0.125 * xm * ep, 0.125 * xm * ep, 0.125 * xp * ep, -0.125 * xp * ep};
    for (int m = 0; m < 8; ++m) { D[m] = dx[m]; D[20 + m] = dy[m]; D[40 + m] = dz[m]; }
    if (fun8) {
      const double fn[8] = {0.125 * xm * em * zm, 0.125 * xm * em * zp, 0.125 * xp * em * zp, 0.125 * xp * em * zm,
                            0.125 * xm * ep * zm, 0.125 * xm * ep * zp, 0.125 * xp * ep * zp, 0.125 * xp * ep * zm};
      for (int m = 0; m < 8; ++m) fun8[m] = fn[m];
    }
  } else if (nod == 20) {
    // corner / mid-edge classes of the 20-node brick in S&G order
    const int sx[20] = {-1, -1, -1, 0, 1, 1, 1, 0, -1, -1, 1, 1, -1, -1, -1, 0, 1, 1, 1, 0};
    const int sy[20] = {-1, -1, -1, -1, -1, -1, -1, -1, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1};
    const int sz[20] = {-1, 0, 1, 1, 1, 0, -1, -1, -1, 1, 1, -1, -1, 0, 1, 1, 1, 0, -1, -1};
    for (int m = 0; m < 20; ++m) {
      const double a = sx[m], b = sy[m], c = sz[m];
      const double x0 = xi * a, e0 = eta * b, z0 = zeta * c;
      double gx, gy, gz;
      if (sx[m] == 0) {         // mid-edge along xi
        gx = -.5 * xi * (1. + e0) * (1. + z0);
        gy = .25 * b * (1. - xi * xi) * (1. + z0);
        gz = .25 * c * (1. - xi * xi) * (1. + e0);
      } else if (sy[m] == 0) {  // mid-edge along eta
        gx = .25 * a * (1. - eta * eta) * (1. + z0);
        gy = -.5 * eta * (1. + x0) * (1. + z0);
        gz = .25 * c * (1. + x0) * (1. - eta * eta);
      } else if (sz[m] == 0) {  // mid-edge along zeta
        gx = .25 * a * (1. + e0) * (1. - zeta * zeta);
        gy = .25 * b * (1. + x0) * (1. - zeta * zeta);
        gz = -.5 * zeta * (1. + x0) * (1. + e0);
      } else {                  // corner
        gx = .125 * a * (1. + e0) * (1. + z0) * (2. * x0 + e0 + z0 - 1.);
        gy = .125 * b * (1. + x0) * (1. + z0) * (x0 + 2. * e0 + z0 - 1.);
        gz = .125 * c * (1. + x0) * (1. + e0) * (x0 + e0 + 2. * z0 - 1.);
      }
      D[m] = gx; D[20 + m] = gy; D[40 + m] = gz;
    }
  } else return 2;
  return 0;
}

// ---- host restatement of the element tables (product code; the oracle has its own) ----
// sample('hexahedron') new_library.f90:1397-1433; shape_der :745-794, :865-896; deemat :1671-1686
int fill_tables(int nod, int nip, double e, double v, double kx, double ky, double kz, ElemTables &T) {
  memset(&T, 0, sizeof T);
  double pts[27][3];
  if (nod == 4) {          // sample('tetrahedron') (new_library.f90:1328-1378): nip = 1 centroid, weight 1/6;
    // nip = 4 / 5 are written with default-real (single-precision) literals there: restated with floats
    for (int i = 0; i < 8; ++i) pts[i][0] = pts[i][1] = pts[i][2] = 0.0;
    if (nip == 1) { pts[0][0] = pts[0][1] = pts[0][2] = 0.25; T.weights[0] = 1.0 / 6.0; }
    else if (nip == 4) {
      const double a = (double).58541020f, b = (double).13819660f;
      for (int i = 0; i < 4; ++i) { pts[i][0] = pts[i][1] = pts[i][2] = b; T.weights[i] = (double)(.25f / 6.f); }
      pts[0][0] = a; pts[1][1] = a; pts[2][2] = a;
    } else if (nip == 5) {
      const double q = (double).25f, hf = (double).5f, sx = (double)(1.f / 6.f);
      pts[0][0] = pts[0][1] = pts[0][2] = q;
      for (int i = 1; i < 5; ++i) pts[i][0] = pts[i][1] = pts[i][2] = sx;
      pts[1][0] = hf; pts[2][1] = hf; pts[3][2] = hf;
      T.weights[0] = (double)(-.8f);
      for (int i = 1; i < 5; ++i) T.weights[i] = (double)(9.f / 20.f);
      for (int i = 0; i < 5; ++i) T.weights[i] = T.weights[i] / (double)6.f;
    } else return 1;
  } else if (nip == 1) { pts[0][0] = pts[0][1] = pts[0][2] = 0.0; T.weights[0] = 8.0; }
  else if (nip == 8) {
    const double r3 = 1.0 / std::sqrt(3.0);
    for (int i = 0; i < 8; ++i) {
      pts[i][0] = (i < 4) ? r3 : -r3;
      pts[i][1] = (i == 0 || i == 1 || i == 4 || i == 6) ? r3 : -r3;
      pts[i][2] = (i == 0 || i == 2 || i == 4 || i == 5) ? r3 : -r3;
      T.weights[i] = 1.0;
    }
  } else if (nip == 27 && nod != 4) {
    // sample('hexahedron'), nip = 27 (new_library.f90:1491-1517): wt = (/5./9.*v,8./9.*v,5./9.*v/) with DEFAULT-REAL
    // outer factors (single-precision quotients, widened) and v(9) = w (x) w, w = (5/9,8/9,5/9) in REAL(iwp) (:1053-1054)
    const double r15 = 0.2 * std::sqrt(15.0);
    const double w[3] = {5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0};
    const double f59 = (double)(5.f / 9.f), f89 = (double)(8.f / 9.f);
    for (int blk = 0; blk < 3; ++blk)
      for (int q = 0; q < 9; ++q) {
        const int i = 9 * blk + q;
        pts[i][0] = (q % 3 == 0) ? -r15 : (q % 3 == 1) ? 0.0 : r15;
        pts[i][2] = (q / 3 == 0) ? r15 : (q / 3 == 1) ? 0.0 : -r15;
        pts[i][1] = blk == 0 ? -r15 : blk == 1 ? 0.0 : r15;
        T.weights[i] = (blk == 1 ? f89 : f59) * (w[q / 3] * w[q % 3]);
      }
  } else return 1;
  T.nip = nip;
  for (int ig = 0; ig < nip; ++ig)
    if (shape_der_at(nod, pts[ig][0], pts[ig][1], pts[ig][2], T.der + ig * 60, (nod == 8 && ig < 8) ? T.fun + ig * 8 : nullptr)) return 2;
  if (nod == 20)           // shape_fun, 3-D nod = 20 (new_library.f90:449-468)
    for (int ig = 0; ig < nip; ++ig) {
      static const int xii[20] = {-1, -1, -1, 0, 1, 1, 1, 0, -1, -1, 1, 1, -1, -1, -1, 0, 1, 1, 1, 0};
      static const int etai[20] = {-1, -1, -1, -1, -1, -1, -1, -1, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1};
      static const int zetai[20] = {-1, 0, 1, 1, 1, 0, -1, -1, -1, 1, 1, -1, -1, 0, 1, 1, 1, 0, -1, -1};
      const double xi = pts[ig][0], eta = pts[ig][1], zeta = pts[ig][2];
      for (int l = 1; l <= 20; ++l) {
        const double xi0 = xi * xii[l - 1], eta0 = eta * etai[l - 1], zeta0 = zeta * zetai[l - 1];
        double f;
        if (l == 4 || l == 8 || l == 16 || l == 20) f = .25 * (1. - xi * xi) * (1. + eta0) * (1. + zeta0);
        else if (l >= 9 && l <= 12) f = .25 * (1. + xi0) * (1. - eta * eta) * (1. + zeta0);
        else if (l == 2 || l == 6 || l == 14 || l == 18) f = .25 * (1. + xi0) * (1. + eta0) * (1. - zeta * zeta);
        else f = .125 * (1. + xi0) * (1. + eta0) * (1. + zeta0) * (xi0 + eta0 + zeta0 - 2);
        T.fun20[ig * 20 + l - 1] = f;
      }
    }
  // deemat, 6x6
  const double v2 = v / (1.0 - v), vv = (1.0 - 2.0 * v) / (1.0 - v) * 0.5;
  for (int i = 0; i < 3; ++i) T.dee[i * 6 + i] = 1.0;
  for (int i = 3; i < 6; ++i) T.dee[i * 6 + i] = vv;
  T.dee[1 * 6 + 0] = T.dee[0 * 6 + 1] = T.dee[2 * 6 + 0] = T.dee[0 * 6 + 2] = T.dee[2 * 6 + 1] = T.dee[1 * 6 + 2] = v2;
  if (e != 0.0) { const double den = 2.0 * (1.0 + v) * vv; for (int i = 0; i < 36; ++i) T.dee[i] = T.dee[i] * e / den; }
  T.kxyz[0] = kx; T.kxyz[1] = ky; T.kxyz[2] = kz;
  return 0;
}

// ---- mat-vec dispatch ----
template <int NTOT, int EPT, int STAGES, bool GATHER>
int launch_matvec_t(pf_handle h, const double *pvec, const State *st, PeerTable *T) {
  using Cfg = MatvecCfg<NTOT, EPT, STAGES>;
  auto kern = k_matvec<NTOT, EPT, STAGES, GATHER>;
  if (int rc_ = ensure_smem(h, kern, Cfg::kSmem)) return rc_;
  const int64_t ntiles = (h->nels + EPT - 1) / EPT;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(h->sm_count, ntiles));
  kern<<<grid, Cfg::kThreads, Cfg::kSmem, h->stream>>>(h->mat_override ? h->mat_override : h->km.p, h->ggl.p, pvec, h->utemp.p,
                                                       (long long)h->nels, st, T);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

template <int NTOT, int EPT, int WARPS, int NBUF, bool GATHER>
int launch_matvec2_t(pf_handle h, const double *pvec, const State *st, PeerTable *T) {
  using Cfg = Matvec2Cfg<NTOT, EPT, WARPS, NBUF>;
  auto kern = k_matvec2<NTOT, EPT, WARPS, NBUF, GATHER>;
  if (int rc_ = ensure_smem(h, kern, Cfg::kSmem)) return rc_;
  const int64_t ntiles = (h->nels + EPT - 1) / EPT;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(h->sm_count, ntiles));
  kern<<<grid, Cfg::kThreads, Cfg::kSmem, h->stream>>>(h->mat_override ? h->mat_override : h->km.p, h->ggl.p, pvec, h->utemp.p,
                                                       (long long)h->nels, st, T);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

constexpr int kMfWarps = 8;   // 256 threads = 256 elements in flight per SM; ~190 registers per thread

int mf_grid(pf_handle h) {
  const int64_t ngroups = (h->nels + 31) / 32;
  return (int)std::max<int64_t>(1, std::min<int64_t>(h->sm_count, (ngroups + kMfWarps - 1) / kMfWarps));
}

template <int NTOT, int EPT, int STAGES, bool GATHER>
int launch_matvec_sym_t(pf_handle h, const double *pvec, const State *st, PeerTable *T) {
  using Cfg = MatvecSymCfg<NTOT, EPT, STAGES>;
  auto kern = k_matvec_sym<NTOT, EPT, STAGES, GATHER>;
  if (int rc_ = ensure_smem(h, kern, Cfg::kSmem)) return rc_;
  const int64_t ntiles = (h->nels + EPT - 1) / EPT;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(h->sm_count, ntiles));
  kern<<<grid, Cfg::kThreads, Cfg::kSmem, h->stream>>>(h->km.p, h->ggl.p, pvec, h->utemp.p, (long long)h->nels, st, T);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

template <int NOD, bool GATHER, int GEOM, int UNR = 2>
int launch_mf_t(pf_handle h, const double *pvec, const State *st, PeerTable *T = nullptr) {
  using Cfg = MfCfg<NOD>;
  auto kern = k_apply_mf<NOD, GATHER, GEOM, kMfWarps, UNR>;
  if (int rc_ = ensure_smem(h, kern, Cfg::smem(kMfWarps))) return rc_;
  kern<<<mf_grid(h), kMfWarps * 32, Cfg::smem(kMfWarps), h->stream>>>(h->coord.p, h->ggl.p, pvec, h->utemp.p, (long long)h->nels, st,
                                                                      h->geom.p, T);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

template <int NTOT, bool GATHER>
int launch_matvec_km_t(pf_handle h, const double *pvec, const State *st, PeerTable *T) {
  const int64_t npass = (h->nels + (64 / NTOT) - 1) / (64 / NTOT);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(npass, (int64_t)h->sm_count * 8));
  k_matvec_km<NTOT, GATHER><<<grid, 64, 0, h->stream>>>(h->km1.p, h->ggl.p, pvec, h->utemp.p, (long long)h->nels, st, T);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

// k_apply_mf2: two lanes per element; 12 warps per SM (3 per scheduler, 168 registers, no spills) with the two halves
// of an element on the two half-warps measured best (profiles/r02_mf_kernel_round2.md); PF_MF=1lane selects k_apply_mf
constexpr int kMf2Warps = 12;
bool mf_two_lanes();
int mf2_grid(pf_handle h) {
  const int64_t nhg = (h->nels + 15) / 16;
  return (int)std::max<int64_t>(1, std::min<int64_t>(h->sm_count, (nhg + kMf2Warps - 1) / kMf2Warps));
}
template <int NOD, bool GATHER, int GEOM, int WARPS, int PAIR>
int launch_mf2_tt(pf_handle h, const double *pvec, const State *st, PeerTable *T) {
  using Cfg = Mf2Cfg<NOD>;
  auto kern = k_apply_mf2<NOD, GATHER, GEOM, WARPS, 2, PAIR>;
  if (int rc_ = ensure_smem(h, kern, Cfg::smem(WARPS))) return rc_;
  const int64_t nhg = (h->nels + 15) / 16;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(h->sm_count, (nhg + WARPS - 1) / WARPS));
  kern<<<grid, WARPS * 32, Cfg::smem(WARPS), h->stream>>>(h->coord.p, h->ggl.p, pvec, h->utemp.p, (long long)h->nels, st, h->geom.p, T);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}
// PF_MF2=1 (adjacent-lane pairs) / PF_MF2W=8|16 select the other measured variants for A/B runs
template <int NOD, bool GATHER, int GEOM>
int launch_mf2_t(pf_handle h, const double *pvec, const State *st, PeerTable *T = nullptr) {
  static const int pair = getenv("PF_MF2") ? atoi(getenv("PF_MF2")) : 16;
  static const int warps = getenv("PF_MF2W") ? atoi(getenv("PF_MF2W")) : kMf2Warps;
  if (GEOM == 2 && GATHER && NOD == 20) {
    if (pair == 1 && warps == 16) return launch_mf2_tt<NOD, GATHER, GEOM, 16, 1>(h, pvec, st, T);
    if (pair == 1) return launch_mf2_tt<NOD, GATHER, GEOM, kMf2Warps, 1>(h, pvec, st, T);
    if (warps == 16) return launch_mf2_tt<NOD, GATHER, GEOM, 16, 16>(h, pvec, st, T);
    if (warps == 8) return launch_mf2_tt<NOD, GATHER, GEOM, 8, 16>(h, pvec, st, T);
  }
  return launch_mf2_tt<NOD, GATHER, GEOM, kMf2Warps, 16>(h, pvec, st, T);
}

// k_apply_mf4 / k_apply_mf3: the FP64 tensor-core kernels (mma.sync.m8n8k4.f64); the warp-specialised k_apply_mf4 is the
// default for both bricks and both matrix-free modes, PF_MF=3 selects the first build (every warp moves its own data).
// PF_MF=2lane / 1lane select the round-2 / round-1 DFMA kernels for A/B runs (each has its own summation order, mirrored
// by the oracle: orc_set_mf_order).
int mf_kernel_choice() {
  static const int c = [] {
    const char *e = getenv("PF_MF");
    if (e && !strcmp(e, "1lane")) return 1;
    if (e && !strcmp(e, "2lane")) return 2;
    if (e && !strcmp(e, "3")) return 3;
    return 4;
  }();
  return c;
}
bool mf_two_lanes() { return mf_kernel_choice() != 1; }
constexpr int kMf3Warps = 12;
template <int NOD, bool GATHER, int GEOM, int WARPS, int BREG>
int launch_mf3_tt(pf_handle h, const double *pvec, const State *st, PeerTable *T) {
  using Cfg = Mf3Cfg<NOD>;
  auto kern = k_apply_mf3<NOD, GATHER, GEOM, WARPS, BREG>;
  if (int rc_ = ensure_smem(h, kern, Cfg::smem(WARPS, GEOM))) return rc_;
  const int64_t npass = (h->nels + Cfg::EPP - 1) / Cfg::EPP;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(h->sm_count, (npass + WARPS - 1) / WARPS));
  kern<<<grid, WARPS * 32, Cfg::smem(WARPS, GEOM), h->stream>>>(h->coord.p, h->ggl.p, pvec, h->utemp.p, (long long)h->nels, st, h->geom.p, T);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}
// PF_MF3W = 16 | 12h (12 warps, phase-1 fragments in registers) | default 12 warps, fragments from the shared-memory table
template <int NOD, bool GATHER, int GEOM>
int launch_mf3_t(pf_handle h, const double *pvec, const State *st, PeerTable *T = nullptr) {
  static const char *sel = getenv("PF_MF3W") ? getenv("PF_MF3W") : "";
  if (!strcmp(sel, "12h")) return launch_mf3_tt<NOD, GATHER, GEOM, 12, 1>(h, pvec, st, T);
  if (!strcmp(sel, "16")) return launch_mf3_tt<NOD, GATHER, GEOM, 16, 0>(h, pvec, st, T);
  return launch_mf3_tt<NOD, GATHER, GEOM, kMf3Warps, 0>(h, pvec, st, T);
}

// k_apply_mf4: the warp-specialised build of the tensor-core kernel (PF_MF=3 selects k_apply_mf3 for A/B runs);
// PF_MF4C = consumer warps per SM (8 or 12)
constexpr int kMf4Cons = 12;
template <int NOD, bool GATHER, int GEOM, int NCW>
int launch_mf4_tt(pf_handle h, const double *pvec, const State *st, PeerTable *T) {
  using Cfg = Mf4Cfg<NOD, NCW>;
  auto kern = k_apply_mf4<NOD, GATHER, GEOM, NCW>;
  if (int rc_ = ensure_smem(h, kern, Cfg::kSmem)) return rc_;
  const int64_t npass = (h->nels + Cfg::EPP - 1) / Cfg::EPP;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(h->sm_count, (npass + Cfg::NCONS - 1) / Cfg::NCONS));
  kern<<<grid, Cfg::kThreads, Cfg::kSmem, h->stream>>>(h->coord.p, h->ggl.p, pvec, h->utemp.p, (long long)h->nels, st, h->geom.p, T,
                                                       nullptr, nullptr, VmParams{0.0, 0.0, 0.0});
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}
template <int NOD, bool GATHER, int GEOM>
int launch_mf4_t(pf_handle h, const double *pvec, const State *st, PeerTable *T = nullptr) {
  static const int ncw = getenv("PF_MF4C") ? atoi(getenv("PF_MF4C")) : kMf4Cons;
  if (ncw == 8) return launch_mf4_tt<NOD, GATHER, GEOM, 8>(h, pvec, st, T);
  return launch_mf4_tt<NOD, GATHER, GEOM, kMf4Cons>(h, pvec, st, T);
}

template <bool GATHER>
int launch_matvec(pf_handle h, const double *pvec, const State *st, PeerTable *T = nullptr) {
  Scope sc(h, K_MATVEC);
  if (h->one_km) {
    switch (h->ntot) {
      case 60: return launch_matvec_km_t<60, GATHER>(h, pvec, st, T);
      case 24: return launch_matvec_km_t<24, GATHER>(h, pvec, st, T);
      case 12: return launch_matvec_km_t<12, GATHER>(h, pvec, st, T);
      case 8: return launch_matvec_km_t<8, GATHER>(h, pvec, st, T);
      case 4: return launch_matvec_km_t<4, GATHER>(h, pvec, st, T);
    }
    return fail(h, 3, "unsupported ntot %d", h->ntot);
  }
  // 20-node bricks: two lanes per element (k_apply_mf2; config C: mode 2 0.754 against 0.867 ms, mode 1 1.30 against 1.37 ms);
  // 8-node bricks keep the one-lane kernel (200^3: 1.79 against 2.17 ms -- too little arithmetic per element to pay for
  // the pairing).  The two kernels add the Gauss points' contributions in different orders; the oracle mirrors each
  // (orc_apply_mf: order by element type).
  if (h->matrix_free && mf_kernel_choice() == 4) {
    if (h->nod == 20) return h->mf_mode == 2 ? launch_mf4_t<20, GATHER, 2>(h, pvec, st, T) : launch_mf4_t<20, GATHER, 0>(h, pvec, st, T);
    return h->mf_mode == 2 ? launch_mf4_t<8, GATHER, 2>(h, pvec, st, T) : launch_mf4_t<8, GATHER, 0>(h, pvec, st, T);
  }
  if (h->matrix_free && mf_kernel_choice() == 3) {
    if (h->nod == 20) return h->mf_mode == 2 ? launch_mf3_t<20, GATHER, 2>(h, pvec, st, T) : launch_mf3_t<20, GATHER, 0>(h, pvec, st, T);
    return h->mf_mode == 2 ? launch_mf3_t<8, GATHER, 2>(h, pvec, st, T) : launch_mf3_t<8, GATHER, 0>(h, pvec, st, T);
  }
  if (h->matrix_free && h->nod == 20 && mf_two_lanes())
    return h->mf_mode == 2 ? launch_mf2_t<20, GATHER, 2>(h, pvec, st, T) : launch_mf2_t<20, GATHER, 0>(h, pvec, st, T);
  if (h->matrix_free) {
    if (h->mf_mode == 2) {
      // PF_TUNE: unroll factor of the node-pair loops (default 2; measured in profiles/r01_mf_kernel_history.md)
      static const int mtune = getenv("PF_TUNE") ? atoi(getenv("PF_TUNE")) : 0;
      if (h->nod == 20 && GATHER && mtune == 1) return launch_mf_t<20, GATHER, 2, 1>(h, pvec, st, T);
      if (h->nod == 20 && GATHER && mtune == 2) return launch_mf_t<20, GATHER, 2, 5>(h, pvec, st, T);
      if (h->nod == 20) return launch_mf_t<20, GATHER, 2>(h, pvec, st, T);
      return launch_mf_t<8, GATHER, 2>(h, pvec, st, T);
    }
    if (h->nod == 20) return launch_mf_t<20, GATHER, 0>(h, pvec, st, T);
    return launch_mf_t<8, GATHER, 0>(h, pvec, st, T);
  }
  // PF_TUNE selects an alternative tile shape (elements per tile x ring slots) for experiments
  static const int tune = getenv("PF_TUNE") ? atoi(getenv("PF_TUNE")) : 0;
  if (h->km_layout == 1) {
    switch (h->ntot) {
      case 60:
        // measured: 10 ring slots 0.95 of HBM peak, 13 slots 0.93, 8 slots 0.94, 2 elements x 6 slots 0.72
        if (tune == 1) return launch_matvec_sym_t<60, 1, 13, GATHER>(h, pvec, st, T);
        if (tune == 2) return launch_matvec_sym_t<60, 1, 8, GATHER>(h, pvec, st, T);
        return launch_matvec_sym_t<60, 1, 10, GATHER>(h, pvec, st, T);
      case 24:   // measured (profiles/r01_symmetric_layout.md): 4 x 16 beats 2 x 32 (64-register cap, spills)
        if (tune == 1) return launch_matvec_sym_t<24, 2, 32, GATHER>(h, pvec, st, T);
        if (tune == 2) return launch_matvec_sym_t<24, 6, 12, GATHER>(h, pvec, st, T);
        return launch_matvec_sym_t<24, 4, 16, GATHER>(h, pvec, st, T);
      case 8: return launch_matvec_sym_t<8, 16, 32, GATHER>(h, pvec, st, T);
    }
    return fail(h, 3, "unsupported ntot %d (supported: 60, 24, 8)", h->ntot);
  }
  switch (h->ntot) {
    case 60: return launch_matvec_t<60, 1, 7, GATHER>(h, pvec, st, T);
    // measured on B200 (profiles/r01_tile_tuning.md): small tiles on many ring slots win --
    // hex8 8x5 -> 0.90 of HBM peak, 2x16 -> 0.99, 1x32 -> 1.00; p123 64x6 -> 0.67, 16x16 -> 0.85
    case 24:
      if (tune == 1) return launch_matvec_t<24, 8, 5, GATHER>(h, pvec, st, T);
      if (tune == 2) return launch_matvec_t<24, 2, 16, GATHER>(h, pvec, st, T);
      return launch_matvec_t<24, 1, 32, GATHER>(h, pvec, st, T);
    case 8:
      // 8x8 matrices (p123 / p124 / p125).  Side traffic (indices 32 B, gathered right-hand sides, utemp 64 B per 512 B
      // matrix) is a fifth of the stream, so ~0.83 of the HBM peak on the matrix bytes is the ceiling.  Measured at config
      // B / at 200^3 (scripts/gpu_r2_15.sh, back to back): one slot per warp 16 elements x 16 warps 0.755 / 0.814;
      // k_matvec2 two slots per warp: 12 x 16 0.775 / 0.834, 8 x 24 0.784 / 0.845 (default), 16 x 12 0.783; three slots:
      // 8 x 16 0.718, 4 x 32 0.715
      if (tune == 1) return launch_matvec_t<8, 16, 16, GATHER>(h, pvec, st, T);
      if (tune == 2) return launch_matvec2_t<8, 12, 16, 2, GATHER>(h, pvec, st, T);
      if (tune == 5) return launch_matvec2_t<8, 16, 12, 2, GATHER>(h, pvec, st, T);
      return launch_matvec2_t<8, 8, 24, 2, GATHER>(h, pvec, st, T);
    case 12: return launch_matvec_t<12, 8, 16, GATHER>(h, pvec, st, T);    // 4-node tetrahedra, elastic: 9 KB tiles
    case 4: return launch_matvec_t<4, 64, 16, GATHER>(h, pvec, st, T);     // 4-node tetrahedra, scalar: 8 KB tiles
  }
  return fail(h, 3, "unsupported ntot %d (supported: 60, 24, 12, 8, 4)", h->ntot);
}

int launch_scatter(pf_handle h, const State *st, bool diag, double *dst, PeerTable *T = nullptr) {
  Scope sc(h, K_SCATTER);
  // grid-stride kernel: exactly one resident wave (k_scatter needs 30 registers: 8 blocks of 256 per SM)
  const int grid = grid_for(h, h->nslots, 256, resident_blocks(h, k_scatter<false>, 256, 8));
  if (diag) k_scatter<true><<<grid, 256, 0, h->stream>>>(h->csr_ptr.p, h->csr_pos.p, h->km.p, dst, (long long)h->nslots, h->ntot, st, h->km_layout);
  else k_scatter<false><<<grid, 256, 0, h->stream>>>(h->csr_ptr.p, h->csr_pos.p, h->utemp.p, dst, (long long)h->nslots, h->ntot, st, 0, T,
                                                     (long long)h->neq_pp);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}


void close_imports(pf_handle h) {
  for (void *p : h->imports) if (p) cudaIpcCloseMemHandle(p);
  h->imports.clear();
  h->peer_ok = false;
}

// map the peers' buffers and fill the PeerTable; `all` = counts matrix, all[r*R+o] = number of
// equations rank r gathers from owner o.  Every rank must reach the same verdict.
int setup_peer(pf_handle h, const std::vector<int64_t> &all) {
  const int R = h->nranks, me = h->rank;
  h->peer_ok = false;
  if (!h->use_peer || R > kMaxRanks) return 0;
  struct Handles { cudaIpcMemHandle_t sync, pext, recv; };
  Handles mine;
  int ok = 1;
  if (cudaIpcGetMemHandle(&mine.sync, h->sync.p) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&mine.pext, h->p_ext.p) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&mine.recv, h->recvbuf.p) != cudaSuccess) ok = 0;
  cudaGetLastError();
  DevBuf<char> d_mine, d_all;
  CU(d_mine.alloc(sizeof(Handles))); CU(d_all.alloc(sizeof(Handles) * (size_t)R));
  CU(cudaMemcpy(d_mine.p, &mine, sizeof mine, cudaMemcpyHostToDevice));
  NC(g_nccl.AllGather(d_mine.p, d_all.p, sizeof(Handles), ncclChar, h->comm, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  std::vector<Handles> hs((size_t)R);
  CU(cudaMemcpy(hs.data(), d_all.p, sizeof(Handles) * (size_t)R, cudaMemcpyDeviceToHost));
  d_mine.release(); d_all.release();

  PeerTable T;
  if (h->ptab_valid) CU(cudaMemcpy(&T, h->ptab.p, sizeof T, cudaMemcpyDeviceToHost));  // keep the sequence counters
  else memset(&T, 0, sizeof T);
  T.rank = me; T.nranks = R; T.ticket[0] = T.ticket[1] = 0;
  for (int r = 0; r < kMaxRanks; ++r) { T.sync[r] = nullptr; T.p_ext[r] = nullptr; T.recv[r] = nullptr; T.fwd_dst_off[r] = T.rev_dst_off[r] = 0; }
  for (int r = 0; r < R && ok; ++r) {
    if (r == me) { T.sync[r] = h->sync.p; T.p_ext[r] = h->p_ext.p; T.recv[r] = h->recvbuf.p; continue; }
    void *ps = nullptr, *pp = nullptr, *pr = nullptr;
    if (cudaIpcOpenMemHandle(&ps, hs[(size_t)r].sync, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
    h->imports.push_back(ps);
    if (cudaIpcOpenMemHandle(&pp, hs[(size_t)r].pext, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
    h->imports.push_back(pp);
    if (cudaIpcOpenMemHandle(&pr, hs[(size_t)r].recv, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
    h->imports.push_back(pr);
    T.sync[r] = (PeerSync *)ps; T.p_ext[r] = (double *)pp; T.recv[r] = (double *)pr;
    int64_t neq_pp_r, st_r;
    pf_calc_neq_pp(h->neq, R, r + 1, &neq_pp_r, &st_r);
    int64_t goff = 0; for (int o = 0; o < me; ++o) goff += all[(size_t)r * R + o];     // peer r's halo offset of owner `me`
    int64_t poff = 0; for (int q = 0; q < me; ++q) poff += all[(size_t)q * R + r];     // owner r's receive offset of source `me`
    T.fwd_dst_off[r] = 1 + neq_pp_r + goff;
    T.rev_dst_off[r] = poff;
  }
  cudaGetLastError();
  for (int r = 0; r < R; ++r) { T.put_off[r] = h->put_off[r]; T.get_off[r] = h->get_off[r]; }
  T.put_off[R] = h->nput; T.get_off[R] = h->nhalo;
  for (int r = R + 1; r <= kMaxRanks; ++r) { T.put_off[r] = h->nput; T.get_off[r] = h->nhalo; }
  // unanimous verdict
  DevBuf<int> d_ok, d_oks;
  CU(d_ok.alloc(1)); CU(d_oks.alloc((size_t)R));
  CU(cudaMemcpy(d_ok.p, &ok, sizeof ok, cudaMemcpyHostToDevice));
  NC(g_nccl.AllGather(d_ok.p, d_oks.p, 1, ncclInt32, h->comm, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  std::vector<int> oks((size_t)R);
  CU(cudaMemcpy(oks.data(), d_oks.p, (size_t)R * 4, cudaMemcpyDeviceToHost));
  d_ok.release(); d_oks.release();
  for (int v : oks) ok = ok && v;
  if (!ok) {
    close_imports(h);
    if (me == 0) fprintf(stderr, "parafem_b200: CUDA IPC peer mapping unavailable, halo exchange stays on NCCL send/recv\n");
    return 0;
  }
  CU(cudaMemcpy(h->ptab.p, &T, sizeof T, cudaMemcpyHostToDevice));
  h->host_tab = T;
  h->ptab_valid = true;
  h->peer_ok = true;
  return 0;
}

// Unanimous verdict before any rank leaves a section that the others continue with collectives: every rank
// contributes its local status; if any failed, ALL return an error (the failing rank its own, the others status 20)
// instead of blocking for ever in the next all-gather / send-recv / flag wait.
int agree(pf_handle h, int local_rc) {
  if (h->nranks == 1 || !h->comm) return local_rc;
  const std::string mine = h->err;
  DevBuf<int> d_one, d_all;
  if (d_one.alloc(1) != cudaSuccess || d_all.alloc((size_t)h->nranks) != cudaSuccess) return local_rc ? local_rc : 10;
  std::vector<int> all((size_t)h->nranks, 0);
  bool ok = cudaMemcpy(d_one.p, &local_rc, sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess &&
            g_nccl.AllGather(d_one.p, d_all.p, 1, ncclInt32, h->comm, h->stream) == ncclSuccess &&
            cudaStreamSynchronize(h->stream) == cudaSuccess &&
            cudaMemcpy(all.data(), d_all.p, all.size() * sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess;
  d_one.release(); d_all.release();
  if (!ok) return local_rc ? local_rc : fail(h, 11, "agree: status exchange failed");
  if (local_rc) { h->err = mine; return local_rc; }
  for (int r = 0; r < h->nranks; ++r)
    if (all[(size_t)r]) return fail(h, 20, "rank %d failed (status %d) in a collective section; this rank returns with it", r, all[(size_t)r]);
  return 0;
}

// tables of the exchanges fused into k_pupdate (forward) and k_dot (reverse); needs h->host_tab (setup_peer)
int build_fused_tables(pf_handle h, const std::vector<int> &put_slot_host, const std::vector<int> &acc_slot_host) {
  const PeerTable &T = h->host_tab;
  const int R = h->nranks;
  std::vector<int64_t> off((size_t)R + 1), dst_off((size_t)R);
  for (int r = 0; r < R; ++r) { off[(size_t)r] = h->put_off[r]; dst_off[(size_t)r] = T.fwd_dst_off[r]; }
  off[(size_t)R] = h->nput;
  const size_t np = (size_t)std::max<int64_t>(h->nput, 1);
  std::vector<unsigned int> bits((size_t)((h->neq_pp + 31) / 32 + 1)), ptr(np + 1);
  std::vector<int> slot0(np), rank(np);
  std::vector<int64_t> dst(np);
  int64_t nu = 0;
  if (pf_make_put_tables(R, h->neq_pp, off.data(), put_slot_host.data(), dst_off.data(), bits.data(), slot0.data(), ptr.data(),
                         rank.data(), dst.data(), &nu))
    return fail(h, 4, "pf_setup_mesh: forward-exchange tables could not be built");
  h->npk = (int)nu;
  CU(h->put_bits.alloc(bits.size())); CU(h->pk_ptr.alloc(ptr.size()));
  CU(h->pk_slot0.alloc(np)); CU(h->pk_rank.alloc(np)); CU(h->pk_dst.alloc(np));
  CU(cudaMemcpy(h->put_bits.p, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->pk_ptr.p, ptr.data(), ptr.size() * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->pk_slot0.p, slot0.data(), np * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->pk_rank.p, rank.data(), np * 4, cudaMemcpyHostToDevice));
  static_assert(sizeof(long long) == sizeof(int64_t), "pk_dst");
  CU(cudaMemcpy(h->pk_dst.p, dst.data(), np * 8, cudaMemcpyHostToDevice));
  const size_t nchunks = (size_t)((h->neq_pp + kChunk - 1) / kChunk);
  std::vector<unsigned int> cptr(nchunks + 1, 0u);
  if (pf_make_acc_chunks(h->neq_pp, kChunk, (int64_t)acc_slot_host.size(), acc_slot_host.data(), cptr.data()))
    return fail(h, 4, "pf_setup_mesh: accumulate table is not ascending");
  CU(h->acc_chunk_ptr.alloc(cptr.size()));
  CU(cudaMemcpy(h->acc_chunk_ptr.p, cptr.data(), cptr.size() * 4, cudaMemcpyHostToDevice));
  return 0;
}

// forward halo exchange: owners send the p values their peers' elements need
int halo_forward(pf_handle h, double *vec_ext, const State *st, bool peer = false) {
  if (h->nranks == 1) return 0;
  Scope sc(h, K_HALO);
  if (peer && vec_ext == h->p_ext.p) {
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((h->nput + 255) / 256, 2 * h->sm_count));
    k_halo_put_peer<<<grid, 256, 0, h->stream>>>(h->ptab.p, h->put_slot.p, vec_ext, (long long)h->nput, st);
    k_halo_wait<<<1, 32, 0, h->stream>>>(h->ptab.p, 0, st);
    h->launches += 2;
    CU(cudaGetLastError());
    return 0;
  }
  if (h->nput > 0) {
    k_halo_pack<<<(int)((h->nput + 255) / 256), 256, 0, h->stream>>>(h->put_slot.p, vec_ext, h->sendbuf.p, (int)h->nput, st);
    h->launches++;
  }
  NC(g_nccl.GroupStart());
  for (int r = 0; r < h->nranks; ++r) {
    if (r == h->rank) continue;
    if (h->put_cnt[r] > 0) NC(g_nccl.Send(h->sendbuf.p + h->put_off[r], (size_t)h->put_cnt[r], ncclDouble, r, h->comm, h->stream));
    if (h->get_cnt[r] > 0) NC(g_nccl.Recv(vec_ext + 1 + h->neq_pp + h->get_off[r], (size_t)h->get_cnt[r], ncclDouble, r, h->comm, h->stream));
  }
  NC(g_nccl.GroupEnd());
  return 0;
}

// reverse halo exchange: partial sums of remote equations go to their owners, which add
// them after their own partial sum, sources in ascending rank order
int halo_reverse(pf_handle h, double *vec_ext, const State *st, bool peer = false) {
  if (h->nranks == 1) return 0;
  Scope sc(h, K_HALO);
  if (peer) {
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((h->nhalo + 255) / 256, 2 * h->sm_count));
    k_halo_rev_peer<<<grid, 256, 0, h->stream>>>(h->ptab.p, vec_ext, (long long)h->neq_pp, (long long)h->nhalo, st);
    k_halo_wait<<<1, 32, 0, h->stream>>>(h->ptab.p, 1, st);
    h->launches += 2;
    if (h->nacc > 0) {
      k_halo_accumulate<<<(h->nacc + 255) / 256, 256, 0, h->stream>>>(h->acc_slot.p, h->acc_ptr.p, h->acc_pos.p, h->recvbuf.p, vec_ext, h->nacc, st);
      h->launches++;
    }
    CU(cudaGetLastError());
    return 0;
  }
  NC(g_nccl.GroupStart());
  for (int r = 0; r < h->nranks; ++r) {
    if (r == h->rank) continue;
    if (h->get_cnt[r] > 0) NC(g_nccl.Send(vec_ext + 1 + h->neq_pp + h->get_off[r], (size_t)h->get_cnt[r], ncclDouble, r, h->comm, h->stream));
    if (h->put_cnt[r] > 0) NC(g_nccl.Recv(h->recvbuf.p + h->put_off[r], (size_t)h->put_cnt[r], ncclDouble, r, h->comm, h->stream));
  }
  NC(g_nccl.GroupEnd());
  if (h->nacc > 0) {
    k_halo_accumulate<<<(h->nacc + 255) / 256, 256, 0, h->stream>>>(h->acc_slot.p, h->acc_ptr.p, h->acc_pos.p, h->recvbuf.p, vec_ext, h->nacc, st);
    h->launches++;
  }
  CU(cudaGetLastError());
  return 0;
}

// all-gather of the ranks' [dot, max, max, 0] partials + fixed-order combine
int combine_scalars(pf_handle h, int mode) {
  if (h->nranks == 1) return 0;
  State *st = h->state.p;
  NC(g_nccl.AllGather(st->loc, h->gath.p, 4, ncclDouble, h->comm, h->stream));
  k_scalars<<<1, 32, 0, h->stream>>>(st, h->gath.p, h->nranks, mode, h->ratio_hist.p);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

// grid of a chunk-looping reduction kernel: exactly the blocks that are resident at once (a second
// partial wave would leave a tail); the reduction tree does not depend on the grid size
template <class K>
int vec_grid(pf_handle h, K kernel) {
  const int per_sm = resident_blocks(h, kernel, kRedThreads, 4);
  const int64_t nchunks = (h->neq_pp + kChunk - 1) / kChunk;
  return (int)std::max<int64_t>(1, std::min<int64_t>(nchunks, (int64_t)h->sm_count * per_sm));
}

// u_ext = A p_ext over this rank's elements + halo exchanges (gather, mat-vec, scatter)
int apply_operator(pf_handle h, const State *st, bool peer = false) {
  int rc;
  if ((rc = halo_forward(h, h->p_ext.p, st, peer))) return rc;
  if ((rc = launch_matvec<true>(h, h->p_ext.p, st))) return rc;
  if ((rc = launch_scatter(h, st, false, h->u_ext.p))) return rc;
  if ((rc = halo_reverse(h, h->u_ext.p, st, peer))) return rc;
  if (h->nfixed > 0) {
    k_fixed_u<<<(h->nfixed + 255) / 256, 256, 0, h->stream>>>(h->fix_slot.p, h->store.p, h->p_ext.p, h->u_ext.p, h->nfixed, st, h->fixed_mode);
    h->launches++;
  }
  return 0;
}

bool fused_peer() {
  static const bool on = !(getenv("PF_FUSE") && !strcmp(getenv("PF_FUSE"), "0"));
  return on;
}
AccTables acc_tables(pf_handle h) {
  return AccTables{h->acc_slot.p, h->acc_ptr.p, h->acc_pos.p, h->acc_chunk_ptr.p, h->recvbuf.p, h->u_ext.p};
}
PutTables put_tables(pf_handle h) {
  return PutTables{h->put_bits.p, h->pk_slot0.p, h->pk_ptr.p, h->pk_rank.p, h->pk_dst.p, h->npk};
}

// One PCG iteration (p121.f90:91-103).  Peer transport: five launches, as on one rank -- the forward exchange of
// the NEXT iteration rides on k_pupdate, k_matvec waits for it in its prologue, the reverse exchange rides on
// k_scatter and k_dot adds the received partial sums before reducing.  NCCL transport: the separate pack / send-recv /
// accumulate / all-gather steps.
int one_iteration(pf_handle h) {
  State *st = h->state.p;
  const int single = h->nranks == 1;
  const long long n = h->neq_pp;
  const bool peer = h->peer_ok && h->use_peer;
  PeerTable *T = peer ? h->ptab.p : nullptr;
  int rc;
  if (peer && !fused_peer()) {
    // PF_FUSE=0: the first build of the peer transport (put / wait / accumulate kernels of their own), kept for A/B runs
    if ((rc = apply_operator(h, st, true))) return rc;
    Scope sc(h, K_VECTOR);
    k_dot<false><<<vec_grid(h, k_dot<false>), kRedThreads, 0, h->stream>>>(h->p_ext.p + 1, h->u_ext.p + 1, n, h->part.p, st, single, 1, T, AccTables{});
    k_pcg_update<<<vec_grid(h, k_pcg_update), kRedThreads, 0, h->stream>>>(h->diag_ext.p + 1, h->p_ext.p + 1, h->u_ext.p + 1, h->x.p, h->r.p,
                                                              h->d.p, n, h->part.p, st, single, h->ratio_hist.p, T);
    k_pupdate<<<grid_for(h, (n + 1) / 2, 256, 8), 256, 0, h->stream>>>(h->d.p, h->p_ext.p + 1, n, st, nullptr, PutTables{});
    h->launches += 3;
    CU(cudaGetLastError());
    return 0;
  }
  if (peer) {
    if ((rc = launch_matvec<true>(h, h->p_ext.p, st, T))) return rc;
    if ((rc = launch_scatter(h, st, false, h->u_ext.p, T))) return rc;
    if (h->nfixed > 0) {   // owned rows only; the partial sums of others for these rows are overridden as well
      k_halo_wait<<<1, 32, 0, h->stream>>>(T, 1, st);
      if (h->nacc > 0) k_halo_accumulate<<<(h->nacc + 255) / 256, 256, 0, h->stream>>>(h->acc_slot.p, h->acc_ptr.p, h->acc_pos.p, h->recvbuf.p, h->u_ext.p, h->nacc, st);
      k_fixed_u<<<(h->nfixed + 255) / 256, 256, 0, h->stream>>>(h->fix_slot.p, h->store.p, h->p_ext.p, h->u_ext.p, h->nfixed, st, h->fixed_mode);
      h->launches += 2 + (h->nacc > 0);
    }
  } else if ((rc = apply_operator(h, st, false))) return rc;
  {
    Scope sc(h, K_VECTOR);
    if (peer && h->nfixed == 0)
      k_dot<true><<<vec_grid(h, k_dot<true>), kRedThreads, 0, h->stream>>>(h->p_ext.p + 1, h->u_ext.p + 1, n, h->part.p, st, single, 1, T, acc_tables(h));
    else
      k_dot<false><<<vec_grid(h, k_dot<false>), kRedThreads, 0, h->stream>>>(h->p_ext.p + 1, h->u_ext.p + 1, n, h->part.p, st, single, 1, T, AccTables{});
    h->launches++;
    if (!peer && (rc = combine_scalars(h, 1))) return rc;
    k_pcg_update<<<vec_grid(h, k_pcg_update), kRedThreads, 0, h->stream>>>(h->diag_ext.p + 1, h->p_ext.p + 1, h->u_ext.p + 1, h->x.p, h->r.p,
                                                              h->d.p, n, h->part.p, st, single, h->ratio_hist.p, T);
    h->launches++;
    if (!peer && (rc = combine_scalars(h, 2))) return rc;
    k_pupdate<<<grid_for(h, (n + 1) / 2, 256, 8), 256, 0, h->stream>>>(h->d.p, h->p_ext.p + 1, n, st, T, peer ? put_tables(h) : PutTables{});
    h->launches++;
  }
  CU(cudaGetLastError());
  return 0;
}

int need_device(pf_handle h) {
  if (!h) return fail(nullptr, 1, "null handle");
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) return fail(h, 10, "cudaSetDevice(%d): %s", h->device, cudaGetErrorString(e));
  return 0;
}

}  // namespace

extern "C" {

int pf_version(void) { return 100; }

int pf_last_error(pf_handle h, char *buf, int len) {
  const std::string &s = h ? h->err : g_last_error;
  if (buf && len > 0) { strncpy(buf, s.c_str(), (size_t)len - 1); buf[len - 1] = 0; }
  return (int)s.size();
}

int pf_nccl_unique_id(void *id128) {
  pf_handle h = nullptr;
  std::string err;
  if (!g_nccl.load(err)) return fail(h, 12, "%s", err.c_str());
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  NC(g_nccl.GetUniqueId(reinterpret_cast<ncclUniqueId *>(id128)));
  return 0;
}

int pf_init(int rank, int nranks, int device, const void *id128, pf_handle *out) {
  pf_handle h = nullptr;
  if (!out || nranks < 1 || rank < 0 || rank >= nranks) return fail(h, 2, "pf_init: bad arguments");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(h, 10, "pf_init: no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(h, 2, "pf_init: device %d out of range (%d devices)", device, ndev);
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(h, 13, "pf_init: device is sm_%d%d; this build targets sm_100a (B200) only", prop.major, prop.minor);
  h = new pf_ctx();
  h->rank = rank; h->nranks = nranks; h->device = device; h->sm_count = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CU(h->state.alloc(1));
  CU(cudaMemset(h->state.p, 0, sizeof(State)));
  CU(h->gath.alloc((size_t)4 * nranks));
  CU(cudaMallocHost((void **)&h->snap_pinned, 2 * sizeof(State)));
  CU(cudaEventCreateWithFlags(&h->snap_ev[0], cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&h->snap_ev[1], cudaEventDisableTiming));
  if (nranks > 1) {
    std::string err;
    if (!g_nccl.load(err)) { int rc = fail(h, 12, "%s", err.c_str()); delete h; return rc; }
    if (!id128) { delete h; return fail(nullptr, 2, "pf_init: nranks > 1 needs the 128-byte NCCL id"); }
    ncclUniqueId id; memcpy(&id, id128, sizeof id);
    NC(g_nccl.CommInitRank(&h->comm, nranks, id, rank));
    CU(h->sync.alloc(1)); CU(cudaMemset(h->sync.p, 0, sizeof(PeerSync)));
    CU(h->ptab.alloc(1)); CU(cudaMemset(h->ptab.p, 0, sizeof(PeerTable)));
    const char *mode = getenv("PF_HALO");     // "nccl": keep every exchange on NCCL send/recv + all-gather
    h->use_peer = !(mode && !strcmp(mode, "nccl"));
  }
  *out = h;
  return 0;
}

int pf_finalize(pf_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  collect_spans(h);
  for (auto e : h->pool) cudaEventDestroy(e);
  if (h->device >= 0 && h->device < 64 && g_ctab_owner[h->device] == h) g_ctab_owner[h->device] = nullptr;
  if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
  if (h->snap_pinned) cudaFreeHost(h->snap_pinned);
  for (auto e : h->snap_ev) if (e) cudaEventDestroy(e);
  h->kc.release(); h->fext.release(); h->x0.release(); h->d1x0.release(); h->d2x0.release();
  h->evpt.release(); h->tensor.release(); h->bdylds.release(); h->oldis.release(); h->totd.release(); h->loads.release(); h->ld0.release(); h->valf.release();
  h->put_bits.release(); h->pk_ptr.release(); h->acc_chunk_ptr.release(); h->pk_slot0.release(); h->pk_rank.release(); h->pk_dst.release();
  close_imports(h);
  if (h->comm) {
    // peers must have closed their mappings of my buffers before I free them
    DevBuf<int> a, b;
    if (a.alloc(1) == cudaSuccess && b.alloc((size_t)h->nranks) == cudaSuccess) {
      g_nccl.AllGather(a.p, b.p, 1, ncclInt32, h->comm, h->stream);
      cudaStreamSynchronize(h->stream);
    }
    a.release(); b.release();
    g_nccl.CommDestroy(h->comm);
  }
  h->sync.release(); h->ptab.release();
  h->coord.release(); h->km.release(); h->kb.release(); h->val_f.release(); h->diag_tmp.release(); h->geom.release(); h->utemp.release(); h->ggl.release(); h->csr_ptr.release(); h->csr_pos.release();
  h->put_slot.release(); h->sendbuf.release(); h->recvbuf.release(); h->acc_slot.release(); h->acc_ptr.release(); h->acc_pos.release();
  h->p_ext.release(); h->u_ext.release(); h->diag_ext.release(); h->r.release(); h->x.release(); h->d.release();
  h->part.release(); h->gath.release(); h->state.release(); h->ratio_hist.release(); h->fix_slot.release(); h->store.release();
  cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int pf_device_info(pf_handle h, int *sm_count, int64_t *free_bytes, int64_t *total_bytes) {
  int rc = need_device(h); if (rc) return rc;
  size_t f = 0, t = 0;
  CU(cudaMemGetInfo(&f, &t));
  if (sm_count) *sm_count = h->sm_count;
  if (free_bytes) *free_bytes = (int64_t)f;
  if (total_bytes) *total_bytes = (int64_t)t;
  return 0;
}

int64_t pf_kernel_launches(pf_handle h) { return h ? h->launches : 0; }

int pf_measure_fp64(pf_handle h, double *tflops) {
  int rc = need_device(h); if (rc) return rc;
  DevBuf<double> out; CU(out.alloc(1));
  const int iters = 20000, threads = 256, blocks = h->sm_count * 8;
  cudaEvent_t e0, e1; CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    CU(cudaEventRecord(e0, h->stream));
    k_fp64_peak<<<blocks, threads, 0, h->stream>>>(out.p, iters, 1.0000001, 1e-9);
    CU(cudaEventRecord(e1, h->stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f; CU(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = 2.0 * 8.0 * iters * (double)threads * blocks / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  h->launches += 4;
  cudaEventDestroy(e0); cudaEventDestroy(e1); out.release();
  *tflops = best;
  return 0;
}

// The mat-vec kernel of the current problem launched `reps` times back to back between ONE pair of events: the kernel
// time without the ~10 us a pair of per-launch event records adds to a 0.1 ms launch (config B).  The right-hand sides
// are whatever p_ext holds (local values; no halo wait, no stopping flag): timing only, utemp is overwritten.
int pf_measure_matvec(pf_handle h, int reps, double *ms_per_launch) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->have_km && h->have_mesh && reps > 0, "needs pf_setup_mesh, element matrices and reps > 0");
  if (h->matrix_free && (rc = ensure_tables(h))) return rc;
  struct NoProfile {                       // per-launch events off for the duration, restored on every return path
    pf_handle h; bool was;
    explicit NoProfile(pf_handle h_) : h(h_), was(h_->profile) { h->profile = false; }
    ~NoProfile() { h->profile = was; }
  } guard(h);
  EventPair ev; CU(ev.create());
  for (int warm = 0; warm < 2 && !rc; ++warm) rc = launch_matvec<true>(h, h->p_ext.p, nullptr);
  if (rc) return rc;
  CU(cudaEventRecord(ev.a, h->stream));
  for (int i = 0; i < reps && !rc; ++i) rc = launch_matvec<true>(h, h->p_ext.p, nullptr);
  CU(cudaEventRecord(ev.b, h->stream));
  CU(cudaEventSynchronize(ev.b));
  if (rc) return rc;
  float ms = 0.f; CU(cudaEventElapsedTime(&ms, ev.a, ev.b));
  *ms_per_launch = (double)ms / reps;
  return 0;
}

int pf_measure_fp64_tensor(pf_handle h, double *tflops) {
  int rc = need_device(h); if (rc) return rc;
  DevBuf<double> out; CU(out.alloc(1));
  const int iters = 4000, threads = 256, blocks = h->sm_count * 2;
  EventPair ev; CU(ev.create());
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    CU(cudaEventRecord(ev.a, h->stream));
    k_fp64_tensor_peak<<<blocks, threads, 0, h->stream>>>(out.p, iters, 1.0000001, 1e-9);
    CU(cudaEventRecord(ev.b, h->stream));
    CU(cudaEventSynchronize(ev.b));
    float ms = 0.f; CU(cudaEventElapsedTime(&ms, ev.a, ev.b));
    const double tf = 2.0 * 256.0 * 9.0 * iters * (double)(threads / 32) * blocks / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  h->launches += 4;
  *tflops = best;
  return 0;
}

int pf_measure_hbm_read(pf_handle h, double *gbs) {
  int rc = need_device(h); if (rc) return rc;
  constexpr int kTile = 28800, kStages = 7;                 // the hex20 tile shape of k_matvec
  const size_t smem = (size_t)kStages * kTile + kStages * 8;
  auto kern = k_stream_read<kTile, kStages>;
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // stream over the element matrices when they are resident (no extra memory), else over a scratch buffer
  DevBuf<double> scratch;
  const double *src = h->km.p;
  size_t bytes = h->km.bytes();
  if (!src || bytes < ((size_t)1 << 30)) {
    size_t f = 0, t = 0;
    CU(cudaMemGetInfo(&f, &t));
    bytes = std::min<size_t>((size_t)8 << 30, f / 2);
    CU(scratch.alloc(bytes / 8));
    CU(cudaMemsetAsync(scratch.p, 0, bytes, h->stream));
    src = scratch.p;
  }
  const long long ntiles = (long long)(bytes / kTile);
  DevBuf<double> out; CU(out.alloc(1));
  cudaEvent_t e0, e1; CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    CU(cudaEventRecord(e0, h->stream));
    kern<<<h->sm_count, kStages * 32, smem, h->stream>>>(src, ntiles, out.p);
    CU(cudaEventRecord(e1, h->stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f; CU(cudaEventElapsedTime(&ms, e0, e1));
    const double g = (double)ntiles * kTile / (ms * 1e-3) / 1e9;
    if (rep > 0 && g > best) best = g;
  }
  h->launches += 4;
  CU(cudaGetLastError());
  cudaEventDestroy(e0); cudaEventDestroy(e1); out.release(); scratch.release();
  *gbs = best;
  return 0;
}

int pf_set_profile(pf_handle h, int on) { if (!h) return 1; h->profile = on != 0; return 0; }
int pf_reset_profile(pf_handle h) {
  if (!h) return 1;
  cudaStreamSynchronize(h->stream);
  collect_spans(h);
  for (int k = 0; k < K_NKINDS; ++k) { h->kind_ms[k] = 0; h->kind_n[k] = 0; }
  return 0;
}
int pf_get_kernel_ms(pf_handle h, int which, double *total_ms, int64_t *launches) {
  if (!h || which < 0 || which >= K_NKINDS) return fail(h, 2, "pf_get_kernel_ms: bad arguments");
  CU(cudaStreamSynchronize(h->stream));
  collect_spans(h);
  if (total_ms) *total_ms = h->kind_ms[which];
  if (launches) *launches = h->kind_n[which];
  return 0;
}

int pf_setup_mesh(pf_handle h, int nod, int nodof, int nip, int64_t nels_pp, const double *g_coord_pp,
                  const int32_t *g_g_pp, int64_t neq, int64_t ieq_start, int64_t neq_pp) {
  int rc = need_device(h); if (rc) return rc;
  const int ntot = nod * nodof;
  const int64_t total = nels_pp * ntot;
  std::vector<int32_t> halo;
  std::vector<int64_t> halo_cnt((size_t)h->nranks, 0);
  int64_t nhalo = 0;

  // ---- section A, local: argument checks, gather table, scatter tables on the device, vectors.  Several ranks:
  // nobody enters the collectives below unless every rank got through (agree) ----
  auto local_a = [&]() -> int {
    NEED((nod == 4 || nod == 8 || nod == 20) && (nodof == 1 || nodof == 3), "nod must be 4 (tetrahedra), 8 or 20 (hexahedra), nodof 1 or 3");
    NEED(nip == 1 || nip == 8 || (nod == 4 && (nip == 4 || nip == 5)) || (nod == 20 && nip == 27),
         "nip must be 1 or 8 (hexahedra; 27 for 20-node bricks), 1, 4 or 5 (tetrahedra)");
    NEED(nod != 4 || nip != 8, "4-node tetrahedra take nip = 1, 4 or 5");
    NEED(nels_pp >= 1 && neq >= 1 && neq_pp >= 0 && ieq_start >= 1, "bad sizes");
    NEED(ntot == 60 || ntot == 24 || ntot == 8 || ntot == 12 || ntot == 4,
         "supported element types: hex20 / hex8 / tet4 elastic, hex8 / tet4 scalar");
    NEED(nels_pp * ntot < (int64_t)0xffffffffu, "nels_pp*ntot exceeds 32-bit table range; use more ranks");
    {
      int64_t c, s;
      pf_calc_neq_pp(neq, h->nranks, h->rank + 1, &c, &s);
      NEED(c == neq_pp && s == ieq_start, "neq_pp/ieq_start do not match calc_neq_pp for this rank");
    }
    return 0;
  };
  if ((rc = agree(h, local_a()))) return rc;
  if (h->nranks > 1) {
    // re-setup: drop the mappings of the peers' old buffers, and make sure every peer has
    // dropped its mappings of mine, before anything is freed
    close_imports(h);
    DevBuf<int> a, b;
    CU(a.alloc(1)); CU(b.alloc((size_t)h->nranks));
    NC(g_nccl.AllGather(a.p, b.p, 1, ncclInt32, h->comm, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    a.release(); b.release();
  }
  auto local_b = [&]() -> int {
    h->nod = nod; h->nodof = nodof; h->nip = nip; h->ntot = ntot;
    h->nels = nels_pp; h->neq = neq; h->ieq_start = ieq_start; h->neq_pp = neq_pp;
    h->have_mesh = false; h->have_km = h->have_precon = false;
    h->transient = h->transient_first = false; h->mat_override = nullptr; h->kb.release();
    h->explicit_ = false; h->plastic = false; h->fixed_mode = 0; h->one_km = false; h->dynamic = false; h->kc.release();
    h->epoch++;                                   // a captured iteration graph of the previous mesh is stale

    // gather table (make_ggl rebuilt from g_g_pp)
    std::vector<int32_t> ggl((size_t)total);
    if (pf_make_ggl(ntot, nels_pp, g_g_pp, neq, h->nranks, h->rank + 1, nullptr, 0, nullptr, halo_cnt.data(), &nhalo))
      return fail(h, 4, "pf_setup_mesh: g_g_pp holds equation numbers outside [0, neq]");
    halo.resize((size_t)std::max<int64_t>(nhalo, 1));
    if (pf_make_ggl(ntot, nels_pp, g_g_pp, neq, h->nranks, h->rank + 1, ggl.data(), (int64_t)halo.size(), halo.data(),
                    halo_cnt.data(), &nhalo))
      return fail(h, 4, "pf_setup_mesh: gather table construction failed");
    h->nhalo = nhalo;
    h->nslots = 1 + neq_pp + nhalo;
    NEED(h->nslots < (int64_t)0x7fffffff, "slot count exceeds int32");

    // CSR of contributions per slot in ascending element order: built on the device below (stable radix sort)
    CU(h->coord.alloc((size_t)nels_pp * nod * 3));
    CU(cudaMemcpy(h->coord.p, g_coord_pp, h->coord.bytes(), cudaMemcpyHostToDevice));
    CU(h->ggl.alloc((size_t)total));
    CU(cudaMemcpy(h->ggl.p, ggl.data(), h->ggl.bytes(), cudaMemcpyHostToDevice));
    {
      // csr_pos = positions e*ntot+k sorted by slot, ascending inside a slot (the sort is stable), the slot-0
      // (restrained) entries dropped; csr_ptr[s] = first position of slot s.  117 M pairs at config C: a few ms
      // instead of ~0.7 s of random-access host passes.  (DevBuf releases on every return path.)
      DevBuf<int> keys;
      DevBuf<unsigned int> iota, vals;
      DevBuf<unsigned char> tmp;
      CU(keys.alloc((size_t)total)); CU(iota.alloc((size_t)total)); CU(vals.alloc((size_t)total));
      k_iota<<<grid_for(h, total, 256), 256, 0, h->stream>>>(iota.p, (long long)total);
      int end_bit = 1;
      while (end_bit < 31 && ((int64_t)1 << end_bit) < h->nslots) ++end_bit;
      size_t tmp_bytes = 0;
      CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, h->ggl.p, keys.p, iota.p, vals.p, (int64_t)total, 0, end_bit, h->stream));
      CU(tmp.alloc(tmp_bytes));
      CU(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, h->ggl.p, keys.p, iota.p, vals.p, (int64_t)total, 0, end_bit, h->stream));
      CU(h->csr_ptr.alloc((size_t)h->nslots + 1));
      k_csr_ptr<<<grid_for(h, h->nslots + 1, 256), 256, 0, h->stream>>>(keys.p, (long long)total, h->csr_ptr.p, (long long)h->nslots);
      h->launches += 2;
      CU(cudaGetLastError());
      unsigned int n0 = 0, nnz = 0;       // entries of slot 0, entries of slots >= 1
      CU(cudaMemcpyAsync(&nnz, h->csr_ptr.p + h->nslots, 4, cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
      n0 = (unsigned int)total - nnz;
      CU(h->csr_pos.alloc(std::max<size_t>(nnz, 1)));
      CU(cudaMemcpyAsync(h->csr_pos.p, vals.p + n0, (size_t)nnz * 4, cudaMemcpyDeviceToDevice, h->stream));
      CU(cudaStreamSynchronize(h->stream));
    }
    CU(h->utemp.alloc((size_t)total));

    const size_t ns = (size_t)h->nslots, nq = (size_t)std::max<int64_t>(neq_pp, 1);
    CU(h->p_ext.alloc(ns)); CU(h->u_ext.alloc(ns)); CU(h->diag_ext.alloc(ns));
    CU(h->r.alloc(nq)); CU(h->x.alloc(nq)); CU(h->d.alloc(nq));
    CU(cudaMemset(h->p_ext.p, 0, ns * 8)); CU(cudaMemset(h->u_ext.p, 0, ns * 8)); CU(cudaMemset(h->diag_ext.p, 0, ns * 8));
    CU(cudaMemset(h->x.p, 0, nq * 8));
    const size_t nchunks = (size_t)((neq_pp + kChunk - 1) / kChunk);
    CU(h->part.alloc(3 * std::max<size_t>(nchunks, 1)));
    return 0;
  };
  if ((rc = agree(h, local_b()))) return rc;

  // halo tables: what I get from each owner is known; tell each owner what to put
  h->get_cnt.assign((size_t)h->nranks, 0); h->get_off.assign((size_t)h->nranks, 0);
  h->put_cnt.assign((size_t)h->nranks, 0); h->put_off.assign((size_t)h->nranks, 0);
  h->nput = 0; h->nacc = 0; h->npk = 0;
  if (h->nranks > 1) {
    const int R = h->nranks;
    int64_t off = 0;
    for (int r = 0; r < R; ++r) { h->get_cnt[r] = halo_cnt[r]; h->get_off[r] = off; off += halo_cnt[r]; }
    // ---- section B, collective: counts matrix by all-gather, wanted equation numbers by grouped send/recv ----
    DevBuf<int64_t> d_cnt, d_all;
    CU(d_cnt.alloc((size_t)R)); CU(d_all.alloc((size_t)R * R));
    CU(cudaMemcpy(d_cnt.p, h->get_cnt.data(), (size_t)R * 8, cudaMemcpyHostToDevice));
    NC(g_nccl.AllGather(d_cnt.p, d_all.p, (size_t)R, ncclInt64, h->comm, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    std::vector<int64_t> all((size_t)R * R);
    CU(cudaMemcpy(all.data(), d_all.p, all.size() * 8, cudaMemcpyDeviceToHost));
    d_cnt.release(); d_all.release();
    off = 0;
    for (int r = 0; r < R; ++r) { h->put_cnt[r] = all[(size_t)r * R + h->rank]; h->put_off[r] = off; off += h->put_cnt[r]; }
    h->nput = off;
    DevBuf<int> d_want, d_asked;
    CU(d_want.alloc((size_t)std::max<int64_t>(nhalo, 1))); CU(d_asked.alloc((size_t)std::max<int64_t>(h->nput, 1)));
    CU(cudaMemcpy(d_want.p, halo.data(), (size_t)nhalo * 4, cudaMemcpyHostToDevice));
    NC(g_nccl.GroupStart());
    for (int r = 0; r < R; ++r) {
      if (r == h->rank) continue;
      if (h->get_cnt[r] > 0) NC(g_nccl.Send(d_want.p + h->get_off[r], (size_t)h->get_cnt[r], ncclInt32, r, h->comm, h->stream));
      if (h->put_cnt[r] > 0) NC(g_nccl.Recv(d_asked.p + h->put_off[r], (size_t)h->put_cnt[r], ncclInt32, r, h->comm, h->stream));
    }
    NC(g_nccl.GroupEnd());
    CU(cudaStreamSynchronize(h->stream));
    std::vector<int> asked((size_t)std::max<int64_t>(h->nput, 1));
    CU(cudaMemcpy(asked.data(), d_asked.p, (size_t)h->nput * 4, cudaMemcpyDeviceToHost));
    d_want.release(); d_asked.release();
    // ---- section C, local: validate what the peers asked for, build the put / accumulate tables ----
    std::vector<int> aslot;
    auto local_c = [&]() -> int {
      NEED(h->put_cnt[h->rank] == 0, "a rank lists its own equations as remote");
      for (int64_t k = 0; k < h->nput; ++k) {
        const int64_t g = asked[(size_t)k];
        NEED(g >= ieq_start && g < ieq_start + neq_pp, "peer asked for an equation this rank does not own");
        asked[(size_t)k] = (int)(g - ieq_start + 1);  // -> slot
      }
      CU(h->put_slot.alloc(asked.size()));
      CU(cudaMemcpy(h->put_slot.p, asked.data(), asked.size() * 4, cudaMemcpyHostToDevice));
      CU(h->sendbuf.alloc(asked.size())); CU(h->recvbuf.alloc(asked.size()));
      // accumulate table: per owned slot, receive-buffer positions in ascending source rank
      // (recvbuf is grouped by source rank ascending, so ascending position == ascending rank)
      std::vector<std::pair<int, unsigned int>> pairs((size_t)h->nput);
      for (int64_t k = 0; k < h->nput; ++k) pairs[(size_t)k] = {asked[(size_t)k], (unsigned int)k};
      std::sort(pairs.begin(), pairs.end());
      std::vector<unsigned int> aptr, apos;
      for (size_t k = 0; k < pairs.size(); ++k) {
        if (k == 0 || pairs[k].first != pairs[k - 1].first) { aslot.push_back(pairs[k].first); aptr.push_back((unsigned int)k); }
        apos.push_back(pairs[k].second);
      }
      aptr.push_back((unsigned int)pairs.size());
      h->nacc = (int)aslot.size();
      CU(h->acc_slot.alloc(std::max<size_t>(aslot.size(), 1))); CU(h->acc_ptr.alloc(aptr.size())); CU(h->acc_pos.alloc(std::max<size_t>(apos.size(), 1)));
      CU(cudaMemcpy(h->acc_slot.p, aslot.data(), aslot.size() * 4, cudaMemcpyHostToDevice));
      CU(cudaMemcpy(h->acc_ptr.p, aptr.data(), aptr.size() * 4, cudaMemcpyHostToDevice));
      CU(cudaMemcpy(h->acc_pos.p, apos.data(), apos.size() * 4, cudaMemcpyHostToDevice));
      return 0;
    };
    if ((rc = agree(h, local_c()))) return rc;
    // ---- section D: peer mappings (collective, ends with its own unanimous verdict), then the tables of the fused
    // exchanges (local) ----
    if ((rc = setup_peer(h, all))) return rc;
    if ((rc = agree(h, h->peer_ok ? build_fused_tables(h, asked, aslot) : 0))) return rc;
  }
  h->have_mesh = true;
  return 0;
}

static size_t km_per_element(pf_handle h) {
  return h->km_layout == 1 ? (size_t)h->ntot * (h->ntot + 1) / 2 : (size_t)h->ntot * h->ntot;
}
static int alloc_km(pf_handle h) {
  const size_t n = (size_t)h->nels * km_per_element(h);
  if (h->km.n != n) CU(h->km.alloc(n));
  return 0;
}

int pf_form_km_elastic(pf_handle h, double e, double v) {
  int rc = need_device(h); if (rc) return rc;
  h->one_km = false;
  h->transient = false; h->explicit_ = false; h->dynamic = false; h->kb.release();
  NEED(h->have_mesh && h->nodof == 3, "needs pf_setup_mesh with nodof = 3");
  NEED(h->nod != 4 || (h->km_layout == 0 && !h->matrix_free), "tetrahedra: reference storkm layout, stored matrices only");
  ElemTables T;
  if (fill_tables(h->nod, h->nip, e, v, 0, 0, 0, T)) return fail(h, 3, "unsupported nod/nip");
  if ((rc = upload_tables(h, T))) return rc;
  double *diag_only = nullptr;
  if (h->matrix_free) {
    // config E: storkm is never stored; only its diagonal is formed (for the preconditioner)
    NEED(h->nip == 8, "the matrix-free variant supports nip = 8");
    h->km.release();
    CU(h->diag_tmp.alloc((size_t)h->nels * h->ntot));
    diag_only = h->diag_tmp.p;
    if (h->mf_mode == 2) {
      // geometric factors: inverse Jacobian (9) + det*w (1) per element and Gauss point
      CU(h->geom.alloc((size_t)((h->nels + 31) / 32) * 32 * 80));   // [group][point][word][lane], padded to whole groups
      // each kernel family writes the factors it reads: the tensor-core kernels invert with one reciprocal per point
      if (mf_kernel_choice() >= 3) rc = h->nod == 20 ? launch_mf3_t<20, false, 1>(h, nullptr, nullptr) : launch_mf3_t<8, false, 1>(h, nullptr, nullptr);
      else if (h->nod == 20 && mf_two_lanes()) rc = launch_mf2_t<20, false, 1>(h, nullptr, nullptr);
      else if (h->nod == 20) rc = launch_mf_t<20, false, 1>(h, nullptr, nullptr);
      else rc = launch_mf_t<8, false, 1>(h, nullptr, nullptr);
      if (rc) return rc;
    } else {
      // mode 1: the factors are rebuilt every call; one 640 B scratch line per resident element slot
      CU(h->geom.alloc((size_t)std::max(mf_grid(h) * kMfWarps, mf2_grid(h) * kMf2Warps) * 32 * 80));
    }
  } else if ((rc = alloc_km(h))) return rc;
  const int grid = (int)std::min<int64_t>(h->nels, (int64_t)h->sm_count * 16);
  // PF_FORM=old: the first (untiled) build of the full matrix, kept for comparison
  static const bool old_form = getenv("PF_FORM") && !strcmp(getenv("PF_FORM"), "old");
  if (h->nod == 4) {
    NEED(!diag_only, "the matrix-free variant exists for the hexahedra only");
    k_form_km_elastic<4, 64><<<grid, 64, 0, h->stream>>>(h->coord.p, h->km.p, (long long)h->nels, nullptr, h->km_layout);
  } else if (h->nip == 27) {
    NEED(!diag_only && h->nod == 20, "the 27-point rule is built for stored matrices of 20-node bricks");
    k_form_km_elastic<20, 128><<<grid, 128, 0, h->stream>>>(h->coord.p, h->km.p, (long long)h->nels, nullptr, h->km_layout);
  } else if (diag_only && !old_form) {
    if (h->nod == 20) k_form_km_tiled<20, 2, 2, 128, false, true><<<grid, 128, 0, h->stream>>>(h->coord.p, diag_only, (long long)h->nels, 0, nullptr, nullptr);
    else k_form_km_tiled<8, 1, 1, 64, false, true><<<grid, 64, 0, h->stream>>>(h->coord.p, diag_only, (long long)h->nels, 0, nullptr, nullptr);
  } else if (!diag_only && !old_form) {
    // hex20: 128-register cap = four resident CTAs per SM (84 B of spills) -- measured 45.5 ms against 59.0 ms with
    // three CTAs at 164 registers on the same box (scripts/gpu_round27.sh); PF_FORM=b3 selects the latter
    static const bool b3 = getenv("PF_FORM") && !strcmp(getenv("PF_FORM"), "b3");
    if (h->nod == 20 && !b3) k_form_km_tiled<20, 2, 2, 128, false, false, 4><<<grid, 128, 0, h->stream>>>(h->coord.p, h->km.p, (long long)h->nels, h->km_layout, nullptr, nullptr);
    else if (h->nod == 20) k_form_km_tiled<20, 2, 2, 128, false><<<grid, 128, 0, h->stream>>>(h->coord.p, h->km.p, (long long)h->nels, h->km_layout, nullptr, nullptr);
    else k_form_km_tiled<8, 1, 1, 64, false><<<grid, 64, 0, h->stream>>>(h->coord.p, h->km.p, (long long)h->nels, h->km_layout, nullptr, nullptr);
  } else if (h->nod == 20) k_form_km_elastic<20, 128><<<grid, 128, 0, h->stream>>>(h->coord.p, h->km.p, (long long)h->nels, diag_only, h->km_layout);
  else k_form_km_elastic<8, 64><<<grid, 64, 0, h->stream>>>(h->coord.p, h->km.p, (long long)h->nels, diag_only, h->km_layout);
  h->launches++;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->stream));
  h->have_km = true; h->have_precon = false;
  return 0;
}

// xx2.f90:169-193: as pf_form_km_elastic with e, v = prop(:,etype_pp(iel)) per element
int pf_form_km_elastic_mat(pf_handle h, int np_types, const double *prop, const int32_t *etype_pp) {
  int rc = need_device(h); if (rc) return rc;
  h->one_km = false;
  NEED(h->have_mesh && h->nodof == 3, "needs pf_setup_mesh with nodof = 3");
  NEED(!h->matrix_free, "the matrix-free variant takes one material (pf_form_km_elastic)");
  NEED(np_types >= 1 && prop && etype_pp, "np_types >= 1, prop(2,np_types) and etype_pp(nels_pp) are required");
  NEED(h->nod != 4 || h->km_layout == 0, "tetrahedra: reference storkm layout only");
  for (int64_t e = 0; e < h->nels; ++e)
    if (etype_pp[e] < 1 || etype_pp[e] > np_types) return fail(h, 4, "pf_form_km_elastic_mat: etype_pp(%lld) = %d outside 1..%d",
                                                               (long long)e + 1, etype_pp[e], np_types);
  h->transient = false; h->explicit_ = false; h->dynamic = false; h->kb.release();
  ElemTables T;
  std::vector<double> dees((size_t)np_types * 36);
  for (int m = 0; m < np_types; ++m) {           // deemat(e,v,dee) per material (xx2.f90:176-180)
    if (fill_tables(h->nod, h->nip, prop[2 * m], prop[2 * m + 1], 0, 0, 0, T)) return fail(h, 3, "unsupported nod/nip");
    memcpy(&dees[(size_t)m * 36], T.dee, sizeof T.dee);
  }
  if ((rc = upload_tables(h, T))) return rc;
  DevBuf<double> d_dee; DevBuf<int> d_etype;
  CU(d_dee.alloc(dees.size())); CU(d_etype.alloc((size_t)h->nels));
  CU(cudaMemcpy(d_dee.p, dees.data(), dees.size() * 8, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_etype.p, etype_pp, (size_t)h->nels * 4, cudaMemcpyHostToDevice));
  if ((rc = alloc_km(h))) return rc;
  const int grid = (int)std::min<int64_t>(h->nels, (int64_t)h->sm_count * 16);
  if (h->nod == 4) k_form_km_elastic<4, 64, true><<<grid, 64, 0, h->stream>>>(h->coord.p, h->km.p, (long long)h->nels, nullptr, h->km_layout, d_dee.p, d_etype.p);
  else if (h->nod == 20) k_form_km_tiled<20, 2, 2, 128, true, false, 4><<<grid, 128, 0, h->stream>>>(h->coord.p, h->km.p, (long long)h->nels, h->km_layout, d_dee.p, d_etype.p);
  else k_form_km_tiled<8, 1, 1, 64, true><<<grid, 64, 0, h->stream>>>(h->coord.p, h->km.p, (long long)h->nels, h->km_layout, d_dee.p, d_etype.p);
  h->launches++;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->stream));
  d_dee.release(); d_etype.release();
  h->have_km = true; h->have_precon = false;
  return 0;
}

int pf_form_kc_laplace(pf_handle h, double kx, double ky, double kz) {
  int rc = need_device(h); if (rc) return rc;
  h->one_km = false;
  h->transient = false; h->explicit_ = false; h->dynamic = false; h->kb.release();
  NEED(h->have_mesh && h->nodof == 1 && (h->nod == 8 || h->nod == 4), "needs pf_setup_mesh with nod = 8 or 4, nodof = 1");
  NEED(h->nod != 4 || h->km_layout == 0, "tetrahedra: reference storkm layout only");
  ElemTables T;
  if (fill_tables(h->nod, h->nip, 0, 0, kx, ky, kz, T)) return fail(h, 3, "unsupported nod/nip");
  if ((rc = upload_tables(h, T))) return rc;
  if ((rc = alloc_km(h))) return rc;
  const int grid = (int)std::min<int64_t>(h->nels, (int64_t)h->sm_count * 32);
  if (h->nod == 8) k_form_kc_laplace<8><<<grid, 64, 0, h->stream>>>(h->coord.p, h->km.p, (long long)h->nels, h->km_layout);
  else k_form_kc_laplace<4><<<grid, 32, 0, h->stream>>>(h->coord.p, h->km.p, (long long)h->nels, h->km_layout);
  h->launches++;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->stream));
  h->have_km = true; h->have_precon = false;
  return 0;
}

// ---- p124: transient heat conduction, implicit theta method ----
int pf_form_k_transient(pf_handle h, double kx, double ky, double kz, double rho, double cp, double theta, double dtim) {
  int rc = need_device(h); if (rc) return rc;
  h->one_km = false;
  NEED(h->have_mesh && h->nodof == 1 && h->nod == 8, "needs pf_setup_mesh with nod = 8, nodof = 1");
  NEED(h->km_layout == 0 && !h->matrix_free, "the transient matrices use the reference storkm layout");
  ElemTables T;
  if (fill_tables(h->nod, h->nip, 0, 0, kx, ky, kz, T)) return fail(h, 3, "unsupported nod/nip");
  T.trans[0] = rho; T.trans[1] = cp; T.trans[2] = theta; T.trans[3] = dtim;
  if ((rc = upload_tables(h, T))) return rc;
  if ((rc = alloc_km(h))) return rc;
  if (h->kb.n != h->km.n) CU(h->kb.alloc(h->km.n));
  const int grid = (int)std::min<int64_t>(h->nels, (int64_t)h->sm_count * 32);
  k_form_k_transient<<<grid, 64, 0, h->stream>>>(h->coord.p, h->km.p, h->kb.p, (long long)h->nels);
  h->launches++;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->stream));
  h->have_km = true; h->have_precon = false; h->transient = true; h->transient_first = false;
  return 0;
}

int pf_get_storkb(pf_handle h, int64_t iel0, int64_t n, double *out) {
  int rc = need_device(h); if (rc) return rc;
  NEED((h->transient || h->dynamic) && h->have_km && iel0 >= 0 && n >= 0 && iel0 + n <= h->nels,
       "needs pf_form_k_transient (storkb_pp) or pf_form_dynamic (store_km*c2 + store_mm*c3) and a local element range");
  const size_t per = (size_t)h->ntot * h->ntot;
  CU(cudaMemcpy(out, h->kb.p + (size_t)iel0 * per, (size_t)n * per * 8, cudaMemcpyDeviceToHost));
  return 0;
}

int pf_transient_start(pf_handle h, double val0, const double *val_f_pp) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->transient && h->have_precon, "needs pf_form_k_transient and pf_build_precon");
  NEED(h->nfixed == 0 || val_f_pp, "fixed freedoms were declared in pf_build_precon: val_f_pp is required");
  // x_pp = val0; x_pp(l) = val_f(k) on the fixed freedoms   (p124.f90:168-173)
  if (h->neq_pp > 0) {
    k_fill<<<grid_for(h, h->neq_pp, 256), 256, 0, h->stream>>>(h->x.p, val0, (long long)h->neq_pp);
    h->launches++;
  }
  if (h->nfixed > 0) {
    CU(h->val_f.alloc((size_t)h->nfixed));
    CU(cudaMemcpyAsync(h->val_f.p, val_f_pp, (size_t)h->nfixed * 8, cudaMemcpyHostToDevice, h->stream));
    k_fixed_transient<<<(h->nfixed + 255) / 256, 256, 0, h->stream>>>(h->fix_slot.p, h->store.p, h->val_f.p, h->x.p - 1, h->nfixed, 0);
    h->launches++;
  }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->stream));
  h->transient_first = true;
  return 0;
}

int pf_transient_step(pf_handle h, const double *loads_pp, double tol, int limit, int *iters, int *converged,
                      double *elapsed_ms) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->transient && h->have_precon, "needs pf_form_k_transient, pf_build_precon and pf_transient_start");
  EventPair ev;
  CU(ev.create());
  const cudaEvent_t e0 = ev.a, e1 = ev.b;
  CU(cudaEventRecord(e0, h->stream));
  // u = storka*x on the first step (p124.f90:174-178), storkb*xnew afterwards (:149-154); x lives in h->x
  CU(cudaMemcpyAsync(h->p_ext.p + 1, h->x.p, (size_t)h->neq_pp * 8, cudaMemcpyDeviceToDevice, h->stream));
  const int nfixed = h->nfixed;
  h->nfixed = 0;                                   // no u = p*store fix-up on this product
  h->mat_override = h->transient_first ? nullptr : h->kb.p;
  rc = apply_operator(h, nullptr);
  h->mat_override = nullptr;
  h->nfixed = nfixed;
  if (rc) return rc;
  if (!h->transient_first && nfixed > 0) {         // u_pp(l) = store_pp(i)*val_f(k)   (:155-160)
    k_fixed_transient<<<(nfixed + 255) / 256, 256, 0, h->stream>>>(h->fix_slot.p, h->store.p, h->val_f.p, h->u_ext.p, nfixed, 1);
    h->launches++;
  }
  // loads_pp = loads_pp + u_pp ; r_pp = loads_pp - r_pp, r_pp = +0.0 off the fixed freedoms (:161,:187-199)
  const double *l = nullptr;
  if (loads_pp) {
    CU(cudaMemcpyAsync(h->d.p, loads_pp, (size_t)h->neq_pp * 8, cudaMemcpyHostToDevice, h->stream));
    l = h->d.p;
  }
  if (h->neq_pp > 0) {
    k_transient_rhs<<<grid_for(h, h->neq_pp, 256), 256, 0, h->stream>>>(h->r.p, l, h->u_ext.p + 1, (long long)h->neq_pp);
    h->launches++;
  }
  if (nfixed > 0) {
    k_fixed_transient<<<(nfixed + 255) / 256, 256, 0, h->stream>>>(h->fix_slot.p, h->store.p, h->val_f.p, h->r.p - 1, nfixed, 2);
    h->launches++;
  }
  CU(cudaGetLastError());
  h->transient_first = false;
  // d = M^-1 r, p = d, x = 0 and the PCG loop (:199-218)
  if ((rc = pf_pcg_run(h, tol, limit, iters, converged, nullptr))) return rc;
  CU(cudaEventRecord(e1, h->stream));
  CU(cudaEventSynchronize(e1));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  if (elapsed_ms) *elapsed_ms = ms;
  return 0;
}

// ---- p125: explicit transient conduction (forward Euler with a lumped mass) ----
int pf_form_k_explicit(pf_handle h, double kx, double ky, double kz, double dtim) {
  int rc = need_device(h); if (rc) return rc;
  h->one_km = false;
  NEED(h->have_mesh && h->nodof == 1 && h->nod == 8, "needs pf_setup_mesh with nod = 8, nodof = 1");
  NEED(h->km_layout == 0 && !h->matrix_free, "store_pm_pp uses the reference storkm layout");
  h->transient = false; h->kb.release();
  ElemTables T;
  if (fill_tables(h->nod, h->nip, 0, 0, kx, ky, kz, T)) return fail(h, 3, "unsupported nod/nip");
  T.trans[3] = dtim;
  if ((rc = upload_tables(h, T))) return rc;
  if ((rc = alloc_km(h))) return rc;
  CU(h->diag_tmp.alloc((size_t)h->nels * h->ntot));
  const int grid = (int)std::min<int64_t>(h->nels, (int64_t)h->sm_count * 32);
  k_form_k_explicit<<<grid, 64, 0, h->stream>>>(h->coord.p, h->km.p, h->diag_tmp.p, (long long)h->nels);
  h->launches++;
  // globma_pp = 1/scatter(globma_tmp)   (p125.f90:78,82)
  h->nfixed = 0;
  {
    Scope sc(h, K_SCATTER);
    k_scatter<false><<<grid_for(h, h->nslots, 256, 16), 256, 0, h->stream>>>(h->csr_ptr.p, h->csr_pos.p, h->diag_tmp.p, h->diag_ext.p,
                                                                              (long long)h->nslots, h->ntot, nullptr);
    h->launches++;
  }
  if ((rc = halo_reverse(h, h->diag_ext.p, nullptr))) return rc;
  if (h->neq_pp > 0) {
    k_invert<<<grid_for(h, h->neq_pp, 256), 256, 0, h->stream>>>(h->diag_ext.p + 1, (long long)h->neq_pp);
    h->launches++;
  }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->stream));
  h->diag_tmp.release();
  h->have_km = true; h->have_precon = true; h->explicit_ = true; h->epoch++;
  return 0;
}

int pf_explicit_start(pf_handle h, double val0) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->explicit_, "needs pf_form_k_explicit");
  if (h->neq_pp > 0) {     // loads_pp = val0 (p125.f90:83); the field lives in the owned part of p_ext
    k_fill<<<grid_for(h, h->neq_pp, 256), 256, 0, h->stream>>>(h->p_ext.p + 1, val0, (long long)h->neq_pp);
    k_fill<<<grid_for(h, h->neq_pp, 256), 256, 0, h->stream>>>(h->x.p, val0, (long long)h->neq_pp);
    h->launches += 2;
  }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int pf_explicit_steps(pf_handle h, int nsteps, double *elapsed_ms) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->explicit_ && nsteps >= 0, "needs pf_form_k_explicit and nsteps >= 0");
  EventPair ev;
  CU(ev.create());
  const cudaEvent_t e0 = ev.a, e1 = ev.b;
  CU(cudaEventRecord(e0, h->stream));
  for (int j = 0; j < nsteps; ++j) {
    // newlo_pp = scatter(MATMUL(store_pm_pp, gather(loads_pp))); loads_pp = newlo_pp*globma_pp  (p125.f90:94-99)
    if ((rc = apply_operator(h, nullptr))) return rc;
    if (h->neq_pp > 0) {
      Scope sc(h, K_VECTOR);
      k_scale<<<grid_for(h, h->neq_pp, 256), 256, 0, h->stream>>>(h->p_ext.p + 1, h->u_ext.p + 1, h->diag_ext.p + 1, (long long)h->neq_pp);
      h->launches++;
    }
  }
  CU(cudaMemcpyAsync(h->x.p, h->p_ext.p + 1, (size_t)h->neq_pp * 8, cudaMemcpyDeviceToDevice, h->stream));
  CU(cudaEventRecord(e1, h->stream));
  CU(cudaEventSynchronize(e1));
  CU(cudaGetLastError());
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  collect_spans(h);
  if (elapsed_ms) *elapsed_ms = ms;
  return 0;
}

static int pcg_run_impl(pf_handle h, double tol, int limit, int *iters, int *converged, double *elapsed_ms, int keep_x);

// ---- p122: 3-D elasto-plasticity, Mohr-Coulomb, viscoplastic strain method ----
// pf_plastic_begin after pf_form_km_elastic(e, v) + pf_build_precon: zero stresses / totals (p122.f90:88-91), the
// critical time step (:92-93), angle constants from the host's libm.
int pf_plastic_begin(pf_handle h, double phi, double c, double psi, double e, double v, double *dt_out) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->have_precon && h->nodof == 3 && !h->matrix_free && h->km_layout == 0 && h->nip == 8 && (h->nod == 8 || h->nod == 20),
       "needs pf_form_km_elastic (hexahedra, nip = 8, reference storkm layout) and pf_build_precon");
  const double pi = std::acos(-1.0);
  const double snph = std::sin(phi * pi / 180.0);                                   // p122.f90:91
  const double dt = 4.0 * (1.0 + v) * (1.0 - 2.0 * v) / (e * (1.0 - 2.0 * v + snph * snph));
  const double phir = phi * 4.0 * std::atan(1.0) / 180.0, psir = psi * 4.0 * std::atan(1.0) / 180.0;   // mocouf / mocouq
  h->pl_par = PlasticParams{std::sin(phir), std::cos(phir), c, std::sin(psir), dt};
  const size_t ng = (size_t)h->nels * h->nip * 6, nq = (size_t)std::max<int64_t>(h->neq_pp, 1);
  CU(h->evpt.alloc(ng)); CU(h->tensor.alloc(ng));
  CU(h->bdylds.alloc(nq)); CU(h->oldis.alloc(nq)); CU(h->totd.alloc(nq)); CU(h->loads.alloc(nq));
  CU(cudaMemsetAsync(h->tensor.p, 0, ng * 8, h->stream)); CU(cudaMemsetAsync(h->evpt.p, 0, ng * 8, h->stream));
  CU(cudaMemsetAsync(h->oldis.p, 0, nq * 8, h->stream)); CU(cudaMemsetAsync(h->totd.p, 0, nq * 8, h->stream));
  CU(cudaMemsetAsync(h->x.p, 0, nq * 8, h->stream)); CU(cudaMemsetAsync(h->bdylds.p, 0, nq * 8, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->plastic = true;
  if (dt_out) *dt_out = dt;
  return 0;
}

// One pass of load_increments (p122.f90:115-205) on the device: the plastic iteration loop, each iteration a PCG
// solve restarted from the current x and a Gauss-point stress update; the host sees two scalars per plastic iteration.
int pf_plastic_increment(pf_handle h, double qinc, const double *ld0_pp, const double *valf_pp, int plasits,
                         double plastol, int cjits, double cjtol, int *plasiters_out, int *cjtot_out, double *elapsed_ms) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->plastic, "needs pf_plastic_begin");
  NEED(plasits >= 1 && cjits >= 1, "plasits and cjits must be >= 1");
  NEED(h->nfixed == 0 || valf_pp, "fixed freedoms were declared in pf_build_precon: valf_pp is required");
  if ((rc = ensure_tables(h))) return rc;
  const long long n = h->neq_pp;
  const int g = grid_for(h, std::max<long long>(n, 1), 256);
  EventPair ev;
  CU(ev.create());
  CU(cudaEventRecord(ev.a, h->stream));
  const double *ld0 = nullptr;
  if (ld0_pp) {
    CU(h->ld0.alloc((size_t)std::max<long long>(n, 1)));
    CU(cudaMemcpyAsync(h->ld0.p, ld0_pp, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    ld0 = h->ld0.p;
  }
  if (h->nfixed > 0) {
    CU(h->valf.alloc((size_t)h->nfixed));
    CU(cudaMemcpyAsync(h->valf.p, valf_pp, (size_t)h->nfixed * 8, cudaMemcpyHostToDevice, h->stream));
  }
  // plasiters = 0; bdylds_pp = zero; evpt_pp = zero; cjtot = 0   (p122.f90:116)
  CU(cudaMemsetAsync(h->bdylds.p, 0, (size_t)std::max<long long>(n, 1) * 8, h->stream));
  CU(cudaMemsetAsync(h->evpt.p, 0, h->evpt.bytes(), h->stream));
  int plasiters = 0, cjtot = 0;
  const int nfixed = h->nfixed;
  for (;;) {
    ++plasiters;
    // loads (p122.f90:119-138)
    if (plasiters == 1) {
      k_plastic_loads<<<g, 256, 0, h->stream>>>(h->loads.p, nullptr, h->bdylds.p, qinc, 0, n);
      if (nfixed > 0) k_plastic_fixed<<<(nfixed + 255) / 256, 256, 0, h->stream>>>(h->fix_slot.p, h->store.p, h->valf.p, h->loads.p - 1, qinc, nfixed, 0);
      if (ld0) k_plastic_loads<<<g, 256, 0, h->stream>>>(h->loads.p, ld0, h->bdylds.p, qinc, 1, n);
    } else {
      k_plastic_loads<<<g, 256, 0, h->stream>>>(h->loads.p, ld0, h->bdylds.p, qinc, 1, n);
      if (nfixed > 0) k_plastic_fixed<<<(nfixed + 255) / 256, 256, 0, h->stream>>>(h->fix_slot.p, h->store.p, h->valf.p, h->loads.p - 1, qinc, nfixed, 1);
    }
    h->launches += 1 + (nfixed > 0) + (plasiters == 1 && ld0);
    // r = loads - A*x   (:139-145); no fixed-freedom fix-up on this product
    CU(cudaMemcpyAsync(h->p_ext.p + 1, h->x.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, h->stream));
    h->nfixed = 0;
    rc = apply_operator(h, nullptr);
    h->nfixed = nfixed;
    if (rc) return rc;
    k_vsub<<<g, 256, 0, h->stream>>>(h->r.p, h->loads.p, h->u_ext.p + 1, n);
    h->launches++;
    // PCG from the current x; fixed rows: u = p*store on the first plastic iteration, 0 afterwards (:154-160)
    const int mode = plasiters == 1 ? 0 : 1;
    if (mode != h->fixed_mode) { h->fixed_mode = mode; h->epoch++; }
    int cjiters = 0, cjconv = 0;
    rc = pcg_run_impl(h, cjtol, cjits, &cjiters, &cjconv, nullptr, 1);
    if (rc) { h->fixed_mode = 0; h->epoch++; return rc; }
    cjtot += cjiters;
    // loads = xnew; checon_par(loads, plastol, plastic_converged, oldis)   (:171-174)
    k_checon<<<vec_grid(h, k_checon), kRedThreads, 0, h->stream>>>(h->x.p, h->oldis.p, n, h->part.p, h->state.p);
    h->launches++;
    std::vector<double> all((size_t)4 * h->nranks, 0.0);
    if (h->nranks > 1) {
      NC(g_nccl.AllGather(h->state.p->loc, h->gath.p, 4, ncclDouble, h->comm, h->stream));
      CU(cudaMemcpyAsync(all.data(), h->gath.p, all.size() * 8, cudaMemcpyDeviceToHost, h->stream));
    } else {
      CU(cudaMemcpyAsync(all.data(), h->state.p->loc, 4 * 8, cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    double maxloads = 0.0, maxdiff = 0.0;
    for (int r = 0; r < h->nranks; ++r) { maxloads = std::max(maxloads, all[(size_t)4 * r + 1]); maxdiff = std::max(maxdiff, all[(size_t)4 * r + 2]); }
    bool conv = (maxdiff / maxloads) <= plastol;
    if (plasiters == 1) conv = false;
    const int last = conv || plasiters == plasits;
    if (last) CU(cudaMemsetAsync(h->bdylds.p, 0, (size_t)std::max<long long>(n, 1) * 8, h->stream));
    // gather(loads) -> elements_4 -> scatter adds into bdylds   (:176-203)
    CU(cudaMemcpyAsync(h->p_ext.p + 1, h->x.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, h->stream));
    if ((rc = halo_forward(h, h->p_ext.p, nullptr))) return rc;
    const int ge = (int)std::min<int64_t>(h->nels, (int64_t)h->sm_count * 16);
    if (h->nod == 20) k_p122_elements<20><<<ge, 64, 0, h->stream>>>(h->coord.p, h->ggl.p, h->p_ext.p, h->evpt.p, h->tensor.p, h->utemp.p, (long long)h->nels, h->pl_par, last);
    else k_p122_elements<8><<<ge, 64, 0, h->stream>>>(h->coord.p, h->ggl.p, h->p_ext.p, h->evpt.p, h->tensor.p, h->utemp.p, (long long)h->nels, h->pl_par, last);
    h->launches++;
    CU(cudaGetLastError());
    if ((rc = launch_scatter(h, nullptr, false, h->u_ext.p))) return rc;
    if ((rc = halo_reverse(h, h->u_ext.p, nullptr))) return rc;
    k_vadd<<<g, 256, 0, h->stream>>>(h->bdylds.p, h->bdylds.p, h->u_ext.p + 1, n);
    h->launches++;
    if (last) break;
  }
  // totd = totd + loads   (:205)
  k_vadd<<<g, 256, 0, h->stream>>>(h->totd.p, h->totd.p, h->x.p, n);
  h->launches++;
  h->fixed_mode = 0; h->epoch++;
  CU(cudaEventRecord(ev.b, h->stream));
  CU(cudaEventSynchronize(ev.b));
  CU(cudaGetLastError());
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, ev.a, ev.b));
  if (plasiters_out) *plasiters_out = plasiters;
  if (cjtot_out) *cjtot_out = cjtot;
  if (elapsed_ms) *elapsed_ms = ms;
  return 0;
}

// totd_pp (neq_pp, may be NULL) and the stress tensor_pp(:,ig,iel) of one Gauss point (local 0-based; may be NULL)
int pf_plastic_get(pf_handle h, double *totd_pp, int64_t iel, int ig, double *tensor6) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->plastic, "needs pf_plastic_begin");
  if (totd_pp) CU(cudaMemcpy(totd_pp, h->totd.p, (size_t)h->neq_pp * 8, cudaMemcpyDeviceToHost));
  if (tensor6) {
    NEED(iel >= 0 && iel < h->nels && ig >= 0 && ig < h->nip, "element / Gauss point outside the local range");
    CU(cudaMemcpy(tensor6, h->tensor.p + ((size_t)iel * h->nip + ig) * 6, 48, cudaMemcpyDeviceToHost));
  }
  return 0;
}

// ---- p1210: forced vibration of an elastic-plastic (von Mises) solid, lumped mass, explicit integration ----
// (programs/5th_ed/p1210/p1210.f90).  No element matrices and no PCG: pf_setup_mesh, then
// pf_vm_explicit_begin (element tables, lumped mass :93-104, zero state :112), pf_vm_explicit_steps (:114-150).
int pf_vm_explicit_begin(pf_handle h, double e, double v, double sbary, double rho, double dtim, double pload, const double *fext_pp) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->have_mesh && h->nodof == 3 && h->nod == 20 && h->nip == 8, "needs pf_setup_mesh with 20-node hexahedra, nodof = 3, nip = 8");
  NEED(dtim > 0.0 && rho > 0.0, "dtim and rho must be positive");
  ElemTables T;
  if (fill_tables(h->nod, h->nip, e, v, 0, 0, 0, T)) return fail(h, 3, "unsupported nod/nip");
  if ((rc = upload_tables(h, T))) return rc;
  h->vm_par = VmParams{e, v, sbary};
  h->vm_dtim = dtim; h->vm_pload = pload;
  const size_t ng = (size_t)std::max<int64_t>(h->nels * h->nip * 6, 1), nq = (size_t)std::max<int64_t>(h->neq_pp, 1);
  CU(h->vm_eten.alloc(ng)); CU(h->vm_ten.alloc(ng));
  CU(h->vm_mm.alloc(nq)); CU(h->vm_fext.alloc(nq)); CU(h->vm_d1.alloc(nq)); CU(h->vm_d2.alloc(nq));
  CU(cudaMemsetAsync(h->vm_eten.p, 0, ng * 8, h->stream)); CU(cudaMemsetAsync(h->vm_ten.p, 0, ng * 8, h->stream));
  CU(cudaMemsetAsync(h->vm_d1.p, 0, nq * 8, h->stream)); CU(cudaMemsetAsync(h->vm_d2.p, 0, nq * 8, h->stream));
  CU(cudaMemsetAsync(h->vm_fext.p, 0, nq * 8, h->stream));
  CU(cudaMemsetAsync(h->p_ext.p, 0, h->p_ext.n * 8, h->stream));
  if (fext_pp && h->neq_pp > 0) CU(cudaMemcpyAsync(h->vm_fext.p, fext_pp, (size_t)h->neq_pp * 8, cudaMemcpyHostToDevice, h->stream));
  // diagonal mass matrix: mm_tmp -> scatter (with the reverse halo exchange) -> mm_pp
  if (h->nels > 0) {
    k_p1210_mass<<<grid_for(h, h->nels, 128), 128, 0, h->stream>>>(h->coord.p, h->utemp.p, (long long)h->nels, rho);
    h->launches++;
  }
  if ((rc = launch_scatter(h, nullptr, false, h->u_ext.p))) return rc;
  if ((rc = halo_reverse(h, h->u_ext.p, nullptr))) return rc;
  CU(cudaMemcpyAsync(h->vm_mm.p, h->u_ext.p + 1, (size_t)h->neq_pp * 8, cudaMemcpyDeviceToDevice, h->stream));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->stream));
  h->have_km = false; h->have_precon = false; h->plastic = false; h->dynamic = false; h->transient = false; h->explicit_ = false;
  h->vm_explicit = true;
  return 0;
}

int pf_vm_explicit_set_form(pf_handle h, int form) {
  int rc = need_device(h); if (rc) return rc;
  NEED(form == 0 || form == 1, "form must be 0 (elements_2 as written) or 1 (operator form, FP64 tensor cores)");
  h->vm_form = form;
  return 0;
}

int pf_vm_explicit_steps(pf_handle h, int nsteps, double *elapsed_ms) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->vm_explicit && nsteps >= 0, "needs pf_vm_explicit_begin and nsteps >= 0");
  if ((rc = ensure_tables(h))) return rc;
  EventPair ev;
  CU(ev.create());
  CU(cudaEventRecord(ev.a, h->stream));
  const long long n = h->neq_pp;
  const int vgrid = grid_for(h, std::max<int64_t>(n, 1), 256);
  const int egrid = (int)std::max<int64_t>(1, std::min<int64_t>(h->nels, (int64_t)h->sm_count * 16));
  for (int j = 0; j < nsteps; ++j) {
    if (n > 0) {
      Scope sc(h, K_VECTOR);
      k_p1210_predict<<<vgrid, 256, 0, h->stream>>>(h->p_ext.p + 1, h->vm_d1.p, h->vm_d2.p, h->vm_dtim, n);
      h->launches++;
    }
    if ((rc = halo_forward(h, h->p_ext.p, nullptr))) return rc;
    if (h->nels > 0) {
      Scope sc(h, K_MATVEC);
      if (h->vm_form == 1) {
        // operator form on the FP64 tensor cores: k_apply_mf4's pipeline with p1210's Gauss-point update (8 consumer warps
        // of 224 registers; coordinates -> Jacobian every step, as the reference recomputes it)
        using Cfg = Mf4Cfg<20, 8>;
        auto kern = k_apply_mf4<20, true, 0, 8, 1>;
        if ((rc = ensure_smem(h, kern, Cfg::kSmem))) return rc;
        const int64_t npass = (h->nels + Cfg::EPP - 1) / Cfg::EPP;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(h->sm_count, (npass + Cfg::NCONS - 1) / Cfg::NCONS));
        kern<<<grid, Cfg::kThreads, Cfg::kSmem, h->stream>>>(h->coord.p, h->ggl.p, h->p_ext.p, h->utemp.p, (long long)h->nels, nullptr,
                                                             nullptr, nullptr, h->vm_eten.p, h->vm_ten.p, h->vm_par);
      } else {
        k_p1210_elements<20><<<egrid, 64, 0, h->stream>>>(h->coord.p, h->ggl.p, h->p_ext.p, h->vm_eten.p, h->vm_ten.p, h->utemp.p,
                                                          (long long)h->nels, h->vm_par);
      }
      h->launches++;
    }
    if ((rc = launch_scatter(h, nullptr, false, h->u_ext.p))) return rc;
    if ((rc = halo_reverse(h, h->u_ext.p, nullptr))) return rc;
    if (n > 0) {
      Scope sc(h, K_VECTOR);
      k_p1210_update<<<vgrid, 256, 0, h->stream>>>(h->u_ext.p + 1, h->vm_fext.p, h->vm_mm.p, h->vm_d1.p, h->vm_d2.p, h->vm_pload,
                                                   h->vm_dtim, n);
      h->launches++;
    }
  }
  CU(cudaEventRecord(ev.b, h->stream));
  CU(cudaEventSynchronize(ev.b));
  CU(cudaGetLastError());
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, ev.a, ev.b));
  collect_spans(h);
  if (elapsed_ms) *elapsed_ms = ms;
  return 0;
}

int pf_vm_explicit_get(pf_handle h, double *x1_pp, double *d1x1_pp, double *d2x1_pp, double *mm_pp) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->vm_explicit, "needs pf_vm_explicit_begin");
  const size_t nb = (size_t)h->neq_pp * 8;
  if (x1_pp) CU(cudaMemcpy(x1_pp, h->p_ext.p + 1, nb, cudaMemcpyDeviceToHost));
  if (d1x1_pp) CU(cudaMemcpy(d1x1_pp, h->vm_d1.p, nb, cudaMemcpyDeviceToHost));
  if (d2x1_pp) CU(cudaMemcpy(d2x1_pp, h->vm_d2.p, nb, cudaMemcpyDeviceToHost));
  if (mm_pp) CU(cudaMemcpy(mm_pp, h->vm_mm.p, nb, cudaMemcpyDeviceToHost));
  return 0;
}

// ---- p129: forced vibration of an elastic solid, implicit theta method, consistent mass ----
int pf_form_dynamic(pf_handle h, double e, double v, double rho, double alpha1, double beta1, double theta, double dtim) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->have_mesh && h->nodof == 3 && (h->nod == 20 || h->nod == 8) && !h->matrix_free && h->km_layout == 0,
       "needs pf_setup_mesh with hexahedra, nodof = 3, the stored path and the reference layout");
  NEED(theta > 0.0 && dtim > 0.0, "theta and dtim must be positive");
  // store_km_pp (p129.f90:85-89) through the stiffness kernels, store_mm_pp (:90-93) through k_form_mass
  if ((rc = pf_form_km_elastic(h, e, v))) return rc;
  const size_t n = h->km.n;
  DevBuf<double> mm;
  CU(mm.alloc(n)); CU(h->kb.alloc(n)); CU(h->kc.alloc(n));
  const int grid = (int)std::min<int64_t>(h->nels, (int64_t)h->sm_count * 16);
  if (h->nod == 20) k_form_mass<20><<<grid, 128, 0, h->stream>>>(h->coord.p, mm.p, (long long)h->nels, rho);
  else k_form_mass<8><<<grid, 128, 0, h->stream>>>(h->coord.p, mm.p, (long long)h->nels, rho);
  // c1..c4 (p129.f90:80-82)
  const double c1 = (1.0 - theta) * dtim, c2 = beta1 - c1, c3 = alpha1 + 1.0 / (theta * dtim), c4 = beta1 + theta * dtim;
  const int g = grid_for(h, (int64_t)n, 256);
  k_lincomb<<<g, 256, 0, h->stream>>>(h->kb.p, h->km.p, c2, mm.p, c3, (long long)n);      // store_km*c2 + store_mm*c3   (:114)
  k_divide<<<g, 256, 0, h->stream>>>(h->kc.p, mm.p, theta, (long long)n);                  // store_mm/theta           (:120)
  k_lincomb<<<g, 256, 0, h->stream>>>(h->km.p, mm.p, c3, h->km.p, c4, (long long)n);      // store_mm*c3 + store_km*c4   (:128)
  h->launches += 4;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->stream));
  h->dynamic = true; h->dyn_theta = theta; h->dyn_dtim = dtim;
  h->have_km = true; h->have_precon = false;
  return 0;
}

int pf_dynamic_start(pf_handle h, const double *fext_pp) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->dynamic && h->have_precon && fext_pp, "needs pf_form_dynamic, pf_build_precon and fext_pp");
  const size_t nq = (size_t)std::max<int64_t>(h->neq_pp, 1);
  CU(h->fext.alloc(nq)); CU(h->x0.alloc(nq)); CU(h->d1x0.alloc(nq)); CU(h->d2x0.alloc(nq));
  CU(cudaMemcpyAsync(h->fext.p, fext_pp, (size_t)h->neq_pp * 8, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemsetAsync(h->x0.p, 0, nq * 8, h->stream)); CU(cudaMemsetAsync(h->d1x0.p, 0, nq * 8, h->stream));
  CU(cudaMemsetAsync(h->d2x0.p, 0, nq * 8, h->stream));                                   // p129.f90:112
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

// One pass of `timesteps` (p129.f90:113-149); load_factor = theta*dtim*cos(omega*t) + c1*cos(omega*(t-dtim)) (:124-125)
int pf_dynamic_step(pf_handle h, double load_factor, double tol, int limit, int *iters, int *converged, double *elapsed_ms) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->dynamic && h->have_precon && h->fext.p, "needs pf_form_dynamic, pf_build_precon and pf_dynamic_start");
  const long long n = h->neq_pp;
  const int g = grid_for(h, std::max<long long>(n, 1), 256);
  EventPair ev;
  CU(ev.create());
  CU(cudaEventRecord(ev.a, h->stream));
  const int nfixed = h->nfixed;
  h->nfixed = 0;
  // u = (store_km*c2 + store_mm*c3) x0 ; vu = (store_mm/theta) d1x0   (:114-123)
  CU(cudaMemcpyAsync(h->p_ext.p + 1, h->x0.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, h->stream));
  h->mat_override = h->kb.p;
  rc = apply_operator(h, nullptr);
  if (!rc) {
    cudaMemcpyAsync(h->d.p, h->u_ext.p + 1, (size_t)n * 8, cudaMemcpyDeviceToDevice, h->stream);
    cudaMemcpyAsync(h->p_ext.p + 1, h->d1x0.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, h->stream);
    h->mat_override = h->kc.p;
    rc = apply_operator(h, nullptr);
  }
  h->mat_override = nullptr;
  h->nfixed = nfixed;
  if (rc) return rc;
  k_dyn_rhs<<<g, 256, 0, h->stream>>>(h->r.p, h->d.p, h->u_ext.p + 1, h->fext.p, load_factor, n);
  h->launches++;
  CU(cudaGetLastError());
  // d = M^-1 loads, p = d, x = 0 and the PCG loop on store_mm*c3 + store_km*c4   (:127-144)
  if ((rc = pcg_run_impl(h, tol, limit, iters, converged, nullptr, 0))) return rc;
  k_dyn_update<<<g, 256, 0, h->stream>>>(h->x.p, h->x0.p, h->d1x0.p, h->d2x0.p, h->dyn_theta, h->dyn_dtim, n);
  h->launches++;
  CU(cudaEventRecord(ev.b, h->stream));
  CU(cudaEventSynchronize(ev.b));
  CU(cudaGetLastError());
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, ev.a, ev.b));
  if (elapsed_ms) *elapsed_ms = ms;
  return 0;
}

int pf_dynamic_get(pf_handle h, double *x_pp, double *d1x_pp, double *d2x_pp) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->dynamic && h->x0.p, "needs pf_dynamic_start");
  if (x_pp) CU(cudaMemcpy(x_pp, h->x0.p, (size_t)h->neq_pp * 8, cudaMemcpyDeviceToHost));
  if (d1x_pp) CU(cudaMemcpy(d1x_pp, h->d1x0.p, (size_t)h->neq_pp * 8, cudaMemcpyDeviceToHost));
  if (d2x_pp) CU(cudaMemcpy(d2x_pp, h->d2x0.p, (size_t)h->neq_pp * 8, cudaMemcpyDeviceToHost));
  return 0;
}

int pf_set_storkm(pf_handle h, const double *storkm_pp) {
  int rc = need_device(h); if (rc) return rc;
  h->one_km = false;
  NEED(h->have_mesh && storkm_pp, "needs pf_setup_mesh");
  NEED(!h->matrix_free, "matrix-free variant: storkm is not stored");
  h->transient = false; h->explicit_ = false; h->dynamic = false; h->kb.release();
  if ((rc = alloc_km(h))) return rc;
  if (h->km_layout == 0) {
    CU(cudaMemcpy(h->km.p, storkm_pp, h->km.bytes(), cudaMemcpyHostToDevice));
  } else {
    // symmetric layout: the lower triangle of every matrix is kept (chunks through a staging buffer)
    const size_t per = (size_t)h->ntot * h->ntot, P = km_per_element(h);
    const int64_t chunk = std::max<int64_t>(1, (int64_t)((256u << 20) / (per * 8)));
    DevBuf<double> stage;
    CU(stage.alloc((size_t)std::min<int64_t>(chunk, h->nels) * per));
    for (int64_t e = 0; e < h->nels; e += chunk) {
      const int64_t n = std::min<int64_t>(chunk, h->nels - e);
      CU(cudaMemcpyAsync(stage.p, storkm_pp + (size_t)e * per, (size_t)n * per * 8, cudaMemcpyHostToDevice, h->stream));
      k_pack_lower<<<grid_for(h, n * (int64_t)per, 256), 256, 0, h->stream>>>(stage.p, h->km.p + (size_t)e * P, (long long)n, h->ntot);
      h->launches++;
      CU(cudaStreamSynchronize(h->stream));
    }
    CU(cudaGetLastError());
    stage.release();
  }
  h->have_km = true; h->have_precon = false;
  return 0;
}

int pf_get_storkm(pf_handle h, int64_t iel0, int64_t n, double *out) {
  int rc = need_device(h); if (rc) return rc;
  NEED(!h->matrix_free, "matrix-free variant: storkm is not stored");
  NEED(h->have_km && iel0 >= 0 && n >= 0 && iel0 + n <= h->nels, "range outside the local elements");
  const size_t per = (size_t)h->ntot * h->ntot;
  if (h->km_layout == 0) {
    CU(cudaMemcpy(out, h->km.p + (size_t)iel0 * per, (size_t)n * per * 8, cudaMemcpyDeviceToHost));
    return 0;
  }
  // symmetric layout: K(i,j) = L(max,min) -- the matrix the product uses
  const size_t P = km_per_element(h);
  const int64_t chunk = std::max<int64_t>(1, (int64_t)((256u << 20) / (per * 8)));
  DevBuf<double> stage;
  CU(stage.alloc((size_t)std::max<int64_t>(1, std::min<int64_t>(chunk, n)) * per));
  for (int64_t e = 0; e < n; e += chunk) {
    const int64_t m = std::min<int64_t>(chunk, n - e);
    k_unpack_lower<<<grid_for(h, m * (int64_t)per, 256), 256, 0, h->stream>>>(h->km.p + (size_t)(iel0 + e) * P, stage.p, (long long)m, h->ntot);
    h->launches++;
    CU(cudaMemcpyAsync(out + (size_t)e * per, stage.p, (size_t)m * per * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  CU(cudaGetLastError());
  stage.release();
  return 0;
}

int pf_set_storkm_layout(pf_handle h, int layout) {
  if (!h) return 1;
  NEED(layout == 0 || layout == 1, "layout must be 0 (reference storkm_pp) or 1 (packed lower triangles)");
  if (layout != h->km_layout) { h->km.release(); h->have_km = false; h->have_precon = false; }
  h->km_layout = layout;
  return 0;
}

int pf_set_matrix_free(pf_handle h, int on) {
  if (!h) return 1;
  NEED(!on || !h->have_mesh || h->nodof == 3, "the matrix-free variant exists for the elastic elements (p121) only");
  NEED(on >= 0 && on <= 2, "mode must be 0 (stored), 1 (recompute from coordinates) or 2 (stored geometric factors)");
  if ((on != 0) != h->matrix_free || on != h->mf_mode) { h->have_km = false; h->have_precon = false; }
  h->matrix_free = on != 0;
  h->mf_mode = on;
  return 0;
}

int pf_build_precon(pf_handle h, int64_t nfixed_pp, const int32_t *no_f_pp, double penalty) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->have_km, "needs element matrices (pf_form_km_elastic / pf_form_kc_laplace / pf_set_storkm)");
  NEED(nfixed_pp >= 0 && nfixed_pp < (1 << 30), "bad nfixed_pp");
  // diag_precon_tmp(i,iel) = storkm(i,i,iel); scatter  (p121.f90:65-69)
  if (h->matrix_free) {
    NEED(h->diag_tmp.n == (size_t)h->nels * h->ntot, "matrix-free: call pf_form_km_elastic first");
    Scope sc(h, K_SCATTER);
    k_scatter<false><<<grid_for(h, h->nslots, 256, 16), 256, 0, h->stream>>>(h->csr_ptr.p, h->csr_pos.p, h->diag_tmp.p, h->diag_ext.p,
                                                                              (long long)h->nslots, h->ntot, nullptr);
    h->launches++;
  } else if ((rc = launch_scatter(h, nullptr, true, h->diag_ext.p))) return rc;
  if ((rc = halo_reverse(h, h->diag_ext.p, nullptr))) return rc;
  // the rest is local to this rank; several ranks agree on the outcome before any of them goes on to a solve
  // (a rank that returned alone would leave the others waiting in the first exchange of the PCG loop)
  auto local = [&]() -> int {
    h->nfixed = (int)nfixed_pp;
    if (h->nfixed > 0) {
      NEED(no_f_pp, "nfixed_pp > 0 needs no_f_pp");
      std::vector<int> slots((size_t)h->nfixed);
      for (int i = 0; i < h->nfixed; ++i) {
        const int64_t g = no_f_pp[i];
        NEED(g >= h->ieq_start && g < h->ieq_start + h->neq_pp, "fixed equation not owned by this rank");
        slots[(size_t)i] = (int)(g - h->ieq_start + 1);
      }
      CU(h->fix_slot.alloc(slots.size())); CU(h->store.alloc(slots.size()));
      CU(cudaMemcpy(h->fix_slot.p, slots.data(), slots.size() * 4, cudaMemcpyHostToDevice));
      k_fixed_penalty<<<(h->nfixed + 255) / 256, 256, 0, h->stream>>>(h->fix_slot.p, h->diag_ext.p, h->store.p, penalty, h->nfixed);
      h->launches++;
    }
    if (h->neq_pp > 0) {
      k_invert<<<grid_for(h, h->neq_pp, 256), 256, 0, h->stream>>>(h->diag_ext.p + 1, (long long)h->neq_pp);
      h->launches++;
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    return 0;
  };
  if ((rc = agree(h, local()))) { h->nfixed = 0; return rc; }
  h->diag_tmp.release();
  h->have_precon = true;
  h->fixed_mode = 0;
  h->epoch++;
  return 0;
}

int pf_get_diag_precon(pf_handle h, double *out) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->have_precon, "needs pf_build_precon");
  CU(cudaMemcpy(out, h->diag_ext.p + 1, (size_t)h->neq_pp * 8, cudaMemcpyDeviceToHost));
  return 0;
}

int pf_get_store(pf_handle h, double *store_pp) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->have_precon, "needs pf_build_precon");
  if (h->nfixed > 0) CU(cudaMemcpy(store_pp, h->store.p, (size_t)h->nfixed * 8, cudaMemcpyDeviceToHost));
  return 0;
}

int pf_pcg_load_rhs(pf_handle h, const double *r_pp) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->have_mesh, "needs pf_setup_mesh");
  CU(cudaMemcpyAsync(h->r.p, r_pp, (size_t)h->neq_pp * 8, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

static int pcg_run_impl(pf_handle h, double tol, int limit, int *iters, int *converged, double *elapsed_ms, int keep_x);

int pf_pcg_run(pf_handle h, double tol, int limit, int *iters, int *converged, double *elapsed_ms) {
  return pcg_run_impl(h, tol, limit, iters, converged, elapsed_ms, 0);
}

// keep_x: the solve starts from the x of the previous one and h->r already holds loads - A*x (p122.f90:139-146)
static int pcg_run_impl(pf_handle h, double tol, int limit, int *iters, int *converged, double *elapsed_ms, int keep_x) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->have_precon, "needs pf_build_precon");
  if (h->matrix_free && (rc = ensure_tables(h))) return rc;
  NEED(limit >= 1, "limit must be >= 1");
  if (h->ratio_cap < limit) { CU(h->ratio_hist.alloc((size_t)limit)); h->ratio_cap = limit; h->epoch++; }
  State init; memset(&init, 0, sizeof init);
  init.tol = tol; init.limit = limit;
  CU(cudaMemcpyAsync(h->state.p, &init, sizeof init, cudaMemcpyHostToDevice, h->stream));
  // d = M^-1 r, p = d, x = 0 (p121.f90:87; p123.f90:132)
  k_pcg_init<<<vec_grid(h, k_pcg_init), kRedThreads, 0, h->stream>>>(h->diag_ext.p + 1, h->r.p, h->d.p, h->p_ext.p + 1, h->x.p,
                                                          (long long)h->neq_pp, h->part.p, h->state.p, h->nranks == 1,
                                                          (h->peer_ok && h->use_peer) ? h->ptab.p : nullptr, keep_x);
  h->launches++;
  if (!(h->peer_ok && h->use_peer) && (rc = combine_scalars(h, 0))) return rc;
  const bool peer = h->peer_ok && h->use_peer;
  if (peer && fused_peer()) {
    // the forward exchange of the first iteration; every later one rides on k_pupdate
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((h->nput + 255) / 256, 2 * h->sm_count));
    k_halo_put_peer<<<grid, 256, 0, h->stream>>>(h->ptab.p, h->put_slot.p, h->p_ext.p, (long long)h->nput, h->state.p);
    h->launches++;
    CU(cudaGetLastError());
  }
  EventPair ev;
  CU(ev.create());
  const cudaEvent_t e0 = ev.a, e1 = ev.b;
  CU(cudaEventRecord(e0, h->stream));  // timest(3), p121.f90:89
  // The iteration (5 launches with constant arguments; sequence numbers and scalars live in device memory) is
  // captured once per problem as a CUDA graph and replayed -- on one rank and, in the peer transport, on N ranks
  // (no NCCL call inside).  PF_GRAPH=0 disables it; so does per-launch profiling.
  static const bool graph_allowed = !(getenv("PF_GRAPH") && !strcmp(getenv("PF_GRAPH"), "0"));
  bool use_graph = graph_allowed && (h->nranks == 1 || (peer && fused_peer())) && !h->profile;
  if (use_graph && (!h->graph_exec || h->graph_epoch != h->epoch)) {
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
    cudaGraph_t g = nullptr;
    const int64_t l0 = h->launches;
    CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
    rc = one_iteration(h);
    const cudaError_t ce = cudaStreamEndCapture(h->stream, &g);
    h->graph_launches = h->launches - l0;
    h->launches = l0;
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    if (ce != cudaSuccess || !g || cudaGraphInstantiate(&h->graph_exec, g, 0) != cudaSuccess) {
      cudaGetLastError();
      h->graph_exec = nullptr;
      use_graph = false;                             // plain stream launches
    }
    if (g) cudaGraphDestroy(g);
    h->graph_epoch = h->epoch;
  }
  use_graph = use_graph && h->graph_exec;
  // The host runs ahead of the device: batches of 8 iterations, and the state snapshot of batch b is only awaited
  // after batch b+1 has been queued (two pinned snapshots) -- the device never idles while the host looks at the
  // stopping flag.  Every kernel returns at once when `done` is set, so a batch queued past convergence costs
  // microseconds.  Never more than `limit` iterations are queued.
  const int batch = 8;
  int queued = 0, slot = 0;
  auto enqueue = [&](int which) -> int {
    const int nb = std::min(batch, limit - queued);
    for (int k = 0; k < nb; ++k) {
      if (use_graph) { CU(cudaGraphLaunch(h->graph_exec, h->stream)); h->launches += h->graph_launches; }
      else if (int r2 = one_iteration(h)) return r2;
    }
    queued += nb;
    CU(cudaMemcpyAsync(&h->snap_pinned[which], h->state.p, sizeof(State), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaEventRecord(h->snap_ev[which], h->stream));
    return 0;
  };
  if ((rc = enqueue(slot))) return rc;
  State snap;
  for (;;) {
    const bool more = queued < limit;
    if (more && (rc = enqueue(slot ^ 1))) return rc;
    CU(cudaEventSynchronize(h->snap_ev[slot]));
    snap = h->snap_pinned[slot];
    if (snap.done || !more) {
      if (more) {                                    // the batch queued ahead drains as no-ops
        CU(cudaEventSynchronize(h->snap_ev[slot ^ 1]));
        snap = h->snap_pinned[slot ^ 1];
      }
      break;
    }
    slot ^= 1;
  }
  CU(cudaEventRecord(e1, h->stream));
  CU(cudaEventSynchronize(e1));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  collect_spans(h);
  h->last_iters = snap.iters;
  h->last_ms = ms;
  if (iters) *iters = snap.iters;
  if (converged) *converged = snap.converged;
  if (elapsed_ms) *elapsed_ms = ms;
  if (snap.fault)
    return fail(h, 14, "pf_pcg_run: a peer rank did not arrive at a halo / reduction flag within %d s; the solve was abandoned",
                (int)(kSpinTimeoutNs / 1000000000ull));
  return 0;
}

int pf_pcg_get_x(pf_handle h, double *xnew_pp) {
  int rc = need_device(h); if (rc) return rc;
  CU(cudaMemcpyAsync(xnew_pp, h->x.p, (size_t)h->neq_pp * 8, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int pf_pcg_solve(pf_handle h, const double *r_pp, double tol, int limit, double *xnew_pp, int *iters,
                 int *converged) {
  int rc;
  if ((rc = pf_pcg_load_rhs(h, r_pp))) return rc;
  if ((rc = pf_pcg_run(h, tol, limit, iters, converged, nullptr))) return rc;
  return pf_pcg_get_x(h, xnew_pp);
}

// PCG_KM (maths.f90:1152-1323): every element shares km(ntot,ntot); the caller supplies the inverted diagonal
// preconditioner of its own equations, as the Fortran routine's argument list does.
int pf_pcg_km(pf_handle h, const double *km, const double *diag_precon_pp, const double *r_pp, double tol, int limit,
              double *xnew_pp, int *iters, int *converged) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->have_mesh && km && diag_precon_pp && r_pp && xnew_pp, "needs pf_setup_mesh and km, diag_precon_pp, r_pp, xnew_pp");
  NEED(!h->matrix_free && h->km_layout == 0, "pcg_km takes the stored path in the reference layout");
  const size_t nk = (size_t)h->ntot * h->ntot;
  CU(h->km1.alloc(nk));
  CU(cudaMemcpyAsync(h->km1.p, km, nk * 8, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->diag_ext.p + 1, diag_precon_pp, (size_t)h->neq_pp * 8, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->km.release();
  h->one_km = true; h->have_km = true; h->have_precon = true; h->nfixed = 0; h->fixed_mode = 0;
  h->transient = false; h->explicit_ = false; h->plastic = false;
  h->epoch++;
  return pf_pcg_solve(h, r_pp, tol, limit, xnew_pp, iters, converged);
}

int pf_get_ratio_history(pf_handle h, double *out, int maxn, int *n) {
  int rc = need_device(h); if (rc) return rc;
  const int m = std::min(maxn, h->last_iters);
  if (m > 0) CU(cudaMemcpy(out, h->ratio_hist.p, (size_t)m * 8, cudaMemcpyDeviceToHost));
  if (n) *n = m;
  return 0;
}

// ---- fine-grained entry points ----
static int upload_owned(pf_handle h, double *ext, const double *host) {
  CU(cudaMemcpyAsync(ext + 1, host, (size_t)h->neq_pp * 8, cudaMemcpyHostToDevice, h->stream));
  return 0;
}

int pf_gather(pf_handle h, const double *p_pp, double *pmul_pp) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->have_mesh, "needs pf_setup_mesh");
  if ((rc = upload_owned(h, h->p_ext.p, p_pp))) return rc;
  if ((rc = halo_forward(h, h->p_ext.p, nullptr))) return rc;
  const int64_t n = h->nels * h->ntot;
  k_gather<<<grid_for(h, n, 256), 256, 0, h->stream>>>(h->ggl.p, h->p_ext.p, h->utemp.p, (long long)n);
  h->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(pmul_pp, h->utemp.p, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int pf_matvec(pf_handle h, const double *pmul_pp, double *utemp_pp) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->have_km, "needs element matrices");
  if (h->matrix_free && (rc = ensure_tables(h))) return rc;
  const size_t n = (size_t)h->nels * h->ntot;
  DevBuf<double> pm;
  CU(pm.alloc(n));
  CU(cudaMemcpyAsync(pm.p, pmul_pp, n * 8, cudaMemcpyHostToDevice, h->stream));
  rc = launch_matvec<false>(h, pm.p, nullptr);
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(utemp_pp, h->utemp.p, n * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) rc = fail(h, 10, "pf_matvec: %s", cudaGetErrorString(e));
  }
  cudaStreamSynchronize(h->stream);
  pm.release();
  return rc;
}

int pf_scatter(pf_handle h, const double *utemp_pp, double *u_pp) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->have_mesh, "needs pf_setup_mesh");
  const size_t n = (size_t)h->nels * h->ntot;
  CU(cudaMemcpyAsync(h->utemp.p, utemp_pp, n * 8, cudaMemcpyHostToDevice, h->stream));
  if ((rc = launch_scatter(h, nullptr, false, h->u_ext.p))) return rc;
  if ((rc = halo_reverse(h, h->u_ext.p, nullptr))) return rc;
  CU(cudaMemcpyAsync(u_pp, h->u_ext.p + 1, (size_t)h->neq_pp * 8, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int pf_apply(pf_handle h, const double *p_pp, double *u_pp) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->have_km, "needs element matrices");
  if (h->matrix_free && (rc = ensure_tables(h))) return rc;
  if ((rc = upload_owned(h, h->p_ext.p, p_pp))) return rc;
  const int nfixed = h->nfixed;
  if (!h->have_precon) h->nfixed = 0;
  rc = apply_operator(h, nullptr);
  h->nfixed = nfixed;
  if (rc) return rc;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(u_pp, h->u_ext.p + 1, (size_t)h->neq_pp * 8, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int pf_dot(pf_handle h, const double *a_pp, const double *b_pp, double *result) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->have_mesh, "needs pf_setup_mesh");
  // p_ext / u_ext double as staging for the two operands
  if ((rc = upload_owned(h, h->p_ext.p, a_pp))) return rc;
  if ((rc = upload_owned(h, h->u_ext.p, b_pp))) return rc;
  k_dot<false><<<vec_grid(h, k_dot<false>), kRedThreads, 0, h->stream>>>(h->p_ext.p + 1, h->u_ext.p + 1, (long long)h->neq_pp, h->part.p, h->state.p,
                                                                   h->nranks == 1, -1, nullptr, AccTables{});
  h->launches++;
  CU(cudaGetLastError());
  std::vector<double> all((size_t)4 * h->nranks, 0.0);
  if (h->nranks > 1) {
    NC(g_nccl.AllGather(h->state.p->loc, h->gath.p, 4, ncclDouble, h->comm, h->stream));
    CU(cudaMemcpyAsync(all.data(), h->gath.p, all.size() * 8, cudaMemcpyDeviceToHost, h->stream));
  } else {
    CU(cudaMemcpyAsync(all.data(), h->state.p->loc, 4 * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  CU(cudaStreamSynchronize(h->stream));
  double s = all[0];
  for (int r = 1; r < h->nranks; ++r) s = s + all[(size_t)4 * r];  // ranks ascending
  *result = s;
  return 0;
}

int pf_sum(pf_handle h, const double *a_pp, double *result) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->have_mesh, "needs pf_setup_mesh");
  // sum_p (maths.f90:271-315) through the dot-product tree: a(i)*1.0 == a(i)
  std::vector<double> ones((size_t)std::max<int64_t>(h->neq_pp, 1), 1.0);
  return pf_dot(h, a_pp, ones.data(), result);
}

int pf_norm(pf_handle h, const double *a_pp, double *result) {
  double s = 0.0;
  int rc = pf_dot(h, a_pp, a_pp, &s);
  if (rc) return rc;
  *result = std::sqrt(s);  // norm_p, maths.f90:258-261
  return 0;
}

static int stress_at(pf_handle h, int64_t iel, const double *pt, double e, double v, double *sigma6) {
  int rc = need_device(h); if (rc) return rc;
  NEED(h->have_mesh && h->nodof == 3 && iel >= 0 && iel < h->nels, "needs an elastic mesh and a local element index");
  // gather xnew into eld (p121.f90:114) for this one element, then one point at the centroid
  const int ntot = h->ntot, nod = h->nod;
  std::vector<int> g((size_t)ntot);
  std::vector<double> co((size_t)nod * 3), eld((size_t)ntot, 0.0);
  CU(cudaMemcpy(g.data(), h->ggl.p + iel * ntot, (size_t)ntot * 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(co.data(), h->coord.p + iel * nod * 3, co.size() * 8, cudaMemcpyDeviceToHost));
  if (h->nranks > 1) {
    // x lives in h->x (owned only): stage it through p_ext so halo values arrive
    CU(cudaMemcpyAsync(h->p_ext.p + 1, h->x.p, (size_t)h->neq_pp * 8, cudaMemcpyDeviceToDevice, h->stream));
    if ((rc = halo_forward(h, h->p_ext.p, nullptr))) return rc;
    CU(cudaStreamSynchronize(h->stream));
  }
  for (int k = 0; k < ntot; ++k) {
    if (g[(size_t)k] == 0) continue;
    const double *src = (h->nranks > 1) ? h->p_ext.p + g[(size_t)k] : h->x.p + (g[(size_t)k] - 1);
    if (h->nranks == 1 && g[(size_t)k] > h->neq_pp) return fail(h, 6, "slot outside owned range");
    CU(cudaMemcpy(&eld[(size_t)k], src, 8, cudaMemcpyDeviceToHost));
  }
  ElemTables T;
  if (fill_tables(nod, 1, e, v, 0, 0, 0, T)) return fail(h, 3, "unsupported nod");
  if (pt && shape_der_at(nod, pt[0], pt[1], pt[2], T.der, nullptr)) return fail(h, 3, "unsupported nod");
  // jac, inverse, deriv with the same operation order as the kernels
  double jac[9], inv[9], deriv[60];
  for (int b = 0; b < 3; ++b)
    for (int a = 0; a < 3; ++a) {
      double s = 0.0;
      for (int m = 0; m < nod; ++m) s = s + T.der[a * 20 + m] * co[(size_t)b * nod + m];
      jac[b * 3 + a] = s;
    }
#define A(r, c) jac[(c - 1) * 3 + (r - 1)]
  double det = A(1, 1) * (A(2, 2) * A(3, 3) - A(3, 2) * A(2, 3));
  det = det - A(1, 2) * (A(2, 1) * A(3, 3) - A(3, 1) * A(2, 3));
  det = det + A(1, 3) * (A(2, 1) * A(3, 2) - A(3, 1) * A(2, 2));
  inv[0] = (A(2, 2) * A(3, 3) - A(3, 2) * A(2, 3)) / det;
  inv[1] = (-(A(2, 1) * A(3, 3)) + A(3, 1) * A(2, 3)) / det;
  inv[2] = (A(2, 1) * A(3, 2) - A(3, 1) * A(2, 2)) / det;
  inv[3] = (-(A(1, 2) * A(3, 3)) + A(3, 2) * A(1, 3)) / det;
  inv[4] = (A(1, 1) * A(3, 3) - A(3, 1) * A(1, 3)) / det;
  inv[5] = (-(A(1, 1) * A(3, 2)) + A(3, 1) * A(1, 2)) / det;
  inv[6] = (A(1, 2) * A(2, 3) - A(2, 2) * A(1, 3)) / det;
  inv[7] = (-(A(1, 1) * A(2, 3)) + A(2, 1) * A(1, 3)) / det;
  inv[8] = (A(1, 1) * A(2, 2) - A(2, 1) * A(1, 2)) / det;
#undef A
  for (int m = 0; m < nod; ++m)
    for (int a = 0; a < 3; ++a) {
      double s = 0.0;
      for (int b = 0; b < 3; ++b) s = s + inv[b * 3 + a] * T.der[b * 20 + m];
      deriv[m * 3 + a] = s;
    }
  // eps = bee*eld, rows of bee from beemat (new_library.f90:976-993)
  double eps[6];
  for (int row = 0; row < 6; ++row) {
    double s = 0.0;
    for (int c = 0; c < ntot; ++c) {
      const int m = c / 3, comp = c % 3;
      const double x = deriv[m * 3], y = deriv[m * 3 + 1], z = deriv[m * 3 + 2];
      double b = 0.0;
      if (comp == 0) b = row == 0 ? x : row == 3 ? y : row == 5 ? z : 0.0;
      else if (comp == 1) b = row == 1 ? y : row == 3 ? x : row == 4 ? z : 0.0;
      else b = row == 2 ? z : row == 4 ? y : row == 5 ? x : 0.0;
      s = s + b * eld[(size_t)c];
    }
    eps[row] = s;
  }
  for (int row = 0; row < 6; ++row) {
    double s = 0.0;
    for (int c = 0; c < 6; ++c) s = s + T.dee[c * 6 + row] * eps[c];
    sigma6[row] = s;
  }
  return 0;
}

int pf_centroid_stress(pf_handle h, int64_t iel, double e, double v, double *sigma6) {
  return stress_at(h, iel, nullptr, e, v, sigma6);
}

int pf_point_stress(pf_handle h, int64_t iel, double xi, double eta, double zeta, double e, double v, double *sigma6) {
  const double pt[3] = {xi, eta, zeta};
  return stress_at(h, iel, pt, e, v, sigma6);
}

int pf_halo_transport(pf_handle h) {
  if (!h || h->nranks == 1) return 0;
  return (h->peer_ok && h->use_peer) ? 2 : 1;
}

int pf_get_last_solve_ms(pf_handle h, double *ms) {
  if (!h || !ms) return 1;
  *ms = h->last_ms;
  return 0;
}

}  // extern "C"

#include "xx3_compat.cuh"
