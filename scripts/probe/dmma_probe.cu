// Probe (round 2): FP64 tensor-core mma.sync on sm_100a -- throughput against the DFMA loop, and the rounding model
// of one m8n8k4 instruction (is D = fma(a3,b3,fma(a2,b2,fma(a1,b1,fma(a0,b0,c)))) bit for bit?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probe/dmma_probe scripts/probe/dmma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
__device__ __forceinline__ void dmma1688(double *d, const double *a, const double *b, const double *c) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
               : "=d"(d[0]), "=d"(d[1]), "=d"(d[2]), "=d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]), "d"(c[0]), "d"(c[1]), "d"(c[2]), "d"(c[3]));
}
__device__ __forceinline__ void dmma16816(double *d, const double *a, const double *b, const double *c) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%16,%17,%18,%19};"
               : "=d"(d[0]), "=d"(d[1]), "=d"(d[2]), "=d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]), "d"(c[0]), "d"(c[1]), "d"(c[2]), "d"(c[3]));
}

template <int CH>
__global__ void k_dmma884(double *out, int iters, double a, double b) {
  double acc[CH][2];
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i][0] = acc[i][1] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma884(acc[i][0], acc[i][1], a, b, acc[i][0], acc[i][1]);
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void k_dmma1688(double *out, int iters, double a, double b) {
  double acc[CH][4], A[4] = {a, a, a, a}, B[2] = {b, b};
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma1688(acc[i], A, B, acc[i]);
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void k_dmma16816(double *out, int iters, double a, double b) {
  double acc[CH][4], A[8] = {a, a, a, a, a, a, a, a}, B[4] = {b, b, b, b};
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma16816(acc[i], A, B, acc[i]);
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void k_dfma(double *out, int iters, double a, double b) {
  double acc[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = fma(acc[i], a, b);
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: DMMA and DFMA interleaved (do they share the pipe?)
template <int CH>
__global__ void k_mixed(double *out, int iters, double a, double b) {
  double acc[CH][2], f[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) { acc[i][0] = acc[i][1] = threadIdx.x * 1e-9 + i; f[i] = i; }
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      dmma884(acc[i][0], acc[i][1], a, b, acc[i][0], acc[i][1]);
#pragma unroll
      for (int r = 0; r < 8; ++r) f[i] = fma(f[i], a, b);
    }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += acc[i][0] + acc[i][1] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// rounding model: one warp, D = A(8x4) B(4x8) + C(8x8)
__global__ void k_one(const double *A, const double *B, const double *C, double *D) {
  const int l = threadIdx.x;
  const double a = A[(l / 4) * 4 + (l % 4)];        // A[row=l/4][k=l%4]
  const double b = B[(l % 4) * 8 + (l / 4)];        // B[k=l%4][n=l/4]
  const int r = l / 4, c = 2 * (l % 4);
  double d0, d1;
  dmma884(d0, d1, a, b, C[r * 8 + c], C[r * 8 + c + 1]);
  D[r * 8 + c] = d0; D[r * 8 + c + 1] = d1;
}
// m16n8k8: A 16x8 row: a0:(r=l/4, k=l%4) a1:(r+8, k) a2:(r, k+4) a3:(r+8, k+4); B 8x8: b0:(k=l%4, n=l/4) b1:(k+4, n); C: c0,c1:(r, 2(l%4)+{0,1}) c2,c3:(r+8, ..)
__global__ void k_one1688(const double *A, const double *B, const double *C, double *D) {
  const int l = threadIdx.x, r = l / 4, q = l % 4;
  double a[4] = {A[r * 8 + q], A[(r + 8) * 8 + q], A[r * 8 + q + 4], A[(r + 8) * 8 + q + 4]};
  double b[2] = {B[q * 8 + r], B[(q + 4) * 8 + r]};
  double c[4] = {C[r * 8 + 2 * q], C[r * 8 + 2 * q + 1], C[(r + 8) * 8 + 2 * q], C[(r + 8) * 8 + 2 * q + 1]}, d[4];
  dmma1688(d, a, b, c);
  D[r * 8 + 2 * q] = d[0]; D[r * 8 + 2 * q + 1] = d[1]; D[(r + 8) * 8 + 2 * q] = d[2]; D[(r + 8) * 8 + 2 * q + 1] = d[3];
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

template <typename F>
static float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 1024 * 4));
  const int iters = 20000;
  for (int warps : {4, 8, 12, 16, 32}) {
    const int thr = warps * 32, blocks = sms;
    float ms;
    ms = time_ms([&] { k_dfma<16><<<blocks, thr>>>(out, iters, 1.0000001, 1e-9); });
    printf("dfma   ch16 warps/SM %2d: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, 2.0 * 16 * iters * thr * blocks / ms * 1e-9);
    ms = time_ms([&] { k_dmma884<9><<<blocks, thr>>>(out, iters, 1.0000001, 1e-9); });
    printf("m8n8k4   ch9 warps/SM %2d: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, 2.0 * 256 * 9 * iters * warps * blocks / ms * 1e-9);
    ms = time_ms([&] { k_dmma884<3><<<blocks, thr>>>(out, iters, 1.0000001, 1e-9); });
    printf("m8n8k4   ch3 warps/SM %2d: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, 2.0 * 256 * 3 * iters * warps * blocks / ms * 1e-9);
    ms = time_ms([&] { k_dmma884<1><<<blocks, thr>>>(out, iters, 1.0000001, 1e-9); });
    printf("m8n8k4   ch1 warps/SM %2d: %8.3f ms  %7.2f TFLOP/s (latency %.1f ns/instr)\n", warps, ms, 2.0 * 256 * 1 * iters * warps * blocks / ms * 1e-9, ms * 1e6 / iters);
    ms = time_ms([&] { k_dmma1688<6><<<blocks, thr>>>(out, iters, 1.0000001, 1e-9); });
    printf("m16n8k8  ch6 warps/SM %2d: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, 2.0 * 1024 * 6 * iters * warps * blocks / ms * 1e-9);
    ms = time_ms([&] { k_dmma16816<6><<<blocks, thr>>>(out, iters, 1.0000001, 1e-9); });
    printf("m16n8k16 ch6 warps/SM %2d: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, 2.0 * 2048 * 6 * iters * warps * blocks / ms * 1e-9);
    ms = time_ms([&] { k_mixed<6><<<blocks, thr>>>(out, iters, 1.0000001, 1e-9); });
    printf("mixed (1 dmma + 8 dfma) ch6 warps/SM %2d: %8.3f ms  %7.2f TFLOP/s total\n", warps, ms,
           (2.0 * 256 * 6 + 2.0 * 8 * 6 * 32) * iters * warps * blocks / ms * 1e-9);
  }
  // rounding model
  double hA[128], hB[64], hC[128], hD[128], *dA, *dB, *dC, *dD;
  CK(cudaMalloc(&dA, sizeof hA)); CK(cudaMalloc(&dB, sizeof hB)); CK(cudaMalloc(&dC, sizeof hC)); CK(cudaMalloc(&dD, sizeof hD));
  srand(12345);
  auto rnd = [] { return (rand() / (double)RAND_MAX - 0.5) * ldexp(1.0, rand() % 9 - 4); };
  long bad_fwd = 0, bad_rev = 0, bad_c_last = 0, tot = 0;
  for (int trial = 0; trial < 2000; ++trial) {
    for (double &v : hA) v = rnd();
    for (double &v : hB) v = rnd();
    for (double &v : hC) v = rnd();
    cudaMemcpy(dA, hA, 32 * 8, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, 32 * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dC, hC, 64 * 8, cudaMemcpyHostToDevice);
    k_one<<<1, 32>>>(dA, dB, dC, dD);
    CK(cudaMemcpy(hD, dD, 64 * 8, cudaMemcpyDeviceToHost));
    for (int r = 0; r < 8; ++r)
      for (int c = 0; c < 8; ++c) {
        double f = hC[r * 8 + c], g = hC[r * 8 + c], h = 0.0;
        for (int k = 0; k < 4; ++k) f = fma(hA[r * 4 + k], hB[k * 8 + c], f);
        for (int k = 3; k >= 0; --k) g = fma(hA[r * 4 + k], hB[k * 8 + c], g);
        h = hA[r * 4] * hB[c];
        for (int k = 1; k < 4; ++k) h = fma(hA[r * 4 + k], hB[k * 8 + c], h);
        h += hC[r * 8 + c];
        ++tot;
        bad_fwd += memcmp(&f, &hD[r * 8 + c], 8) != 0;
        bad_rev += memcmp(&g, &hD[r * 8 + c], 8) != 0;
        bad_c_last += memcmp(&h, &hD[r * 8 + c], 8) != 0;
      }
  }
  printf("m8n8k4 rounding: of %ld entries, mismatches vs fma chain k=0..3 from c: %ld; k=3..0: %ld; c added last: %ld\n", tot, bad_fwd, bad_rev, bad_c_last);
  bad_fwd = 0; tot = 0; long bad_split = 0;
  for (int trial = 0; trial < 2000; ++trial) {
    for (double &v : hA) v = rnd();
    for (double &v : hB) v = rnd();
    for (double &v : hC) v = rnd();
    cudaMemcpy(dA, hA, 128 * 8, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, 64 * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dC, hC, 128 * 8, cudaMemcpyHostToDevice);
    k_one1688<<<1, 32>>>(dA, dB, dC, dD);
    CK(cudaMemcpy(hD, dD, 128 * 8, cudaMemcpyDeviceToHost));
    for (int r = 0; r < 16; ++r)
      for (int c = 0; c < 8; ++c) {
        double f = hC[r * 8 + c];
        for (int k = 0; k < 8; ++k) f = fma(hA[r * 8 + k], hB[k * 8 + c], f);
        ++tot;
        bad_fwd += memcmp(&f, &hD[r * 8 + c], 8) != 0;
      }
  }
  printf("m16n8k8 rounding: of %ld entries, mismatches vs fma chain k=0..7 from c: %ld\n", tot, bad_fwd);
  (void)bad_split;
  return 0;
}
