// xx3_compat.cuh -- include/parafem_xx3_compat.h: the six entry points ParaFEM's xx3 driver binds today
// (programs/dev/xx3/xx3.f90:56-148 -> xx3/cuda_helpers.cu:183-368), same names and signatures, on top of this
// library's mat-vec.  Included at the end of device.cu (one translation unit: the kernels share c_tab).
#pragma once
#include "parafem_xx3_compat.h"

namespace {

// lhs(:,e) = MATMUL(matrix(:,:,e), rhs(:,e)) for any n_row x n_col: one thread per row, j ascending from 0.0,
// separate multiply and add (consecutive threads read consecutive rows of a column: coalesced)
__global__ void k_matvec_any(const double *__restrict__ mat, const double *__restrict__ rhs, double *__restrict__ lhs,
                             long long n_mat, int n_row, int n_col) {
  const long long total = n_mat * n_row;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const long long e = q / n_row;
    const int i = (int)(q - e * n_row);
    const double *K = mat + e * (long long)n_row * n_col, *p = rhs + e * (long long)n_col;
    double s = 0.0;
    for (int j = 0; j < n_col; ++j) s = s + K[(long long)j * n_row + i] * p[j];
    lhs[q] = s;
  }
}

template <int NTOT, int EPT, int STAGES>
cudaError_t xx3_ring(const double *mat, const double *rhs, double *lhs, long long n_mat, int sms) {
  using Cfg = pf::MatvecCfg<NTOT, EPT, STAGES>;
  auto kern = pf::k_matvec<NTOT, EPT, STAGES, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem);
  if (e != cudaSuccess) return e;
  const long long ntiles = (n_mat + EPT - 1) / EPT;
  const int grid = (int)std::max<long long>(1, std::min<long long>(sms, ntiles));
  kern<<<grid, Cfg::kThreads, Cfg::kSmem>>>(mat, nullptr, rhs, lhs, n_mat, nullptr, nullptr);
  return cudaGetLastError();
}

int xx3_fail(const char *msg, cudaError_t e) {
  printf("%s (%s)\n", msg, cudaGetErrorString(e));
  return EXIT_FAILURE;
}

}  // namespace

extern "C" {

int set_gpu(const int *device_id) {
  const cudaError_t e = cudaSetDevice(*device_id);
  return e == cudaSuccess ? EXIT_SUCCESS : xx3_fail("Failed to set device!", e);
}

int allocate_memory_on_gpu(const int *n_elements, const int *element_size, void **device_pointer) {
  const cudaError_t e = cudaMalloc(device_pointer, (size_t)*element_size * (size_t)*n_elements);
  return e == cudaSuccess ? EXIT_SUCCESS : xx3_fail("Device memory failed to allocate!", e);
}

int free_memory_on_gpu(void **device_pointer) {
  const cudaError_t e = cudaFree(*device_pointer);
  return e == cudaSuccess ? EXIT_SUCCESS : xx3_fail("Device memory failed to deallocate!", e);
}

int copy_data_to_gpu(const int *n_elements, const int *element_size, const void *host_data, void **device_pointer) {
  const cudaError_t e = cudaMemcpy(*device_pointer, host_data, (size_t)*n_elements * (size_t)*element_size, cudaMemcpyHostToDevice);
  return e == cudaSuccess ? EXIT_SUCCESS : xx3_fail("Failed to copy data to device!", e);
}

int copy_data_from_gpu(const int *n_elements, const int *element_size, void *host_data, void **device_pointer) {
  const cudaError_t e = cudaMemcpy(host_data, *device_pointer, (size_t)*n_elements * (size_t)*element_size, cudaMemcpyDeviceToHost);
  return e == cudaSuccess ? EXIT_SUCCESS : xx3_fail("Failed to copy data from device!", e);
}

int matrix_vector_multiplies(int *n_mat, int *n_row, int *n_col, void **d_lhs_vector, void **d_matrix, void **d_rhs_vector) {
  if (*n_mat < 1 || *n_row < 1 || *n_col < 1) { printf("Failed to launch multiply kernel! (bad sizes)\n"); return EXIT_FAILURE; }
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return xx3_fail("Failed to launch multiply kernel!", e);
  // The reference's naming: d_lhs_vector is the LEFT-HAND OPERAND (pmul_pp goes in, xx3.f90:496-500) and
  // d_rhs_vector receives the products (utemp_pp comes out, xx3.f90:525-529; cuda_helpers.cu:172-176).
  const double *mat = static_cast<const double *>(*d_matrix), *rhs = static_cast<const double *>(*d_lhs_vector);
  double *lhs = static_cast<double *>(*d_rhs_vector);
  const long long n = *n_mat;
  if (*n_row == *n_col && *n_row == 60) e = xx3_ring<60, 1, 7>(mat, rhs, lhs, n, sms);
  else if (*n_row == *n_col && *n_row == 24) e = xx3_ring<24, 1, 32>(mat, rhs, lhs, n, sms);
  else if (*n_row == *n_col && *n_row == 8) e = xx3_ring<8, 16, 16>(mat, rhs, lhs, n, sms);
  else {
    const long long total = n * *n_row;
    const int grid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)sms * 8));
    k_matvec_any<<<grid, 256>>>(mat, rhs, lhs, n, *n_row, *n_col);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) return xx3_fail("Failed to launch multiply kernel!", e);
  e = cudaDeviceSynchronize();                       // the reference's call is synchronous (cuda_helpers.cu:360)
  return e == cudaSuccess ? EXIT_SUCCESS : xx3_fail("Failed to synchronise!", e);
}

}  // extern "C"
