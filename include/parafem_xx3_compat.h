/*
 * parafem_xx3_compat.h -- the reference's EXISTING CUDA boundary, symbol for symbol.
 *
 * ParaFEM's GPU driver xx3 (parafem/src/programs/dev/xx3/xx3.f90) already binds six C functions
 * through `interface ... bind(C)` (xx3.f90:56-148) and links them from xx3/cuda_helpers.cu:183-368.
 * libparafem_b200.so exports the same names with the same signatures (all scalars by reference,
 * device pointers as `type(c_ptr)` passed by reference, return EXIT_SUCCESS / EXIT_FAILURE and a
 * printf message), so xx3.f90 links against it UNCHANGED:
 *
 *     $(FC) xx3.o -o xx3 $(LIB_MPI) -L$(PF_B200)/parafem_b200 -lparafem_b200       (instead of cuda_helpers.o)
 *
 * What differs behind the names:
 *   - byte counts are 64-bit products (cuda_helpers.cu:205,242,264 multiply two ints and overflow past
 *     596 523 20-node elements);
 *   - matrix_vector_multiplies runs the bulk-copy-ring mat-vec of this library (k_matvec, the kernel of
 *     pf_pcg_solve) for the element sizes of p121 / p123 (n_row = n_col = 60, 24, 8) and a plain row-per-
 *     thread kernel otherwise; either way, in the reference's naming, d_rhs_vector(:,e) =
 *     MATMUL(d_matrix(:,:,e), d_lhs_vector(:,e)) -- pmul_pp is uploaded to d_lhs_vector (xx3.f90:496-500), utemp_pp
 *     is read back from d_rhs_vector (:525-529), the kernel reads lhs_vector and writes rhs_vector
 *     (cuda_helpers.cu:168-176) -- with the column
 *     sweep j ascending and separate multiply / add, i.e. the bits of the Fortran loop at xx3.f90:476-480
 *     (the reference's kernel MultiMatVecMultiply1, cuda_helpers.cu:144-177, sums in the same order);
 *   - the call is synchronous as in the reference (cudaDeviceSynchronize, cuda_helpers.cu:360).
 * This is the fine-grained, PCIe-per-iteration interface xx3 has today; the state-holding API of
 * parafem_b200.h is the one to move to.
 */
#ifndef PARAFEM_XX3_COMPAT_H
#define PARAFEM_XX3_COMPAT_H
#ifdef __cplusplus
extern "C" {
#endif

int set_gpu(const int *device_id);                                              /* cuda_helpers.cu:183-196 */
int allocate_memory_on_gpu(const int *n_elements, const int *element_size,
                           void **device_pointer);                               /* :199-214 */
int free_memory_on_gpu(void **device_pointer);                                   /* :217-230 */
int copy_data_to_gpu(const int *n_elements, const int *element_size,
                     const void *host_data, void **device_pointer);              /* :233-252 */
int copy_data_from_gpu(const int *n_elements, const int *element_size,
                       void *host_data, void **device_pointer);                  /* :255-274 */
int matrix_vector_multiplies(int *n_mat, int *n_row, int *n_col, void **d_lhs_vector,
                             void **d_matrix, void **d_rhs_vector);              /* :277-368 */

#ifdef __cplusplus
}
#endif
#endif
