// gq_internal.cuh -- declarations shared between the translation units of libgqb200.
#pragma once
#include "gq_common.cuh"

namespace gq {

// hsq_exact.cu
int launch_minmax_init(uint32_t *keys, int n_seg, cudaStream_t st);
int hsq_search_exact(const float *grad, int64_t n_chunks, int d, const float *codebook, int K,
                     void *codes, int code_bytes, float *u_out, const int64_t *seg_start, int n_seg,
                     uint32_t *minmax_keys, cudaStream_t st);

// hsq_tc.cu (tcgen05 path; d == 16, K == 256)
bool hsq_tc_supported(int d, int K, int code_bytes);
size_t hsq_tc_workspace_bytes(int64_t n_chunks);
int hsq_search_tc(const float *grad, int64_t n_chunks, int d, const float *codebook, int K,
                  void *codes, int code_bytes, float *u_out, const int64_t *seg_start, int n_seg,
                  uint32_t *minmax_keys, void *workspace, size_t workspace_bytes, cudaStream_t st);

// hsq_tail.cu
int launch_seg_minmax(const float *u, int64_t n, const int64_t *seg_start, int n_seg, uint32_t *keys,
                      cudaStream_t st);
int launch_norm_quantize(const float *u, int64_t n, const int64_t *seg_start, int n_seg, int n_bit,
                         int random, const float *uniforms, uint64_t seed, uint64_t offset, void *l,
                         int l_bytes, float *lbub, const uint32_t *keys, cudaStream_t st);
int launch_norm_dequantize(const void *l, int l_bytes, int64_t n, const int64_t *seg_start, int n_seg,
                           int n_bit, const float *lbub, float *out, cudaStream_t st);
int hsq_decode_reduce(const void *codes, int code_bytes, const void *l, int l_bytes, const float *lbub,
                      const float *norms_f32, int64_t user_stride, int n_users, int64_t n_chunks, int d,
                      const float *codebook, int K, const int64_t *seg_start, int n_seg, int n_bit,
                      int mean, int accumulate, float *out, cudaStream_t st);
int launch_f32_reduce_users(const float *in, int64_t user_stride, int n_users, int64_t n, int mean,
                            int accumulate, float *out, cudaStream_t st);
int launch_axpy(const float *a, const float *b, float alpha, int64_t n, float *out, int sub,
                cudaStream_t st);

}  // namespace gq
