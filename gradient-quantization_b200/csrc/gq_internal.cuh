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

// Per-chunk epilogue shared by all search kernels: store code and u, fold u into
// the per-tensor min/max keys (lb/ub of the norm quantizer).
template <typename CodeT>
__device__ __forceinline__ void search_epilogue(bool valid, int64_t c, int best_k, float best_u,
                                                CodeT *__restrict__ codes, float *__restrict__ u_out,
                                                const int64_t *__restrict__ seg_start, int n_seg,
                                                uint32_t *__restrict__ minmax_keys, SegCache &segc)
{
    if (valid) {
        codes[c] = (CodeT)best_k;
        u_out[c] = best_u;
    }
    if (minmax_keys == nullptr) return;
    int seg = valid ? cached_segment(segc, seg_start, n_seg, c) : -1;
    // warp-uniform fast path: every valid lane in the same tensor
    int seg0 = __shfl_sync(0xffffffffu, seg, 0);
    bool uniform = __all_sync(0xffffffffu, (seg == seg0) || !valid) && (seg0 >= 0);
    if (uniform) {
        float mn = warp_min(valid ? best_u : INFINITY);
        float mx = warp_max(valid ? best_u : -INFINITY);
        if ((threadIdx.x & 31) == 0) {
            atomicMin(minmax_keys + 2 * seg0, float_to_key(mn));
            atomicMax(minmax_keys + 2 * seg0 + 1, float_to_key(mx));
        }
    } else if (valid) {
        atomicMin(minmax_keys + 2 * seg, float_to_key(best_u));
        atomicMax(minmax_keys + 2 * seg + 1, float_to_key(best_u));
    }
}


// Running per-tensor min/max of u kept in registers while a warp walks CONSECUTIVE chunks
// (the tcgen05 kernel gives every CTA a contiguous range of tiles): one atomic pair per warp
// per tensor instead of one per tile.  Both functions are warp-collective.
struct MinMaxAcc {
    int seg = -1;
    float mn = INFINITY, mx = -INFINITY;
};
__device__ __forceinline__ void minmax_flush_warp(MinMaxAcc &a, uint32_t *__restrict__ keys)
{
    const uint32_t have = __ballot_sync(0xffffffffu, a.seg >= 0);
    if (have != 0u) {
        const int seg0 = __shfl_sync(0xffffffffu, a.seg, __ffs(have) - 1);
        const bool uniform = __all_sync(0xffffffffu, a.seg == seg0 || a.seg < 0);
        if (uniform) {
            const float mn = warp_min(a.mn), mx = warp_max(a.mx);
            if ((threadIdx.x & 31) == 0) {
                atomicMin(keys + 2 * seg0, float_to_key(mn));
                atomicMax(keys + 2 * seg0 + 1, float_to_key(mx));
            }
        } else if (a.seg >= 0) {
            atomicMin(keys + 2 * a.seg, float_to_key(a.mn));
            atomicMax(keys + 2 * a.seg + 1, float_to_key(a.mx));
        }
    }
    a.seg = -1;
    a.mn = INFINITY;
    a.mx = -INFINITY;
}
__device__ __forceinline__ void minmax_add_warp(MinMaxAcc &a, bool valid, int seg, float u,
                                                uint32_t *__restrict__ keys)
{
    if (__any_sync(0xffffffffu, valid && a.seg >= 0 && seg != a.seg)) minmax_flush_warp(a, keys);
    if (valid) {
        a.seg = seg;
        a.mn = fminf(a.mn, u);
        a.mx = fmaxf(a.mx, u);
    }
}

}  // namespace gq
