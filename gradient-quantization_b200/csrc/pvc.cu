// pvc.cu -- probabilistic vector compressor search (stage 2 of the residual
// compressor).  INTENDED semantics of
// compressors/probabilistic_vector_compressor.py:42-65 (the shipped class cannot
// run: wrong codebook directory and argmin on a bool tensor):
//   p = pinv(C^T) v ; l1 = sum_k |p_k| ; code = first k with
//   cumsum(|p|/l1)_k >= r - 1e-5 (last k if none) ; u = sign(p_code) * l1.
// Sums run sequentially in ascending k (a defined order shared with the CPU
// oracle; parity with the reference is unpinned for this class).
#include "gq_internal.cuh"

namespace gq {

template <int D>
__device__ __forceinline__ float pvc_score(const float4 *__restrict__ cw, const float (&v)[D])
{
    float4 c0 = cw[0];
    float acc = __fmul_rn(c0.x, v[0]);
    acc = __fmaf_rn(c0.y, v[1], acc);
    acc = __fmaf_rn(c0.z, v[2], acc);
    acc = __fmaf_rn(c0.w, v[3], acc);
#pragma unroll
    for (int q = 1; q < D / 4; ++q) {
        float4 cq = cw[q];
        acc = __fmaf_rn(cq.x, v[4 * q + 0], acc);
        acc = __fmaf_rn(cq.y, v[4 * q + 1], acc);
        acc = __fmaf_rn(cq.z, v[4 * q + 2], acc);
        acc = __fmaf_rn(cq.w, v[4 * q + 3], acc);
    }
    return acc;
}

template <int D, typename CodeT>
__global__ void __launch_bounds__(256)
pvc_search_kernel(const float *__restrict__ grad, int64_t n_chunks, const float *__restrict__ dagger,
                  int K, int k_tile, const float *__restrict__ uniforms, uint64_t seed, uint64_t offset,
                  CodeT *__restrict__ codes, float *__restrict__ u_out)
{
    extern __shared__ float4 s_dg4[];
    constexpr int D4 = D / 4;
    const int tid = threadIdx.x;
    for (int64_t base = (int64_t)blockIdx.x * 256; base < n_chunks; base += (int64_t)gridDim.x * 256) {
        const int64_t c = base + tid;
        const bool valid = c < n_chunks;
        float v[D];
#pragma unroll
        for (int q = 0; q < D4; ++q) {
            float4 t = valid ? ld_stream_f4(reinterpret_cast<const float4 *>(grad + c * D) + q)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
        float l1 = 0.0f;
        float cum = 0.0f, thr = 0.0f, sel_p = 0.0f, last_p = 0.0f;
        int code = -1;
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1 && valid) {
                float r = uniforms ? __ldg(uniforms + c) : philox_uniform(seed, offset, (uint64_t)c);
                thr = __fsub_rn(r, 1e-5f);
            }
            for (int k0 = 0; k0 < K; k0 += k_tile) {
                const int kt = min(k_tile, K - k0);
                __syncthreads();
                const float4 *dg = reinterpret_cast<const float4 *>(dagger) + (int64_t)k0 * D4;
                for (int i = tid; i < kt * D4; i += 256) s_dg4[i] = __ldg(dg + i);
                __syncthreads();
                if (pass == 0) {
                    for (int k = 0; k < kt; ++k) l1 = __fadd_rn(l1, fabsf(pvc_score<D>(s_dg4 + k * D4, v)));
                } else if (code < 0) {
                    for (int k = 0; k < kt; ++k) {
                        float p = pvc_score<D>(s_dg4 + k * D4, v);
                        last_p = p;
                        cum = __fadd_rn(cum, __fdiv_rn(fabsf(p), l1));
                        if (cum >= thr) { code = k0 + k; sel_p = p; break; }
                    }
                }
            }
        }
        if (code < 0) { code = K - 1; sel_p = last_p; }
        if (valid) {
            codes[c] = (CodeT)code;
            float sg = (float)((sel_p > 0.0f) - (sel_p < 0.0f));
            u_out[c] = __fmul_rn(sg, l1);
        }
    }
}

template <int D, typename CodeT>
static int launch_pvc_d(const float *grad, int64_t n_chunks, const float *dagger, int K,
                        const float *uniforms, uint64_t seed, uint64_t offset, CodeT *codes, float *u_out,
                        cudaStream_t st)
{
    const int max_tile = (64 * 1024) / (D * 4);
    const int k_tile = K <= max_tile ? K : max_tile;
    const size_t smem = (size_t)k_tile * D * 4;
    auto kern = pvc_search_kernel<D, CodeT>;
    GQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (n_chunks + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 4;
    int grid = (int)(blocks < cap ? blocks : cap);
    kern<<<grid < 1 ? 1 : grid, 256, smem, st>>>(grad, n_chunks, dagger, K, k_tile, uniforms, seed, offset,
                                                 codes, u_out);
    GQ_LAUNCH_CHECK("pvc_search");
    return GQ_OK;
}

template <typename CodeT>
static int launch_pvc(const float *grad, int64_t n_chunks, int d, const float *dagger, int K,
                      const float *uniforms, uint64_t seed, uint64_t offset, CodeT *codes, float *u_out,
                      cudaStream_t st)
{
    switch (d) {
#define GQ_CASE(DD) case DD: return launch_pvc_d<DD, CodeT>(grad, n_chunks, dagger, K, uniforms, seed, offset, codes, u_out, st);
        GQ_CASE(4) GQ_CASE(8) GQ_CASE(12) GQ_CASE(16) GQ_CASE(24) GQ_CASE(32) GQ_CASE(48) GQ_CASE(64)
#undef GQ_CASE
        default: break;
    }
    set_error("pvc_search: chunk dim %d not supported (4, 8, 12, 16, 24, 32, 48, 64)", d);
    return GQ_ERR_UNSUPPORTED;
}

}  // namespace gq

using namespace gq;

extern "C" int gq_pvc_search(const float *grad, int64_t n_chunks, int d, const float *dagger, int K,
                             const float *uniforms, uint64_t philox_seed, uint64_t philox_offset,
                             void *codes, int code_bytes, float *u_out, gq_stream_t stream)
{
    GQ_REQUIRE(n_chunks >= 0 && d >= 1 && K >= 1, "bad sizes");
    GQ_REQUIRE(n_chunks == 0 || (grad && dagger && codes && u_out), "null pointer");
    GQ_REQUIRE(code_bytes == 1 || code_bytes == 4, "code_bytes must be 1 or 4");
    GQ_REQUIRE(code_bytes == 4 || K <= 256, "uint8 codes need K <= 256");
    GQ_REQUIRE(((uintptr_t)grad & 15) == 0, "gradient must be 16-byte aligned");
    if (n_chunks == 0) return GQ_OK;
    if (code_bytes == 1)
        return launch_pvc<uint8_t>(grad, n_chunks, d, dagger, K, uniforms, philox_seed, philox_offset,
                                   (uint8_t *)codes, u_out, as_stream(stream));
    return launch_pvc<int32_t>(grad, n_chunks, d, dagger, K, uniforms, philox_seed, philox_offset,
                               (int32_t *)codes, u_out, as_stream(stream));
}
