// hsq_exact.cu -- fp32 CUDA-core nearest-codeword search (any d <= 128, any K).
//
// This is the bit-exact definition of the search on the GPU: every score is the
// sequential ascending-j fp32 FMA chain the reference's CPU torch.mm produces
// (compressors/nearest_neighbor_compressor.py:68), the winner is
// argmax |score| with the first index winning ties (:69-72) and u is the signed
// score of the winner (:73).  The tcgen05 kernel (hsq_tc.cu) uses the same
// chain to rescore its candidates, and falls back to this kernel for shapes it
// does not cover.  FMA-pipe bound: 2*K flop per gradient element.
#include "gq_internal.cuh"

namespace gq {

constexpr int kSearchThreads = 256;

// One thread per chunk, chunk in registers, codebook tile in shared memory
// (broadcast LDS.128).  Four independent FMA chains per iteration for ILP.
template <int D, typename CodeT>
__global__ void __launch_bounds__(kSearchThreads)
hsq_search_exact_kernel(const float *__restrict__ grad, int64_t n_chunks,
                        const float *__restrict__ codebook, int K, int k_tile,
                        CodeT *__restrict__ codes, float *__restrict__ u_out,
                        const int64_t *__restrict__ seg_start, int n_seg,
                        uint32_t *__restrict__ minmax_keys)
{
    extern __shared__ float4 s_cb4[];  // [k_tile][D/4]
    constexpr int D4 = D / 4;
    const int tid = threadIdx.x;
    const bool single_tile = (k_tile >= K);

    if (single_tile) {
        const float4 *cb4 = reinterpret_cast<const float4 *>(codebook);
        for (int i = tid; i < K * D4; i += kSearchThreads) s_cb4[i] = __ldg(cb4 + i);
        __syncthreads();
    }

    SegCache segc;
    for (int64_t base = (int64_t)blockIdx.x * kSearchThreads; base < n_chunks;
         base += (int64_t)gridDim.x * kSearchThreads) {
        const int64_t c = base + tid;
        const bool valid = c < n_chunks;
        float v[D];
        if (valid) {
            const float4 *g4 = reinterpret_cast<const float4 *>(grad + c * D);
#pragma unroll
            for (int q = 0; q < D4; ++q) {
                float4 t = ld_stream_f4(g4 + q);
                v[4 * q + 0] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < D; ++j) v[j] = 0.0f;
        }
        int best_bits = -1;  // |score| bit pattern of the current winner (as int: NaN > inf > finite)
        int best_k = 0;
        float best_u = 0.0f;

        for (int k0 = 0; k0 < K; k0 += k_tile) {
            const int kt = min(k_tile, K - k0);
            if (!single_tile) {
                __syncthreads();
                const float4 *cb4 = reinterpret_cast<const float4 *>(codebook) + (int64_t)k0 * D4;
                for (int i = tid; i < kt * D4; i += kSearchThreads) s_cb4[i] = __ldg(cb4 + i);
                __syncthreads();
            }
            int k = 0;
            for (; k + 4 <= kt; k += 4) {
                float p[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float4 *cw = s_cb4 + (k + t) * D4;
                    float4 c0 = cw[0];
                    float acc = __fmul_rn(c0.x, v[0]);
                    acc = __fmaf_rn(c0.y, v[1], acc);
                    acc = __fmaf_rn(c0.z, v[2], acc);
                    acc = __fmaf_rn(c0.w, v[3], acc);
#pragma unroll
                    for (int q = 1; q < D4; ++q) {
                        float4 cq = cw[q];
                        acc = __fmaf_rn(cq.x, v[4 * q + 0], acc);
                        acc = __fmaf_rn(cq.y, v[4 * q + 1], acc);
                        acc = __fmaf_rn(cq.z, v[4 * q + 2], acc);
                        acc = __fmaf_rn(cq.w, v[4 * q + 3], acc);
                    }
                    p[t] = acc;
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    int ab = __float_as_int(p[t]) & 0x7fffffff;
                    if (ab > best_bits) { best_bits = ab; best_k = k0 + k + t; best_u = p[t]; }
                }
            }
            for (; k < kt; ++k) {
                const float4 *cw = s_cb4 + k * D4;
                float4 c0 = cw[0];
                float acc = __fmul_rn(c0.x, v[0]);
                acc = __fmaf_rn(c0.y, v[1], acc);
                acc = __fmaf_rn(c0.z, v[2], acc);
                acc = __fmaf_rn(c0.w, v[3], acc);
#pragma unroll
                for (int q = 1; q < D4; ++q) {
                    float4 cq = cw[q];
                    acc = __fmaf_rn(cq.x, v[4 * q + 0], acc);
                    acc = __fmaf_rn(cq.y, v[4 * q + 1], acc);
                    acc = __fmaf_rn(cq.z, v[4 * q + 2], acc);
                    acc = __fmaf_rn(cq.w, v[4 * q + 3], acc);
                }
                int ab = __float_as_int(acc) & 0x7fffffff;
                if (ab > best_bits) { best_bits = ab; best_k = k0 + k; best_u = acc; }
            }
        }
        search_epilogue<CodeT>(valid, c, best_k, best_u, codes, u_out, seg_start, n_seg, minmax_keys, segc);
    }
}

// Any d (not a multiple of 4, or > 64): chunk staged in shared memory with an
// odd row pitch (conflict-free), codeword read from shared memory as scalars.
// Slow path for the escalated dims of nearest_neighbor_compressor.py:27-29.
template <typename CodeT>
__global__ void __launch_bounds__(128)
hsq_search_generic_kernel(const float *__restrict__ grad, int64_t n_chunks, int d,
                          const float *__restrict__ codebook, int K, int k_tile,
                          CodeT *__restrict__ codes, float *__restrict__ u_out,
                          const int64_t *__restrict__ seg_start, int n_seg,
                          uint32_t *__restrict__ minmax_keys)
{
    extern __shared__ float s_gen[];
    const int pitch = d | 1;
    float *s_v = s_gen;                  // [128][pitch]
    float *s_cb = s_gen + 128 * pitch;   // [k_tile][d]
    const int tid = threadIdx.x;

    SegCache segc;
    for (int64_t base = (int64_t)blockIdx.x * 128; base < n_chunks; base += (int64_t)gridDim.x * 128) {
        __syncthreads();
        // coalesced staging of 128 chunks
        const int64_t e0 = base * d;
        const int64_t e_end = min((base + 128) * (int64_t)d, n_chunks * (int64_t)d);
        for (int64_t e = e0 + tid; e < e_end; e += 128) {
            int64_t rel = e - e0;
            s_v[(rel / d) * pitch + (rel % d)] = grad[e];
        }
        const int64_t c = base + tid;
        const bool valid = c < n_chunks;
        int best_bits = -1, best_k = 0;
        float best_u = 0.0f;
        const float *myv = s_v + tid * pitch;
        for (int k0 = 0; k0 < K; k0 += k_tile) {
            const int kt = min(k_tile, K - k0);
            __syncthreads();
            for (int i = tid; i < kt * d; i += 128) s_cb[i] = __ldg(codebook + (int64_t)k0 * d + i);
            __syncthreads();
            if (valid) {
                for (int k = 0; k < kt; ++k) {
                    const float *cw = s_cb + k * d;
                    float acc = __fmul_rn(cw[0], myv[0]);
                    for (int j = 1; j < d; ++j) acc = __fmaf_rn(cw[j], myv[j], acc);
                    int ab = __float_as_int(acc) & 0x7fffffff;
                    if (ab > best_bits) { best_bits = ab; best_k = k0 + k; best_u = acc; }
                }
            }
        }
        search_epilogue<CodeT>(valid, c, best_k, best_u, codes, u_out, seg_start, n_seg, minmax_keys, segc);
    }
}

__global__ void minmax_init_kernel(uint32_t *keys, int n_seg, uint32_t *barrier, const Rider rider)
{
    pdl_launch_dependents();
    pdl_wait();   // the keys may still be read by the previous step's quantize kernel
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && barrier != nullptr) *barrier = 0u;
    if (i < n_seg) {
        keys[2 * i] = GQ_KEY_MIN_INIT;
        keys[2 * i + 1] = GQ_KEY_MAX_INIT;
    }
    rider_run(rider, i, (int64_t)gridDim.x * blockDim.x);
}

int launch_minmax_init_rider(uint32_t *keys, int n_seg, cudaStream_t st, uint32_t *barrier, const Rider &rider)
{
    int64_t blocks = (n_seg + 127) / 128;
    const int64_t rb = (rider.n + 127) / 128;
    if (rb > blocks) blocks = rb;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    GQ_CUDA(launch_pdl(minmax_init_kernel, dim3((unsigned)blocks), dim3(128), 0, st, keys, n_seg, barrier, rider));
    return GQ_OK;
}

int launch_minmax_init(uint32_t *keys, int n_seg, cudaStream_t st, uint32_t *barrier)
{
    Rider none = {};
    return launch_minmax_init_rider(keys, n_seg, st, barrier, none);
}

template <int D, typename CodeT>
static int launch_exact_d(const float *grad, int64_t n_chunks, const float *codebook, int K,
                          CodeT *codes, float *u_out, const int64_t *seg_start, int n_seg,
                          uint32_t *minmax_keys, cudaStream_t st)
{
    // codebook tile: whole codebook when it fits in 64 KB, else 64 KB tiles
    const int max_tile = (64 * 1024) / (D * 4);
    const int k_tile = K <= max_tile ? K : max_tile;
    const size_t smem = (size_t)k_tile * D * 4;
    auto kern = hsq_search_exact_kernel<D, CodeT>;
    GQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (n_chunks + kSearchThreads - 1) / kSearchThreads;
    int64_t cap = (int64_t)sm_count() * 8;
    int grid = (int)(blocks < cap ? blocks : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, kSearchThreads, smem, st>>>(grad, n_chunks, codebook, K, k_tile, codes, u_out,
                                             seg_start, n_seg, minmax_keys);
    GQ_LAUNCH_CHECK("hsq_search_exact");
    return GQ_OK;
}

template <typename CodeT>
static int launch_exact(const float *grad, int64_t n_chunks, int d, const float *codebook, int K,
                        CodeT *codes, float *u_out, const int64_t *seg_start, int n_seg,
                        uint32_t *minmax_keys, cudaStream_t st)
{
    switch (d) {
#define GQ_CASE(DD) case DD: return launch_exact_d<DD, CodeT>(grad, n_chunks, codebook, K, codes, u_out, seg_start, n_seg, minmax_keys, st);
        GQ_CASE(4) GQ_CASE(8) GQ_CASE(12) GQ_CASE(16) GQ_CASE(24) GQ_CASE(32) GQ_CASE(48) GQ_CASE(64)
#undef GQ_CASE
        default: break;
    }
    GQ_REQUIRE(d >= 1 && d <= 128, "hsq_search: chunk dim %d not supported (1..128)", d);
    const int pitch = d | 1;
    int k_tile = (32 * 1024) / (d * 4);
    if (k_tile > K) k_tile = K;
    if (k_tile < 1) k_tile = 1;
    const size_t smem = (size_t)(128 * pitch + k_tile * d) * 4;
    auto kern = hsq_search_generic_kernel<CodeT>;
    GQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (n_chunks + 127) / 128;
    int64_t cap = (int64_t)sm_count() * 8;
    int grid = (int)(blocks < cap ? blocks : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, 128, smem, st>>>(grad, n_chunks, d, codebook, K, k_tile, codes, u_out, seg_start,
                                  n_seg, minmax_keys);
    GQ_LAUNCH_CHECK("hsq_search_generic");
    return GQ_OK;
}

int hsq_search_exact(const float *grad, int64_t n_chunks, int d, const float *codebook, int K,
                     void *codes, int code_bytes, float *u_out, const int64_t *seg_start, int n_seg,
                     uint32_t *minmax_keys, cudaStream_t st)
{
    if (n_chunks == 0) return GQ_OK;
    if (code_bytes == 1)
        return launch_exact<uint8_t>(grad, n_chunks, d, codebook, K, (uint8_t *)codes, u_out,
                                     seg_start, n_seg, minmax_keys, st);
    return launch_exact<int32_t>(grad, n_chunks, d, codebook, K, (int32_t *)codes, u_out, seg_start,
                                 n_seg, minmax_keys, st);
}

}  // namespace gq
