// hsq_tail.cu -- everything of the HSQ codec after the search:
//   * per-tensor min/max of u (when the search kernel did not fold it in)
//   * n-bit norm quantization   (probabilistic_scalar_compressor.py:12-27)
//   * norm dequantization       (probabilistic_scalar_compressor.py:29-33)
//   * fused decode-and-reduce over users
//       (nearest_neighbor_compressor.py:80-90 + ps_quantizer.py:48 /
//        ring_quantizer.py:31-32)
// All HBM-bound.  Algorithmic bytes per gradient element (d = chunk dim, U users):
//   quantize: (4 u + 1 l [+4 r]) / d        decode-reduce: 4 + 2U/d (+4 if accumulate)
#include "gq_common.cuh"

namespace gq {

// ------------------------------------------------------- segmented min/max ---
__global__ void __launch_bounds__(256)
seg_minmax_kernel(const float *__restrict__ u, int64_t n, const int64_t *__restrict__ seg_start,
                  int n_seg, uint32_t *__restrict__ keys)
{
    for (int64_t base = (int64_t)blockIdx.x * 256; base < n; base += (int64_t)gridDim.x * 256) {
        int64_t i = base + threadIdx.x;
        bool valid = i < n;
        float x = valid ? u[i] : 0.0f;
        int seg = valid ? find_segment(seg_start, n_seg, i) : -1;
        int seg0 = __shfl_sync(0xffffffffu, seg, 0);
        bool uniform = __all_sync(0xffffffffu, (seg == seg0) || !valid) && (seg0 >= 0);
        if (uniform) {
            float mn = warp_min(valid ? x : INFINITY);
            float mx = warp_max(valid ? x : -INFINITY);
            if ((threadIdx.x & 31) == 0) {
                atomicMin(keys + 2 * seg0, float_to_key(mn));
                atomicMax(keys + 2 * seg0 + 1, float_to_key(mx));
            }
        } else if (valid) {
            atomicMin(keys + 2 * seg, float_to_key(x));
            atomicMax(keys + 2 * seg + 1, float_to_key(x));
        }
    }
}

// ---------------------------------------------------------- norm quantize ---
template <typename LT>
__global__ void __launch_bounds__(256)
norm_quantize_kernel(const float *__restrict__ u, int64_t n, const int64_t *__restrict__ seg_start,
                     int n_seg, float s, int random, const float *__restrict__ uniforms,
                     uint64_t seed, uint64_t offset, LT *__restrict__ l, float *__restrict__ lbub,
                     const uint32_t *__restrict__ keys)
{
    // block 0 also publishes lb/ub as floats (the reference returns them, :27)
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < 2 * n_seg; i += 256) lbub[i] = key_to_float(keys[i]);
    }
    for (int64_t base = (int64_t)blockIdx.x * 256; base < n; base += (int64_t)gridDim.x * 256) {
        int64_t i = base + threadIdx.x;
        if (i >= n) continue;
        int seg = find_segment(seg_start, n_seg, i);
        float lb = key_to_float(__ldg(keys + 2 * seg));
        float ub = key_to_float(__ldg(keys + 2 * seg + 1));
        float r = 0.0f;
        if (random) r = uniforms ? __ldg(uniforms + i) : philox_uniform(seed, offset, (uint64_t)i);
        l[i] = (LT)psc_level(u[i], lb, ub, s, random, r);
    }
}

template <typename LT>
__global__ void __launch_bounds__(256)
norm_dequantize_kernel(const LT *__restrict__ l, int64_t n, const int64_t *__restrict__ seg_start,
                       int n_seg, float s, const float *__restrict__ lbub, float *__restrict__ out)
{
    for (int64_t base = (int64_t)blockIdx.x * 256; base < n; base += (int64_t)gridDim.x * 256) {
        int64_t i = base + threadIdx.x;
        if (i >= n) continue;
        int seg = find_segment(seg_start, n_seg, i);
        out[i] = psc_value((int)l[i], __ldg(lbub + 2 * seg), __ldg(lbub + 2 * seg + 1), s);
    }
}

static int grid_for(int64_t n, int per_block, int waves = 8)
{
    int64_t blocks = (n + per_block - 1) / per_block;
    int64_t cap = (int64_t)sm_count() * waves;
    int64_t g = blocks < cap ? blocks : cap;
    return (int)(g < 1 ? 1 : g);
}

int launch_seg_minmax(const float *u, int64_t n, const int64_t *seg_start, int n_seg, uint32_t *keys,
                      cudaStream_t st)
{
    if (n == 0) return GQ_OK;
    seg_minmax_kernel<<<grid_for(n, 256), 256, 0, st>>>(u, n, seg_start, n_seg, keys);
    GQ_LAUNCH_CHECK("seg_minmax");
    return GQ_OK;
}

int launch_norm_quantize(const float *u, int64_t n, const int64_t *seg_start, int n_seg, int n_bit,
                         int random, const float *uniforms, uint64_t seed, uint64_t offset, void *l,
                         int l_bytes, float *lbub, const uint32_t *keys, cudaStream_t st)
{
    const float s = (float)(1u << n_bit);
    const int grid = grid_for(n > 0 ? n : 1, 256);
    if (l_bytes == 1)
        norm_quantize_kernel<uint8_t><<<grid, 256, 0, st>>>(u, n, seg_start, n_seg, s, random, uniforms,
                                                            seed, offset, (uint8_t *)l, lbub, keys);
    else
        norm_quantize_kernel<int32_t><<<grid, 256, 0, st>>>(u, n, seg_start, n_seg, s, random, uniforms,
                                                            seed, offset, (int32_t *)l, lbub, keys);
    GQ_LAUNCH_CHECK("norm_quantize");
    return GQ_OK;
}

int launch_norm_dequantize(const void *l, int l_bytes, int64_t n, const int64_t *seg_start, int n_seg,
                           int n_bit, const float *lbub, float *out, cudaStream_t st)
{
    if (n == 0) return GQ_OK;
    const float s = (float)(1u << n_bit);
    const int grid = grid_for(n, 256);
    if (l_bytes == 1)
        norm_dequantize_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t *)l, n, seg_start, n_seg, s, lbub, out);
    else
        norm_dequantize_kernel<int32_t><<<grid, 256, 0, st>>>((const int32_t *)l, n, seg_start, n_seg, s, lbub, out);
    GQ_LAUNCH_CHECK("norm_dequantize");
    return GQ_OK;
}

// ------------------------------------------------------- decode-and-reduce ---
// One warp owns 32 consecutive chunks.  Phase A: lane <-> chunk, per user load
// (code, l) and dequantize the norm once.  Phase B: lane <-> float4 of the
// output row; (code, norm) of the owning chunk arrive by shuffle, the codeword
// comes from shared memory (K*d*4 <= 64 KB) or L1/L2, products are added in
// user order u = 0..U-1 with separately rounded mul and add, exactly like
// torch.mul + stack().mean(0) of the reference.
constexpr int kDecodeWarps = 8;
constexpr int kMaxUsersUnrolled = 8;

template <int D, typename CodeT, typename LT, bool CB_SMEM>
__global__ void __launch_bounds__(kDecodeWarps * 32)
hsq_decode_reduce_kernel(const CodeT *__restrict__ codes, const LT *__restrict__ l,
                         const float *__restrict__ lbub, const float *__restrict__ norms_f32,
                         int64_t user_stride, int n_users, int64_t n_chunks,
                         const float *__restrict__ codebook, int K,
                         const int64_t *__restrict__ seg_start, int n_seg, float s, int n_bit,
                         int mean, int accumulate, float *__restrict__ out)
{
    extern __shared__ float4 s_cbd[];
    constexpr int D4 = D / 4;         // float4 per chunk
    constexpr int ROWS = D4;          // warp rows (32 float4 each) per 32 chunks
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const float4 *cb4g = reinterpret_cast<const float4 *>(codebook);
    if (CB_SMEM) {
        for (int i = threadIdx.x; i < K * D4; i += kDecodeWarps * 32) s_cbd[i] = __ldg(cb4g + i);
        __syncthreads();
    }
    const float inv_users = 1.0f;  // (division is done with __fdiv_rn below)
    (void)inv_users;
    const int64_t n_groups = (n_chunks + 31) / 32;
    for (int64_t g = (int64_t)blockIdx.x * kDecodeWarps + warp; g < n_groups;
         g += (int64_t)gridDim.x * kDecodeWarps) {
        const int64_t c = g * 32 + lane;
        const bool valid = c < n_chunks;
        int seg = 0;
        if (valid && n_bit != 32) seg = find_segment(seg_start, n_seg, c);
        float4 acc[ROWS];
#pragma unroll
        for (int i = 0; i < ROWS; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);

        for (int u = 0; u < n_users; ++u) {
            int code = 0;
            float norm = 0.0f;
            if (valid) {
                const char *cu = reinterpret_cast<const char *>(codes) + u * user_stride;
                code = (int)reinterpret_cast<const CodeT *>(cu)[c];
                if (n_bit == 32) {
                    const char *nu = reinterpret_cast<const char *>(norms_f32) + u * user_stride;
                    norm = reinterpret_cast<const float *>(nu)[c];
                } else {
                    const char *lu = reinterpret_cast<const char *>(l) + u * user_stride;
                    const char *bu = reinterpret_cast<const char *>(lbub) + u * user_stride;
                    int lv = (int)reinterpret_cast<const LT *>(lu)[c];
                    const float *b = reinterpret_cast<const float *>(bu);
                    norm = psc_value(lv, __ldg(b + 2 * seg), __ldg(b + 2 * seg + 1), s);
                }
            }
#pragma unroll
            for (int i = 0; i < ROWS; ++i) {
                const int f = i * 32 + lane;      // float4 index inside the 32-chunk group
                const int slot = f / D4;          // owning chunk (lane of phase A)
                const int part = f % D4;
                const int cd = __shfl_sync(0xffffffffu, code, slot);
                const float nm = __shfl_sync(0xffffffffu, norm, slot);
                float4 cw = CB_SMEM ? s_cbd[cd * D4 + part] : __ldg(cb4g + (int64_t)cd * D4 + part);
                float4 pr;
                pr.x = __fmul_rn(cw.x, nm); pr.y = __fmul_rn(cw.y, nm);
                pr.z = __fmul_rn(cw.z, nm); pr.w = __fmul_rn(cw.w, nm);
                if (u == 0) {
                    acc[i] = pr;
                } else {
                    acc[i].x = __fadd_rn(acc[i].x, pr.x); acc[i].y = __fadd_rn(acc[i].y, pr.y);
                    acc[i].z = __fadd_rn(acc[i].z, pr.z); acc[i].w = __fadd_rn(acc[i].w, pr.w);
                }
            }
        }
        const float nu = (float)n_users;
        float4 *o4 = reinterpret_cast<float4 *>(out) + g * 32 * D4;
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
            const int f = i * 32 + lane;
            const int64_t chunk = g * 32 + f / D4;
            if (chunk >= n_chunks) continue;
            float4 r = acc[i];
            if (mean) {
                r.x = __fdiv_rn(r.x, nu); r.y = __fdiv_rn(r.y, nu);
                r.z = __fdiv_rn(r.z, nu); r.w = __fdiv_rn(r.w, nu);
            }
            if (accumulate) {
                float4 o = o4[f];
                r.x = __fadd_rn(o.x, r.x); r.y = __fadd_rn(o.y, r.y);
                r.z = __fadd_rn(o.z, r.z); r.w = __fadd_rn(o.w, r.w);
            }
            o4[f] = r;
        }
    }
}

// scalar fallback for chunk dims that are not a multiple of 4
template <typename CodeT, typename LT>
__global__ void __launch_bounds__(256)
hsq_decode_reduce_generic_kernel(const CodeT *__restrict__ codes, const LT *__restrict__ l,
                                 const float *__restrict__ lbub, const float *__restrict__ norms_f32,
                                 int64_t user_stride, int n_users, int64_t n_chunks, int d,
                                 const float *__restrict__ codebook,
                                 const int64_t *__restrict__ seg_start, int n_seg, float s, int n_bit,
                                 int mean, int accumulate, float *__restrict__ out)
{
    const int64_t n = n_chunks * (int64_t)d;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256) {
        const int64_t c = e / d;
        const int j = (int)(e % d);
        int seg = (n_bit != 32) ? find_segment(seg_start, n_seg, c) : 0;
        float acc = 0.0f;
        for (int u = 0; u < n_users; ++u) {
            const char *cu = reinterpret_cast<const char *>(codes) + u * user_stride;
            int code = (int)reinterpret_cast<const CodeT *>(cu)[c];
            float norm;
            if (n_bit == 32) {
                const char *nu = reinterpret_cast<const char *>(norms_f32) + u * user_stride;
                norm = reinterpret_cast<const float *>(nu)[c];
            } else {
                const char *lu = reinterpret_cast<const char *>(l) + u * user_stride;
                const char *bu = reinterpret_cast<const char *>(lbub) + u * user_stride;
                const float *b = reinterpret_cast<const float *>(bu);
                norm = psc_value((int)reinterpret_cast<const LT *>(lu)[c], __ldg(b + 2 * seg),
                                 __ldg(b + 2 * seg + 1), s);
            }
            float pr = __fmul_rn(__ldg(codebook + (int64_t)code * d + j), norm);
            acc = (u == 0) ? pr : __fadd_rn(acc, pr);
        }
        if (mean) acc = __fdiv_rn(acc, (float)n_users);
        if (accumulate) acc = __fadd_rn(out[e], acc);
        out[e] = acc;
    }
}

template <int D, typename CodeT, typename LT>
static int launch_decode_d(const void *codes, const void *l, const float *lbub, const float *norms_f32,
                           int64_t user_stride, int n_users, int64_t n_chunks, const float *codebook,
                           int K, const int64_t *seg_start, int n_seg, int n_bit, int mean,
                           int accumulate, float *out, cudaStream_t st)
{
    const float s = (n_bit == 32) ? 1.0f : (float)(1u << n_bit);
    const size_t cb_bytes = (size_t)K * D * 4;
    const int64_t n_groups = (n_chunks + 31) / 32;
    int64_t blocks = (n_groups + kDecodeWarps - 1) / kDecodeWarps;
    if (cb_bytes <= 64 * 1024) {
        auto kern = hsq_decode_reduce_kernel<D, CodeT, LT, true>;
        GQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cb_bytes));
        // persistent: the codebook is staged once per CTA
        int64_t cap = (int64_t)sm_count() * (cb_bytes <= 16 * 1024 ? 8 : 3);
        int grid = (int)(blocks < cap ? blocks : cap);
        kern<<<grid < 1 ? 1 : grid, kDecodeWarps * 32, cb_bytes, st>>>(
            (const CodeT *)codes, (const LT *)l, lbub, norms_f32, user_stride, n_users, n_chunks,
            codebook, K, seg_start, n_seg, s, n_bit, mean, accumulate, out);
    } else {
        auto kern = hsq_decode_reduce_kernel<D, CodeT, LT, false>;
        int64_t cap = (int64_t)sm_count() * 8;
        int grid = (int)(blocks < cap ? blocks : cap);
        kern<<<grid < 1 ? 1 : grid, kDecodeWarps * 32, 0, st>>>(
            (const CodeT *)codes, (const LT *)l, lbub, norms_f32, user_stride, n_users, n_chunks,
            codebook, K, seg_start, n_seg, s, n_bit, mean, accumulate, out);
    }
    GQ_LAUNCH_CHECK("hsq_decode_reduce");
    return GQ_OK;
}

template <typename CodeT, typename LT>
static int launch_decode(const void *codes, const void *l, const float *lbub, const float *norms_f32,
                         int64_t user_stride, int n_users, int64_t n_chunks, int d,
                         const float *codebook, int K, const int64_t *seg_start, int n_seg, int n_bit,
                         int mean, int accumulate, float *out, cudaStream_t st)
{
    switch (d) {
#define GQ_CASE(DD) case DD: return launch_decode_d<DD, CodeT, LT>(codes, l, lbub, norms_f32, user_stride, n_users, n_chunks, codebook, K, seg_start, n_seg, n_bit, mean, accumulate, out, st);
        GQ_CASE(4) GQ_CASE(8) GQ_CASE(12) GQ_CASE(16) GQ_CASE(24) GQ_CASE(32) GQ_CASE(48) GQ_CASE(64)
#undef GQ_CASE
        default: break;
    }
    const float s = (n_bit == 32) ? 1.0f : (float)(1u << n_bit);
    hsq_decode_reduce_generic_kernel<CodeT, LT><<<grid_for(n_chunks * (int64_t)d, 256), 256, 0, st>>>(
        (const CodeT *)codes, (const LT *)l, lbub, norms_f32, user_stride, n_users, n_chunks, d,
        codebook, seg_start, n_seg, s, n_bit, mean, accumulate, out);
    GQ_LAUNCH_CHECK("hsq_decode_reduce_generic");
    return GQ_OK;
}

int hsq_decode_reduce(const void *codes, int code_bytes, const void *l, int l_bytes, const float *lbub,
                      const float *norms_f32, int64_t user_stride, int n_users, int64_t n_chunks, int d,
                      const float *codebook, int K, const int64_t *seg_start, int n_seg, int n_bit,
                      int mean, int accumulate, float *out, cudaStream_t st)
{
    if (n_chunks == 0) return GQ_OK;
#define GQ_GO(CT, LTT) return launch_decode<CT, LTT>(codes, l, lbub, norms_f32, user_stride, n_users, n_chunks, d, codebook, K, seg_start, n_seg, n_bit, mean, accumulate, out, st)
    if (code_bytes == 1 && l_bytes == 1) GQ_GO(uint8_t, uint8_t);
    if (code_bytes == 1 && l_bytes == 4) GQ_GO(uint8_t, int32_t);
    if (code_bytes == 4 && l_bytes == 1) GQ_GO(int32_t, uint8_t);
    GQ_GO(int32_t, int32_t);
#undef GQ_GO
}

// ------------------------------------------------------ fp32 user reduction ---
__global__ void __launch_bounds__(256)
f32_reduce_users_kernel(const float *__restrict__ in, int64_t user_stride, int n_users, int64_t n,
                        int mean, int accumulate, float *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        float acc = in[i];
        for (int u = 1; u < n_users; ++u)
            acc = __fadd_rn(acc, *reinterpret_cast<const float *>(
                                     reinterpret_cast<const char *>(in) + u * user_stride + 4 * i));
        if (mean) acc = __fdiv_rn(acc, (float)n_users);
        if (accumulate) acc = __fadd_rn(out[i], acc);
        out[i] = acc;
    }
}

int launch_f32_reduce_users(const float *in, int64_t user_stride, int n_users, int64_t n, int mean,
                            int accumulate, float *out, cudaStream_t st)
{
    if (n == 0) return GQ_OK;
    f32_reduce_users_kernel<<<grid_for(n, 256), 256, 0, st>>>(in, user_stride, n_users, n, mean,
                                                              accumulate, out);
    GQ_LAUNCH_CHECK("f32_reduce_users");
    return GQ_OK;
}

// ------------------------------------------------------------- elementwise ---
__global__ void __launch_bounds__(256)
axpy_kernel(const float *__restrict__ a, const float *__restrict__ b, float alpha, int64_t n,
            float *__restrict__ out, int sub)
{
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        out[i] = sub ? __fsub_rn(a[i], b[i]) : __fadd_rn(a[i], __fmul_rn(alpha, b[i]));
    }
}

int launch_axpy(const float *a, const float *b, float alpha, int64_t n, float *out, int sub,
                cudaStream_t st)
{
    if (n == 0) return GQ_OK;
    axpy_kernel<<<grid_for(n, 256), 256, 0, st>>>(a, b, alpha, n, out, sub);
    GQ_LAUNCH_CHECK("axpy");
    return GQ_OK;
}

}  // namespace gq
