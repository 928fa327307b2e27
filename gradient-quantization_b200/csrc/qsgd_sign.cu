// qsgd_sign.cu -- QSGD / TernGrad and SignSGD: HBM-bound elementwise codecs.
//   QSGD   : compressors/qsgd_compressor.py:42-71
//   SignSGD: compressors/signsgd_compressor.py:8-12
// Algorithmic bytes per gradient element: QSGD encode 4 (+4 when uniforms are
// supplied) + bits/8 + 4/dim; QSGD decode-reduce 4 + U*(bits/8 + 4/dim);
// sign encode 4 + 0.25; sign decode-reduce 4 + 0.25 U.
#include "gq_internal.cuh"

namespace gq {

static int grid_for(int64_t n, int per_block, int waves = 8)
{
    int64_t blocks = (n + per_block - 1) / per_block;
    int64_t cap = (int64_t)sm_count() * waves;
    int64_t g = blocks < cap ? blocks : cap;
    return (int)(g < 1 ? 1 : g);
}

__device__ __forceinline__ int64_t chunk_of(const int64_t *__restrict__ chunk_start, int64_t n_chunks,
                                            int dim, int64_t i)
{
    return chunk_start ? (int64_t)find_segment(chunk_start, (int)n_chunks, i) : i / dim;
}

// ------------------------------------------------------------ chunk L-inf ---
// norm[m] = max |v| over chunk m (qsgd_compressor.py:49).  |v| bit patterns are
// monotone as unsigned ints, so one atomicMax per warp (or lane) suffices.
__global__ void __launch_bounds__(256)
chunk_absmax_kernel(const float *__restrict__ v, int64_t n, const int64_t *__restrict__ chunk_start,
                    int64_t n_chunks, int dim, uint32_t *__restrict__ norm_bits)
{
    for (int64_t base = (int64_t)blockIdx.x * 256; base < n; base += (int64_t)gridDim.x * 256) {
        int64_t i = base + threadIdx.x;
        bool valid = i < n;
        uint32_t a = valid ? (__float_as_uint(v[i]) & 0x7fffffffu) : 0u;
        int64_t m = valid ? chunk_of(chunk_start, n_chunks, dim, i) : -1;
        int64_t m0 = __shfl_sync(0xffffffffu, m, 0);
        bool uniform = __all_sync(0xffffffffu, (m == m0) || !valid) && (m0 >= 0);
        if (uniform) {
            uint32_t w = __reduce_max_sync(0xffffffffu, a);
            if ((threadIdx.x & 31) == 0) atomicMax(norm_bits + m0, w);
        } else if (valid) {
            atomicMax(norm_bits + m, a);
        }
    }
}

// --------------------------------------------------------------- quantize ---
// level/sign of one element (qsgd_compressor.py:50-63), exact op order.
// returns level; nan_level flags the 0/0 case (reference: int cast of NaN = INT_MIN).
__device__ __forceinline__ int qsgd_level(float x, float nm, float s, int random, float r, bool &is_nan)
{
    float scaled = fabsf(__fdiv_rn(x, nm)) * s;
    is_nan = (scaled != scaled);
    if (is_nan) return 0;
    float c = fminf(fmaxf(scaled, 0.0f), s - 1.0f);
    int li = (int)c;
    if (random) {
        float prob = __fsub_rn(scaled, (float)li);
        li += (prob > r) ? 1 : 0;
    }
    return li;
}

template <int BITS>  // 0: no packed output; 4, 8, 16
__global__ void __launch_bounds__(256)
qsgd_quantize_kernel(const float *__restrict__ v, int64_t n, const int64_t *__restrict__ chunk_start,
                     int64_t n_chunks, int dim, float s, int random,
                     const float *__restrict__ uniforms, uint64_t seed, uint64_t offset,
                     const float *__restrict__ norm, uint8_t *__restrict__ signs,
                     int32_t *__restrict__ l, void *__restrict__ packed)
{
    const int64_t n4 = (n + 3) / 4;
    for (int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x; q < n4; q += (int64_t)gridDim.x * 256) {
        const int64_t i0 = q * 4;
        float x[4], r[4];
        const bool full = (i0 + 3 < n);
        if (full) {
            float4 t = ld_stream_f4(reinterpret_cast<const float4 *>(v) + q);
            x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
        } else {
#pragma unroll
            for (int t = 0; t < 4; ++t) x[t] = (i0 + t < n) ? v[i0 + t] : 0.0f;
        }
        if (random) {
            if (uniforms) {
                if (full) {
                    float4 t = ld_stream_f4(reinterpret_cast<const float4 *>(uniforms) + q);
                    r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
                } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t) r[t] = (i0 + t < n) ? uniforms[i0 + t] : 0.0f;
                }
            } else if (((offset + (uint64_t)i0) & 3u) == 0) {
                uint4 w = philox4x32_10(seed, (offset + (uint64_t)i0) >> 2);
                r[0] = u01(w.x); r[1] = u01(w.y); r[2] = u01(w.z); r[3] = u01(w.w);
            } else {
#pragma unroll
                for (int t = 0; t < 4; ++t) r[t] = philox_uniform(seed, offset, (uint64_t)(i0 + t));
            }
        } else {
            r[0] = r[1] = r[2] = r[3] = 0.0f;
        }
        uint32_t pk[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int64_t i = i0 + t;
            if (i >= n) { pk[t] = 0; continue; }
            const int64_t m = chunk_of(chunk_start, n_chunks, dim, i);
            const float nm = __ldg(norm + m);
            bool is_nan;
            int li = qsgd_level(x[t], nm, s, random, r[t], is_nan);
            const uint32_t sg = (x[t] > 0.0f) ? 1u : 0u;
            if (signs) signs[i] = (uint8_t)sg;
            if (l) l[i] = is_nan ? (int32_t)0x80000000 : li;
            pk[t] = (BITS > 0) ? ((sg << (BITS - 1)) | (uint32_t)li) : 0u;
        }
        if (BITS == 4) {
            reinterpret_cast<uint16_t *>(packed)[q] =
                (uint16_t)(pk[0] | (pk[1] << 4) | (pk[2] << 8) | (pk[3] << 12));
        } else if (BITS == 8) {
            reinterpret_cast<uint32_t *>(packed)[q] = pk[0] | (pk[1] << 8) | (pk[2] << 16) | (pk[3] << 24);
        } else if (BITS == 16) {
            reinterpret_cast<uint2 *>(packed)[q] = make_uint2(pk[0] | (pk[1] << 16), pk[2] | (pk[3] << 16));
        }
    }
}

int qsgd_wire_bits(int n_bit) { return n_bit <= 2 ? 4 : (n_bit <= 6 ? 8 : 16); }

int qsgd_encode(const float *grad, int64_t n, const int64_t *chunk_start, int64_t n_chunks, int dim,
                int n_bit, int random, const float *uniforms, uint64_t seed, uint64_t offset, float *norm,
                uint8_t *signs, int32_t *l, void *packed, cudaStream_t st)
{
    if (n == 0) return GQ_OK;
    GQ_CUDA(cudaMemsetAsync(norm, 0, (size_t)n_chunks * 4, st));
    chunk_absmax_kernel<<<grid_for(n, 256), 256, 0, st>>>(grad, n, chunk_start, n_chunks, dim,
                                                          reinterpret_cast<uint32_t *>(norm));
    GQ_LAUNCH_CHECK("chunk_absmax");
    const float s = (float)(1u << n_bit);
    const int grid = grid_for((n + 3) / 4, 256);
    const int bits = packed ? qsgd_wire_bits(n_bit) : 0;
#define GQ_Q(B) qsgd_quantize_kernel<B><<<grid, 256, 0, st>>>(grad, n, chunk_start, n_chunks, dim, s, random, uniforms, seed, offset, norm, signs, l, packed)
    if (bits == 0) GQ_Q(0);
    else if (bits == 4) GQ_Q(4);
    else if (bits == 8) GQ_Q(8);
    else GQ_Q(16);
#undef GQ_Q
    GQ_LAUNCH_CHECK("qsgd_quantize");
    return GQ_OK;
}

// ----------------------------------------------------------------- decode ---
// value = (float(l) * (2*sign - 1)) * norm / s   (qsgd_compressor.py:69-70)
__device__ __forceinline__ float qsgd_value(int li, uint32_t sg, float nm, float s)
{
    float sv = __fmul_rn((float)li, __fsub_rn(__fmul_rn(2.0f, (float)sg), 1.0f));
    return __fdiv_rn(__fmul_rn(sv, nm), s);
}

template <int BITS>
__global__ void __launch_bounds__(256)
qsgd_decode_reduce_kernel(const float *__restrict__ norm, const void *__restrict__ packed,
                          int64_t user_stride, int n_users, int64_t n,
                          const int64_t *__restrict__ chunk_start, int64_t n_chunks, int dim, float s,
                          int mean, int accumulate, float *__restrict__ out)
{
    const int64_t n4 = (n + 3) / 4;
    for (int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x; q < n4; q += (int64_t)gridDim.x * 256) {
        const int64_t i0 = q * 4;
        int64_t m[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) m[t] = (i0 + t < n) ? chunk_of(chunk_start, n_chunks, dim, i0 + t) : 0;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int u = 0; u < n_users; ++u) {
            const char *pu = reinterpret_cast<const char *>(packed) + u * user_stride;
            const float *nu = reinterpret_cast<const float *>(reinterpret_cast<const char *>(norm) + u * user_stride);
            uint32_t pk[4];
            if (BITS == 4) {
                uint32_t w = reinterpret_cast<const uint16_t *>(pu)[q];
                pk[0] = w & 15u; pk[1] = (w >> 4) & 15u; pk[2] = (w >> 8) & 15u; pk[3] = (w >> 12) & 15u;
            } else if (BITS == 8) {
                uint32_t w = reinterpret_cast<const uint32_t *>(pu)[q];
                pk[0] = w & 255u; pk[1] = (w >> 8) & 255u; pk[2] = (w >> 16) & 255u; pk[3] = w >> 24;
            } else {
                uint2 w = reinterpret_cast<const uint2 *>(pu)[q];
                pk[0] = w.x & 0xffffu; pk[1] = w.x >> 16; pk[2] = w.y & 0xffffu; pk[3] = w.y >> 16;
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const uint32_t sg = pk[t] >> (BITS - 1);
                const int li = (int)(pk[t] & ((1u << (BITS - 1)) - 1u));
                float val = qsgd_value(li, sg, __ldg(nu + m[t]), s);
                acc[t] = (u == 0) ? val : __fadd_rn(acc[t], val);
            }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (i0 + t >= n) continue;
            float r = acc[t];
            if (mean) r = __fdiv_rn(r, (float)n_users);
            if (accumulate) r = (accumulate == 2) ? __fsub_rn(out[i0 + t], r) : __fadd_rn(out[i0 + t], r);
            out[i0 + t] = r;
        }
    }
}

int qsgd_decode_reduce(const float *norm, const void *packed, int64_t user_stride, int n_users,
                       int64_t n, const int64_t *chunk_start, int64_t n_chunks, int dim, int n_bit,
                       int mean, int accumulate, float *out, cudaStream_t st)
{
    if (n == 0) return GQ_OK;
    const float s = (float)(1u << n_bit);
    const int grid = grid_for((n + 3) / 4, 256);
    const int bits = qsgd_wire_bits(n_bit);
#define GQ_D(B) qsgd_decode_reduce_kernel<B><<<grid, 256, 0, st>>>(norm, packed, user_stride, n_users, n, chunk_start, n_chunks, dim, s, mean, accumulate, out)
    if (bits == 4) GQ_D(4);
    else if (bits == 8) GQ_D(8);
    else GQ_D(16);
#undef GQ_D
    GQ_LAUNCH_CHECK("qsgd_decode_reduce");
    return GQ_OK;
}

__global__ void __launch_bounds__(256)
qsgd_decode_unpacked_kernel(const float *__restrict__ norm, const uint8_t *__restrict__ signs,
                            const int32_t *__restrict__ l, int64_t n,
                            const int64_t *__restrict__ chunk_start, int64_t n_chunks, int dim, float s,
                            float *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const int64_t m = chunk_of(chunk_start, n_chunks, dim, i);
        out[i] = qsgd_value(l[i], (uint32_t)signs[i], __ldg(norm + m), s);
    }
}

int qsgd_decode_unpacked(const float *norm, const uint8_t *signs, const int32_t *l, int64_t n,
                         const int64_t *chunk_start, int64_t n_chunks, int dim, int n_bit, float *out,
                         cudaStream_t st)
{
    if (n == 0) return GQ_OK;
    qsgd_decode_unpacked_kernel<<<grid_for(n, 256), 256, 0, st>>>(norm, signs, l, n, chunk_start, n_chunks,
                                                                  dim, (float)(1u << n_bit), out);
    GQ_LAUNCH_CHECK("qsgd_decode_unpacked");
    return GQ_OK;
}

// ------------------------------------------------------------------- sign ---
// torch.sign -> {-1, 0, +1}; packed: 2 bits per element (0 -> 0, 1 -> +1, 2 -> -1).
__global__ void __launch_bounds__(256)
sign_encode_kernel(const float *__restrict__ v, int64_t n, float *__restrict__ out_f32,
                   uint8_t *__restrict__ packed)
{
    const int64_t n4 = (n + 3) / 4;
    for (int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x; q < n4; q += (int64_t)gridDim.x * 256) {
        const int64_t i0 = q * 4;
        float x[4];
        if (i0 + 3 < n) {
            float4 t = ld_stream_f4(reinterpret_cast<const float4 *>(v) + q);
            x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
        } else {
#pragma unroll
            for (int t = 0; t < 4; ++t) x[t] = (i0 + t < n) ? v[i0 + t] : 0.0f;
        }
        uint32_t byte = 0;
        float sg[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int pos = x[t] > 0.0f, neg = x[t] < 0.0f;
            sg[t] = (float)(pos - neg);
            byte |= (uint32_t)(pos | (neg << 1)) << (2 * t);
        }
        if (packed) packed[q] = (uint8_t)byte;
        if (out_f32) {
            if (i0 + 3 < n) {
                reinterpret_cast<float4 *>(out_f32)[q] = make_float4(sg[0], sg[1], sg[2], sg[3]);
            } else {
#pragma unroll
                for (int t = 0; t < 4; ++t) if (i0 + t < n) out_f32[i0 + t] = sg[t];
            }
        }
    }
}

__global__ void __launch_bounds__(256)
sign_decode_reduce_kernel(const uint8_t *__restrict__ packed, int64_t user_stride, int n_users,
                          int64_t n, int mean, int accumulate, float *__restrict__ out)
{
    const int64_t n4 = (n + 3) / 4;
    for (int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x; q < n4; q += (int64_t)gridDim.x * 256) {
        const int64_t i0 = q * 4;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int u = 0; u < n_users; ++u) {
            const uint32_t b = packed[u * user_stride + q];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const uint32_t c = (b >> (2 * t)) & 3u;
                const float val = (float)((int)(c & 1u) - (int)(c >> 1));
                acc[t] = (u == 0) ? val : __fadd_rn(acc[t], val);
            }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (i0 + t >= n) continue;
            float r = acc[t];
            if (mean) r = __fdiv_rn(r, (float)n_users);
            if (accumulate) r = (accumulate == 2) ? __fsub_rn(out[i0 + t], r) : __fadd_rn(out[i0 + t], r);
            out[i0 + t] = r;
        }
    }
}

int sign_encode(const float *grad, int64_t n, float *out_f32, uint8_t *packed, cudaStream_t st)
{
    if (n == 0) return GQ_OK;
    sign_encode_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, st>>>(grad, n, out_f32, packed);
    GQ_LAUNCH_CHECK("sign_encode");
    return GQ_OK;
}

int sign_decode_reduce(const uint8_t *packed, int64_t user_stride, int n_users, int64_t n, int mean,
                       int accumulate, float *out, cudaStream_t st)
{
    if (n == 0) return GQ_OK;
    sign_decode_reduce_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, st>>>(packed, user_stride, n_users, n,
                                                                          mean, accumulate, out);
    GQ_LAUNCH_CHECK("sign_decode_reduce");
    return GQ_OK;
}

}  // namespace gq

using namespace gq;

extern "C" {

int gq_qsgd_wire_bits(int n_bit) { return qsgd_wire_bits(n_bit); }

int gq_qsgd_encode(const float *grad, int64_t n, const int64_t *chunk_start, int64_t n_chunks, int dim,
                   int n_bit, int random, const float *uniforms, uint64_t philox_seed,
                   uint64_t philox_offset, float *norm, uint8_t *signs, int32_t *l, void *packed,
                   gq_stream_t stream)
{
    GQ_REQUIRE(n >= 0 && n_chunks >= 0, "negative size");
    GQ_REQUIRE(n_bit >= 1 && n_bit <= 14, "n_bit %d out of range 1..14", n_bit);
    GQ_REQUIRE(chunk_start || (dim >= 1 && n_chunks * (int64_t)dim == n),
               "n (%lld) != n_chunks (%lld) * dim (%d)", (long long)n, (long long)n_chunks, dim);
    GQ_REQUIRE(n == 0 || (grad && norm), "null pointer");
    GQ_REQUIRE(packed || (signs && l), "need packed or (signs and l) outputs");
    GQ_REQUIRE(((uintptr_t)grad & 15) == 0, "gradient must be 16-byte aligned");
    return qsgd_encode(grad, n, chunk_start, n_chunks, dim, n_bit, random, uniforms, philox_seed,
                       philox_offset, norm, signs, l, packed, as_stream(stream));
}

int gq_qsgd_decode_reduce(const float *norm, const void *packed, int64_t user_stride_bytes, int n_users,
                          int64_t n, const int64_t *chunk_start, int64_t n_chunks, int dim, int n_bit,
                          int mean, int accumulate, float *out, gq_stream_t stream)
{
    GQ_REQUIRE(n >= 0 && n_users >= 1, "bad sizes");
    GQ_REQUIRE(n_bit >= 1 && n_bit <= 14, "n_bit %d out of range 1..14", n_bit);
    GQ_REQUIRE(chunk_start || (dim >= 1 && n_chunks * (int64_t)dim == n), "n != n_chunks * dim");
    GQ_REQUIRE(n == 0 || (norm && packed && out), "null pointer");
    return qsgd_decode_reduce(norm, packed, user_stride_bytes, n_users, n, chunk_start, n_chunks, dim,
                              n_bit, mean, accumulate, out, as_stream(stream));
}

int gq_qsgd_decode_unpacked(const float *norm, const uint8_t *signs, const int32_t *l, int64_t n,
                            const int64_t *chunk_start, int64_t n_chunks, int dim, int n_bit, float *out,
                            gq_stream_t stream)
{
    GQ_REQUIRE(n >= 0, "bad sizes");
    GQ_REQUIRE(chunk_start || (dim >= 1 && n_chunks * (int64_t)dim == n), "n != n_chunks * dim");
    GQ_REQUIRE(n == 0 || (norm && signs && l && out), "null pointer");
    return qsgd_decode_unpacked(norm, signs, l, n, chunk_start, n_chunks, dim, n_bit, out,
                                as_stream(stream));
}

int gq_sign_encode(const float *grad, int64_t n, float *out_f32, uint8_t *packed, gq_stream_t stream)
{
    GQ_REQUIRE(n >= 0 && (n == 0 || grad), "bad arguments");
    GQ_REQUIRE(out_f32 || packed, "need at least one output");
    GQ_REQUIRE(((uintptr_t)grad & 15) == 0, "gradient must be 16-byte aligned");
    return sign_encode(grad, n, out_f32, packed, as_stream(stream));
}

int gq_sign_decode_reduce(const uint8_t *packed, int64_t user_stride_bytes, int n_users, int64_t n,
                          int mean, int accumulate, float *out, gq_stream_t stream)
{
    GQ_REQUIRE(n >= 0 && n_users >= 1 && (n == 0 || (packed && out)), "bad arguments");
    return sign_decode_reduce(packed, user_stride_bytes, n_users, n, mean, accumulate, out,
                              as_stream(stream));
}

}  // extern "C"
