// gather.cu -- multi-tensor gather: the per-parameter gradients autograd leaves in param.grad
// (main.py:229-230) are copied into the codec arena by ONE kernel per 128 tensors, driven by a
// pointer table passed by value (no device-side table to build, no per-tensor launches, no
// framework op on the drop-in path).  HBM-bound: 8 bytes per element.
// Error feedback rides along: with feedback != 0 the destination already holds the user's error
// state e and becomes g + scale * e (ps_quantizer.py:35, ring_quantizer.py:34: grad += scale * error,
// same two roundings), optionally written back to the source tensors as well (the reference
// mutates param.grad in place) -- no separate axpy sweep over the model.
#include "gq_internal.cuh"

namespace gq {

constexpr int kGatherMax = 128;      // tensors per launch (the table travels as a kernel parameter)
constexpr int kGatherTile = 4096;    // elements per block

struct GatherTable {
    const float *src[kGatherMax];
    int64_t dst_off[kGatherMax];     // element offset in dst
    int64_t size[kGatherMax];
    int tile_prefix[kGatherMax + 1]; // blocks before tensor t
    int n;
    int feedback;                    // 0: dst = src; 1: dst = src + scale * dst; 2: same, and src = that sum too
    float scale;
};

__global__ void __launch_bounds__(256)
gather_f32_kernel(const __grid_constant__ GatherTable T, float *__restrict__ dst)
{
    int lo = 0, hi = T.n;            // tile_prefix[lo] <= blockIdx.x < tile_prefix[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (T.tile_prefix[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
    }
    const float *s = T.src[lo];
    float *d = dst + T.dst_off[lo];
    const int64_t n = T.size[lo];
    const int64_t begin = (int64_t)((int)blockIdx.x - T.tile_prefix[lo]) * kGatherTile;
    const int64_t end = min(begin + kGatherTile, n);
    if (T.feedback) {
        const float scale = T.scale;
        const bool back = T.feedback == 2;
        if ((((uintptr_t)s | (uintptr_t)d) & 15) == 0) {
            const int64_t v_end = begin + ((end - begin) & ~(int64_t)3);
            for (int64_t i = begin + 4 * threadIdx.x; i < v_end; i += 4 * 256) {
                const float4 a = *reinterpret_cast<const float4 *>(s + i);
                const float4 e = *reinterpret_cast<const float4 *>(d + i);
                const float4 r = make_float4(__fadd_rn(a.x, __fmul_rn(scale, e.x)), __fadd_rn(a.y, __fmul_rn(scale, e.y)),
                                             __fadd_rn(a.z, __fmul_rn(scale, e.z)), __fadd_rn(a.w, __fmul_rn(scale, e.w)));
                *reinterpret_cast<float4 *>(d + i) = r;
                if (back) *reinterpret_cast<float4 *>(const_cast<float *>(s) + i) = r;
            }
            for (int64_t i = v_end + threadIdx.x; i < end; i += 256) {
                const float r = __fadd_rn(s[i], __fmul_rn(scale, d[i]));
                d[i] = r;
                if (back) const_cast<float *>(s)[i] = r;
            }
        } else {
            for (int64_t i = begin + threadIdx.x; i < end; i += 256) {
                const float r = __fadd_rn(s[i], __fmul_rn(scale, d[i]));
                d[i] = r;
                if (back) const_cast<float *>(s)[i] = r;
            }
        }
        return;
    }
    if ((((uintptr_t)s | (uintptr_t)d) & 15) == 0) {
        const int64_t v_end = begin + ((end - begin) & ~(int64_t)3);
        for (int64_t i = begin + 4 * threadIdx.x; i < v_end; i += 4 * 256)
            *reinterpret_cast<float4 *>(d + i) = ld_stream_f4(reinterpret_cast<const float4 *>(s + i));
        for (int64_t i = v_end + threadIdx.x; i < end; i += 256) d[i] = s[i];
    } else {
        for (int64_t i = begin + threadIdx.x; i < end; i += 256) d[i] = s[i];
    }
}

}  // namespace gq

using namespace gq;

extern "C" int gq_gather_f32(const void *const *src_ptrs, const int64_t *dst_offsets, const int64_t *sizes,
                             int n_tensors, float *dst, int feedback, float scale, gq_stream_t stream)
{
    GQ_REQUIRE(feedback >= 0 && feedback <= 2, "feedback must be 0, 1 or 2");
    GQ_REQUIRE(n_tensors >= 0 && (n_tensors == 0 || (src_ptrs && dst_offsets && sizes && dst)), "bad arguments");
    cudaStream_t st = as_stream(stream);
    for (int t0 = 0; t0 < n_tensors; t0 += kGatherMax) {
        GatherTable T = {};
        int blocks = 0, m = 0;
        for (int t = t0; t < n_tensors && m < kGatherMax; ++t) {
            GQ_REQUIRE(sizes[t] >= 0 && (sizes[t] == 0 || src_ptrs[t]), "tensor %d: null source", t);
            GQ_REQUIRE(((uintptr_t)src_ptrs[t] & 3) == 0, "tensor %d: source is not 4-byte aligned", t);
            if (sizes[t] == 0) continue;
            T.src[m] = reinterpret_cast<const float *>(src_ptrs[t]);
            T.dst_off[m] = dst_offsets[t];
            T.size[m] = sizes[t];
            T.tile_prefix[m] = blocks;
            blocks += (int)((sizes[t] + kGatherTile - 1) / kGatherTile);
            ++m;
        }
        T.tile_prefix[m] = blocks;
        T.n = m;
        T.feedback = feedback;
        T.scale = scale;
        if (blocks == 0) continue;
        gather_f32_kernel<<<blocks, 256, 0, st>>>(T, dst);
        GQ_LAUNCH_CHECK("gather_f32");
    }
    return GQ_OK;
}
