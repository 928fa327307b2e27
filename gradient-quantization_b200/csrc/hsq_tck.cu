// hsq_tck.cu -- tcgen05 (TF32) nearest-codeword search for LARGE codebooks: d == 16, K = 512 .. 4096
// (a multiple of 256; BASELINE config 4's K = 2^12), int32 codes, exact fp32 rescoring.
//
// The [128 x 16] x [16 x K] contraction of nearest_neighbor_compressor.py:68 runs as T = K / 256
// N-tiles per 128-chunk row tile: the gradient tile (8 KB, TMA, SWIZZLE_64B) stays in shared memory
// while the TF32-rounded codebook streams through a 4-deep ring of 16 KB tiles (TMA from an
// L2-resident copy prepared by hsq_tck_prep_kernel), 2 x tcgen05.mma (M128 N256 K8) per N-tile into
// one of two 256-column TMEM buffers.  Two epilogue groups of four warps, one per TMEM buffer, one
// thread per row: each N-tile's 256 approximate scores are reduced to 16 "coarse" maxima (one per
// 16 codewords) that go to shared memory (K / 16 floats per row), plus a running row maximum.  After
// the last N-tile of a row tile the two groups exchange their maxima, and every coarse group whose
// maximum lies within 2 * eps of the row maximum is rescored exactly (ascending-j fmaf chain over the
// fp32 codebook, read through L1/L2; first index wins) -- the same exactness argument as hsq_tc.cu.
// A dedicated warp issues the MMAs (it waits, with the low-latency mbarrier wait, for the four warps
// of the buffer's epilogue group to release it; letting the last of them issue the next MMA itself was
// the first design and left the buffer idle while that warp also waited for the operand tiles).
// Bound: tensor pipe (2 * K flop per element: 32 MMAs per row tile at K = 4096) with the epilogue's
// 16 first passes per row tile close behind.
#include <cuda.h>
#include <stdlib.h>

#include <mutex>

#include "gq_internal.cuh"
#include "tc_ptx.cuh"

namespace gq {
namespace tck {

using namespace tcptx;

constexpr int kD = 16;
constexpr int kTileM = 128;
constexpr int kTileN = 256;
constexpr int kCoarse = 16;                        // codewords per coarse group
constexpr int kMaxK = 4096;
constexpr int kBStages = 4;
constexpr uint32_t kATileBytes = kTileM * kD * 4;  // 8192
constexpr uint32_t kBTileBytes = kTileN * kD * 4;  // 16384
constexpr uint32_t kOffA = 0;
constexpr uint32_t kOffB = 2 * kATileBytes;
constexpr uint32_t kOffCm = kOffB + kBStages * kBTileBytes;            // coarse maxima: [K/64][128 rows] float4
constexpr float kMargin = 2.0f * (1.5f / 1024.0f + 4.0e-6f);           // see hsq_tc.cu
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((kTileN >> 3) << 17) | ((kTileM >> 4) << 24);
constexpr int kThreads = 128 + 256;                // 4 control warps + 2 epilogue groups of 4 warps

__host__ __device__ inline uint32_t cm_bytes(int K) { return (uint32_t)(K / 64) * kTileM * 16u; }
__host__ __device__ inline uint32_t off_misc(int K) { return kOffCm + cm_bytes(K); }
// misc region: mbarriers (afull[2] aempty[2] bfull[4] bempty[4] tfull[2] tempty[2]) | tmem ptr | s_amax[2][128] |
//              s_best[128] x (bits, k, u)
constexpr uint32_t kMiscBytes = 16 * 8 + 16 + 16 + 2 * 128 * 4 + 128 * 12;
__host__ __device__ inline uint32_t smem_bytes(int K) { return off_misc(K) + kMiscBytes + 1024; }

// TF32-rounded copy of the codebook (row-major [K, 16]) and the largest codeword norm
__global__ void __launch_bounds__(256)
hsq_tck_prep_kernel(const float *__restrict__ codebook, int K, float *__restrict__ cb_tf32, uint32_t *__restrict__ cn_key)
{
    const int k = blockIdx.x * 256 + threadIdx.x;
    float c2 = 0.0f;
    if (k < K) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(codebook) + k * 4 + u);
            c2 = fmaf(v.x, v.x, c2); c2 = fmaf(v.y, v.y, c2); c2 = fmaf(v.z, v.z, c2); c2 = fmaf(v.w, v.w, c2);
            uint4 t;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t.x) : "f"(v.x));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t.y) : "f"(v.y));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t.z) : "f"(v.z));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t.w) : "f"(v.w));
            reinterpret_cast<uint4 *>(cb_tf32)[k * 4 + u] = t;
        }
    }
    float n = sqrtf(c2);
    if (!(n < 3.0e38f)) n = __int_as_float(0x7f800000);
    n = warp_max(n);
    if ((threadIdx.x & 31) == 0) atomicMax(cn_key, __float_as_uint(n));   // non-negative floats order like uints
}

// exact fp32 score of codeword k (global memory, L1/L2-resident), same chain as hsq_exact.cu
__device__ __forceinline__ float exact_score(const float *__restrict__ codebook, int k, const float (&v)[kD])
{
    const float4 *row = reinterpret_cast<const float4 *>(codebook) + k * 4;
    const float4 c0 = __ldg(row), c1 = __ldg(row + 1), c2 = __ldg(row + 2), c3 = __ldg(row + 3);
    float acc = __fmul_rn(c0.x, v[0]);
    acc = __fmaf_rn(c0.y, v[1], acc);  acc = __fmaf_rn(c0.z, v[2], acc);  acc = __fmaf_rn(c0.w, v[3], acc);
    acc = __fmaf_rn(c1.x, v[4], acc);  acc = __fmaf_rn(c1.y, v[5], acc);  acc = __fmaf_rn(c1.z, v[6], acc);
    acc = __fmaf_rn(c1.w, v[7], acc);  acc = __fmaf_rn(c2.x, v[8], acc);  acc = __fmaf_rn(c2.y, v[9], acc);
    acc = __fmaf_rn(c2.z, v[10], acc); acc = __fmaf_rn(c2.w, v[11], acc); acc = __fmaf_rn(c3.x, v[12], acc);
    acc = __fmaf_rn(c3.y, v[13], acc); acc = __fmaf_rn(c3.z, v[14], acc); acc = __fmaf_rn(c3.w, v[15], acc);
    return acc;
}

// max |x| over 16 accumulators as a depth-3 tree of 3-input maxima (a serial chain of eight
// dependent FMNMX3 made the first pass latency-bound: one warp per scheduler and group)
__device__ __forceinline__ float absmax16(const uint32_t (&s)[16])
{
    auto a = [&](int i) { return fabsf(__uint_as_float(s[i])); };
    const float m0 = fmaxf(fmaxf(a(0), a(1)), a(2)), m1 = fmaxf(fmaxf(a(3), a(4)), a(5));
    const float m2 = fmaxf(fmaxf(a(6), a(7)), a(8)), m3 = fmaxf(fmaxf(a(9), a(10)), a(11));
    const float m4 = fmaxf(fmaxf(a(12), a(13)), a(14));
    const float n0 = fmaxf(fmaxf(m0, m1), m2), n1 = fmaxf(fmaxf(m3, m4), a(15));
    return fmaxf(n0, n1);
}

__device__ __forceinline__ void named_barrier(int id, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
hsq_search_tck_kernel(const __grid_constant__ CUtensorMap map_grad, const __grid_constant__ CUtensorMap map_cb,
                      const float *__restrict__ codebook, int K, int n_chunks, int32_t *__restrict__ codes,
                      float *__restrict__ u_out, const int64_t *__restrict__ seg_start, int n_seg,
                      uint32_t *__restrict__ minmax_keys, const uint32_t *__restrict__ cn_key)
{
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *s_a = smem + kOffA;
    uint8_t *s_b = smem + kOffB;
    float4 *s_cm = reinterpret_cast<float4 *>(smem + kOffCm);
    uint8_t *misc = smem + off_misc(K);
    const uint32_t bar_afull = smem_u32(misc);
    const uint32_t bar_aempty = bar_afull + 16;
    const uint32_t bar_bfull = bar_aempty + 16;
    const uint32_t bar_bempty = bar_bfull + 32;
    const uint32_t bar_tfull = bar_bempty + 32;
    const uint32_t bar_tempty = bar_tfull + 16;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(misc + 16 * 8 + 16);
    float *s_amax = reinterpret_cast<float *>(misc + 16 * 8 + 32);
    int *s_best = reinterpret_cast<int *>(misc + 16 * 8 + 32 + 2 * 128 * 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int T = K / kTileN;                    // N-tiles per row tile (even)
    const int n_tiles = (n_chunks + kTileM - 1) / kTileM;
    const int tq = n_tiles / (int)gridDim.x, trem = n_tiles % (int)gridDim.x;
    const int tile0 = (int)blockIdx.x * tq + min((int)blockIdx.x, trem);
    const int my_tiles = tq + (((int)blockIdx.x < trem) ? 1 : 0);
    const int total = my_tiles * T;              // (row tile, N-tile) pairs of this CTA, index j = rtl * T + t

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_afull + 8 * i, 1);
            mbar_init(bar_aempty + 8 * i, 8);    // one arrive per epilogue warp
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 4);    // the four warps of the buffer's epilogue group
        }
        for (int i = 0; i < kBStages; ++i) {
            mbar_init(bar_bfull + 8 * i, 1);
            mbar_init(bar_bempty + 8 * i, 1);    // tcgen05.commit of the tile's MMAs
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    pdl_wait();   // the prepared codebook copy, the gradient and the min/max keys come from predecessors
    const float cn = __uint_as_float(__ldcg(cn_key));
    const float margin = kMargin * (1.0f + 1.0e-5f) * cn;

    // MMA of pair j: A = row tile j / T (ring of 2), B = ring slot j & 3, accumulators in TMEM buffer j & 1
    auto issue_mma = [&](int j) {
        const int rtl = j / T;
        mbar_wait_sleep(bar_afull + 8 * (rtl & 1), (rtl >> 1) & 1);
        mbar_wait_sleep(bar_bfull + 8 * (j & 3), (j >> 2) & 1);
        if (j >= 2) mbar_wait_hw(bar_tempty + 8 * (j & 1), ((j - 2) >> 1) & 1);   // the buffer's previous pass is read out
        tc_fence_after();
        const uint64_t adesc = make_desc(smem_u32(s_a + (rtl & 1) * kATileBytes));
        const uint64_t bdesc = make_desc(smem_u32(s_b + (j & 3) * kBTileBytes));
        const uint32_t taddr = tmem_base + (uint32_t)((j & 1) * kTileN);
        mma_tf32(taddr, adesc, bdesc, 0u, kIdesc);
        mma_tf32(taddr, adesc + 2, bdesc + 2, 1u, kIdesc);
        mma_commit(bar_tfull + 8 * (j & 1));
        mma_commit(bar_bempty + 8 * (j & 3));     // the operand slot may be refilled once these MMAs are done
    };

    if (warp == 0) {
        // ------------------------------------------------------ TMA producer ---
        if (lane == 0) {
            for (int rtl = 0; rtl < my_tiles; ++rtl) {
                mbar_wait_sleep(bar_aempty + 8 * (rtl & 1), ((rtl >> 1) & 1) ^ 1);
                mbar_expect_tx(bar_afull + 8 * (rtl & 1), kATileBytes);
                tma_load_2d(smem_u32(s_a + (rtl & 1) * kATileBytes), &map_grad, bar_afull + 8 * (rtl & 1), 0,
                            (tile0 + rtl) * kTileM);
                for (int t = 0; t < T; ++t) {
                    const int j = rtl * T + t;
                    mbar_wait_sleep(bar_bempty + 8 * (j & 3), ((j >> 2) & 1) ^ 1);
                    mbar_expect_tx(bar_bfull + 8 * (j & 3), kBTileBytes);
                    tma_load_2d(smem_u32(s_b + (j & 3) * kBTileBytes), &map_cb, bar_bfull + 8 * (j & 3), 0, t * kTileN);
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer ---
        if (lane == 0) {
            for (int j = 0; j < total; ++j) issue_mma(j);
        }
    } else if (warp >= 4) {
        // ----------------------------------------------------------- epilogue ---
        const int e = (warp - 4) >> 2;           // group = TMEM buffer
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        SegCache segc;
        MinMaxAcc mm;
        for (int rtl = 0; rtl < my_tiles; ++rtl) {
            const int c = (tile0 + rtl) * kTileM + row;
            const bool valid = c < n_chunks;
            float amax = 0.0f;
            for (int t = e; t < T; t += 2) {
                const int j = rtl * T + t;
                mbar_wait_hw(bar_tfull + 8 * e, (j >> 1) & 1);
                __syncwarp();
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(e * kTileN);
                // 256 approximate scores -> 16 coarse maxima (one per 16 codewords), software-pipelined loads
                float cm[16];
                uint32_t sa[16], sb[16];
                tmem_ld16(taddr, sa);
                tmem_ld_wait16(sa);
#pragma unroll
                for (int h = 0; h < 8; ++h) {
                    tmem_ld16(taddr + h * 32 + 16, sb);
                    cm[2 * h] = absmax16(sa);
                    tmem_ld_wait16(sb);
                    if (h + 1 < 8) tmem_ld16(taddr + h * 32 + 32, sa);
                    cm[2 * h + 1] = absmax16(sb);
                    if (h + 1 < 8) tmem_ld_wait16(sa);
                }
                // release the TMEM buffer (pair j + 2 is issued by the MMA warp as soon as all four warps have)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * e);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    s_cm[(t * 4 + q) * kTileM + row] = make_float4(cm[4 * q], cm[4 * q + 1], cm[4 * q + 2], cm[4 * q + 3]);
                    amax = fmaxf(amax, fmaxf(fmaxf(cm[4 * q], cm[4 * q + 1]), fmaxf(cm[4 * q + 2], cm[4 * q + 3])));
                }
            }
            // ---------------- end of the row tile: exchange the maxima, find and rescore the candidates ---
            s_amax[e * kTileM + row] = amax;
            named_barrier(1, 256);
            amax = fmaxf(amax, s_amax[(e ^ 1) * kTileM + row]);
            float v[kD];
            {
                const uint32_t arow = smem_u32(s_a) + (rtl & 1) * kATileBytes + row * 64;
                const int sw = (row >> 1) & 3;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float4 t4;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(t4.x), "=f"(t4.y), "=f"(t4.z), "=f"(t4.w) : "r"(arow + ((u ^ sw) << 4)));
                    v[4 * u] = t4.x; v[4 * u + 1] = t4.y; v[4 * u + 2] = t4.z; v[4 * u + 3] = t4.w;
                }
            }
            float n2 = 0.0f;
#pragma unroll
            for (int jx = 0; jx < kD; ++jx) n2 = fmaf(v[jx], v[jx], n2);
            // the gradient tile may be refilled once every epilogue warp has its rows in registers
            const uint32_t all_loaded = __ballot_sync(0xffffffffu, !(n2 < 0.0f));
            if (lane == 0) mbar_arrive(bar_aempty + 8 * (rtl & 1) + ((all_loaded == 0u) ? 64u : 0u));
            const float thr = amax - margin * sqrtf(n2);
            uint32_t any = 0u;
#pragma unroll
            for (int jx = 0; jx < kD; ++jx) any |= __float_as_uint(v[jx]);
            const bool zero = (any << 1) == 0u;
            // rows whose bound cannot be trusted (non-finite, or so small that operands may have been flushed)
            const bool special = !zero && (!(n2 < 3.0e38f) || !(amax < 3.0e38f) || n2 < 1.0e-30f);
            int best_bits = -1, best_k = 0;
            float best_u = 0.0f;
            auto rescore16 = [&](int k0) {
#pragma unroll 1
                for (int i = 0; i < kCoarse; i += 4) {
                    float p[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) p[x] = exact_score(codebook, k0 + i + x, v);
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const int ab = __float_as_int(p[x]) & 0x7fffffff;
                        if (ab > best_bits) { best_bits = ab; best_k = k0 + i + x; best_u = p[x]; }
                    }
                }
            };
            if (zero) {
                if (e == 0) rescore16(0);          // +-0 against every codeword: codeword 0 wins
            } else if (special) {
                if (e == 0) {
#pragma unroll 1
                    for (int k0 = 0; k0 < K; k0 += kCoarse) rescore16(k0);
                }
            } else {
                // Candidate coarse groups of this group's N-tiles as a bit mask first (16 bits per N-tile, two
                // N-tiles per word), THEN one rescoring pass per set bit: the warp runs max-over-lanes passes
                // (about three), not one pass per distinct position any of its 32 lanes asks for (about forty).
                uint32_t cmask[kMaxK / kTileN / 4];
#pragma unroll
                for (int w = 0; w < kMaxK / kTileN / 4; ++w) cmask[w] = 0u;
#pragma unroll
                for (int tt = 0; tt < kMaxK / kTileN / 2; ++tt) {
                    const int t = 2 * tt + e;
                    if (t < T) {
                        uint32_t bits = 0u;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 w = s_cm[(t * 4 + q) * kTileM + row];
                            bits |= (w.x >= thr ? 1u : 0u) << (4 * q);
                            bits |= (w.y >= thr ? 1u : 0u) << (4 * q + 1);
                            bits |= (w.z >= thr ? 1u : 0u) << (4 * q + 2);
                            bits |= (w.w >= thr ? 1u : 0u) << (4 * q + 3);
                        }
                        cmask[tt >> 1] |= bits << (16 * (tt & 1));
                    }
                }
#pragma unroll
                for (int w = 0; w < kMaxK / kTileN / 4; ++w) {
                    uint32_t m = cmask[w];
                    while (m != 0u) {
                        const int bit = __ffs((int)m) - 1;
                        m &= m - 1u;
                        const int tt = 2 * w + (bit >> 4);
                        rescore16((2 * tt + e) * kTileN + (bit & 15) * kCoarse);
                    }
                }
            }
            __syncwarp();
            // merge the two groups' winners: larger |score|, then the lower index
            if (e == 1) {
                s_best[row * 3] = best_bits;
                s_best[row * 3 + 1] = best_k;
                s_best[row * 3 + 2] = __float_as_int(best_u);
            }
            named_barrier(2, 256);
            if (e == 0) {
                const int ob = s_best[row * 3], ok = s_best[row * 3 + 1];
                const float ou = __int_as_float(s_best[row * 3 + 2]);
                if (ob > best_bits || (ob == best_bits && ob >= 0 && ok < best_k)) { best_bits = ob; best_k = ok; best_u = ou; }
                if (valid) {
                    codes[c] = best_k;
                    u_out[c] = best_u;
                }
                if (minmax_keys != nullptr) {
                    const int seg = valid ? cached_segment(segc, seg_start, n_seg, (int64_t)c) : -1;
                    minmax_add_warp(mm, valid, seg, best_u, minmax_keys);
                }
            }
        }
        if (e == 0 && minmax_keys != nullptr) minmax_flush_warp(mm, minmax_keys);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ------------------------------------------------------------------ host side ---
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// rows x 16 fp32 matrix, box = 16 x box_rows, 64-byte swizzle, out-of-range rows read as zero
static int make_map(CUtensorMap *map, const float *base, int64_t rows, int box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return GQ_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)kD, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kD * 4};
    cuuint32_t box[2] = {(cuuint32_t)kD, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (base %p, rows %lld)", (int)r, (const void *)base,
                  (long long)rows);
        return GQ_ERR_CUDA;
    }
    return GQ_OK;
}

}  // namespace tck

bool hsq_tck_supported(int d, int K, int code_bytes)
{
    if (d != tck::kD || K < 512 || K > tck::kMaxK || (K % 512) != 0 || code_bytes != 4) return false;
    if (const char *e = getenv("GQ_TCK")) {
        if (atoi(e) == 0) return false;
    }
    return hsq_tc_supported(16, 256, 1) && tck::encode_fn() != nullptr;
}

size_t hsq_tck_workspace_bytes(int d, int K)
{
    if (d != tck::kD || K < 512 || K > tck::kMaxK || (K % 512) != 0) return 0;
    return (size_t)K * d * 4 + 256;
}

// workspace: [K * 64 bytes TF32 codebook | 256 bytes: max norm key], 256-byte aligned
int hsq_search_tck(const float *grad, int64_t n_chunks, const float *codebook, int K, void *codes, float *u_out,
                   const int64_t *seg_start, int n_seg, uint32_t *minmax_keys, void *workspace, size_t workspace_bytes,
                   cudaStream_t st)
{
    using namespace tck;
    GQ_REQUIRE(workspace && workspace_bytes >= hsq_tck_workspace_bytes(kD, K) && ((uintptr_t)workspace & 255) == 0,
               "tcgen05 large-codebook search: workspace too small or misaligned");
    GQ_REQUIRE(((uintptr_t)grad & 15) == 0 && ((uintptr_t)codebook & 15) == 0, "TMA needs 16-byte aligned bases");
    GQ_REQUIRE(n_chunks > 0 && n_chunks < ((int64_t)1 << 31) - 256, "n_chunks out of range for one tensor map");
    float *cb_tf32 = reinterpret_cast<float *>(workspace);
    uint32_t *cn_key = reinterpret_cast<uint32_t *>((char *)workspace + (size_t)K * kD * 4);
    GQ_CUDA(cudaMemsetAsync(cn_key, 0, 4, st));
    hsq_tck_prep_kernel<<<(K + 255) / 256, 256, 0, st>>>(codebook, K, cb_tf32, cn_key);
    GQ_LAUNCH_CHECK("hsq_tck_prep");
    CUtensorMap mg, mc;
    int e = make_map(&mg, grad, n_chunks, kTileM);
    if (e) return e;
    e = make_map(&mc, cb_tf32, K, kTileN);
    if (e) return e;
    const int n_tiles = (int)((n_chunks + kTileM - 1) / kTileM);
    int sms = sm_count();
    if (const char *g = getenv("GQ_TC_GRID")) {
        int v = atoi(g);
        if (v > 0 && v < sms) sms = v;
    }
    const int grid = n_tiles < sms ? n_tiles : sms;
    const size_t smem = smem_bytes(K);
    GQ_CUDA(cudaFuncSetAttribute(hsq_search_tck_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GQ_CUDA(launch_pdl(hsq_search_tck_kernel, dim3(grid), dim3(kThreads), smem, st, mg, mc, codebook, K, (int)n_chunks,
                       (int32_t *)codes, u_out, seg_start, n_seg, minmax_keys, (const uint32_t *)cn_key));
    GQ_LAUNCH_CHECK("hsq_search_tck");
    return GQ_OK;
}

}  // namespace gq
