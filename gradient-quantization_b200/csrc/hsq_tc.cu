// hsq_tc.cu -- tcgen05 (TF32) nearest-codeword search with exact fp32 rescoring,
// for the headline shape d == 16, K == 256 (uint8 codes).
//
// Per 128-chunk tile the [128 x 16] x [16 x 256] inner-product contraction of
// nearest_neighbor_compressor.py:68 runs on the 5th-gen tensor cores:
//   TMA  : gradient tile (8 KB, row = chunk = 64 B, SWIZZLE_64B) -> 6-stage smem ring.
//          The codebook operand (16 KB, rounded to nearest TF32) and four exact, bank-conflict-free
//          "planes" of the codebook for rescoring (128 KB) are written once per CTA.
//   MMA  : 2 x tcgen05.mma kind::tf32 (M=128, N=256, K=8), fp32 accumulators in
//          TMEM, two 256-column buffers so the next tile's MMA overlaps the epilogue
//   EPI  : three groups of four warps take tiles round-robin; each of the 128 rows is one thread:
//          software-pipelined tcgen05.ld of its 256 approximate scores, of which only max|.| per
//          group of 4 codewords is kept (the scores never touch HBM), then it RESCORES in exact
//          fp32 every codeword of every group whose maximum lies within 2*eps of the row maximum.
// Bit-exactness argument: |approx_k - exact_k| <= eps(v) for all k  =>  the exact
// argmax (and every exact tie) has approx >= max_approx - 2 eps, so it is inside
// the rescored set; the rescoring is the same ascending-j FMA chain and the same
// "first index wins" rule as hsq_exact.cu, hence identical codes and u.
// eps(v) = (1.5 * 2^-10 + 4e-6) * ||v||_2 * max_k ||c_k||_2 -- see kMargin; rows with a non-finite
// or vanishing norm rescore all 256 codewords (all-zero rows: codeword 0).
#include <cuda.h>
#include <stdlib.h>
#include <time.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "gq_internal.cuh"
#include "tc_ptx.cuh"

namespace gq {
namespace tc {

constexpr int kD = 16;
constexpr int kK = 256;
constexpr int kTileM = 128;
constexpr int kStages = 6;   // ring of smem stages AND of TMEM full/empty barriers: a multiple of 2 (TMEM
                             // buffers) and of kEpiGroups, so each barrier is always waited on by the same warps
constexpr int kMaxEpiGroups = 3;   // epilogue groups of 4 warps, round-robin over tiles (2 or 3)
constexpr uint32_t kTileBytes = kTileM * kD * 4;  // 8192
constexpr uint32_t kCbBytes = kK * kD * 4;        // 16384
constexpr int kGroup = 4;                         // codewords per rescoring group
constexpr int kNumGroups = kK / kGroup;           // 64
constexpr uint32_t kPlaneBytes = kK * 128;        // one rescoring plane: 128-byte slot per codeword
constexpr int kPlanes = 4;
// Rescoring margin, 2 * eps / (||v|| max_k ||c_k||).  The tensor core reads the top 19 bits of an
// fp32 operand (TF32, truncation -- measured: adding the two truncation remainders back with extra
// MMAs leaves a residual of 1.1e-6 ||v||): the gradient operand is therefore off by < 2^-10 |v_j|.
// The codebook operand is rounded to nearest TF32 in shared memory once per CTA (cvt.rna), off by
// <= 2^-11 |c_kj|.  Every product is thus within (2^-10 + 2^-11 + 2^-21) |v_j c_kj|, the sum within
// 1.5 * 2^-10 ||v|| ||c_k|| (Cauchy-Schwarz); the fp32 accumulation of 16 terms adds
// < 16 * 2^-23 ||v|| ||c_k||; 4e-6 covers that, the rounding of ||v|| and of the threshold itself.
constexpr float kMargin = 2.0f * (1.5f / 1024.0f + 4.0e-6f);

static_assert(kStages % 2 == 0 && kStages % 3 == 0, "barrier ring must be a multiple of the buffer and group counts");
constexpr uint32_t kOffA = 0;
constexpr uint32_t kOffCb = kStages * kTileBytes;
constexpr uint32_t kOffPlanes = kOffCb + kCbBytes;
constexpr uint32_t kOffBar = kOffPlanes + kPlanes * kPlaneBytes;
constexpr uint32_t kSmemBytes = kOffBar + 256 + 1024;  // + alignment slack

// instruction descriptor: D fp32, A/B TF32, both K-major, N = 256, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((kK >> 3) << 17) | ((kTileM >> 4) << 24);

using namespace tcptx;

// Exact fp32 score of codeword k against v (same chain as hsq_exact.cu).
// Rescoring reads are lane-divergent (every lane wants a different codeword), so the
// codebook is replicated in shared memory in 4 "planes" laid out such that the 8 lanes of
// a quarter-warp always hit 8 different 16-byte bank groups, whatever codewords they ask for:
// lane class c = lane & 7, rotation r = c >> 1, half h = c & 1; plane r stores codeword k in a
// 128-byte slot, twice (h = 0, 1), with 16-byte unit u at bank group 4h + ((u + r) & 3).
// cw_base = plane r + 64 h (per lane), uo[u] = 16 * ((u + r) & 3).
__device__ __forceinline__ float exact_score(const uint8_t *__restrict__ cw_base, const uint32_t (&uo)[4], int k,
                                             const float (&v)[kD])
{
    const uint8_t *row = cw_base + k * 128;
    const float4 c0 = *reinterpret_cast<const float4 *>(row + uo[0]);
    const float4 c1 = *reinterpret_cast<const float4 *>(row + uo[1]);
    const float4 c2 = *reinterpret_cast<const float4 *>(row + uo[2]);
    const float4 c3 = *reinterpret_cast<const float4 *>(row + uo[3]);
    float acc = __fmul_rn(c0.x, v[0]);
    acc = __fmaf_rn(c0.y, v[1], acc);  acc = __fmaf_rn(c0.z, v[2], acc);  acc = __fmaf_rn(c0.w, v[3], acc);
    acc = __fmaf_rn(c1.x, v[4], acc);  acc = __fmaf_rn(c1.y, v[5], acc);  acc = __fmaf_rn(c1.z, v[6], acc);
    acc = __fmaf_rn(c1.w, v[7], acc);  acc = __fmaf_rn(c2.x, v[8], acc);  acc = __fmaf_rn(c2.y, v[9], acc);
    acc = __fmaf_rn(c2.z, v[10], acc); acc = __fmaf_rn(c2.w, v[11], acc); acc = __fmaf_rn(c3.x, v[12], acc);
    acc = __fmaf_rn(c3.y, v[13], acc); acc = __fmaf_rn(c3.z, v[14], acc); acc = __fmaf_rn(c3.w, v[15], acc);
    return acc;
}

// Optional fused tail: after a grid-wide barrier (all CTAs are resident: grid <= SM count,
// 1 CTA/SM) every CTA quantizes the norms of its own chunk range, so encode is ONE launch.
struct TcTail {
    uint8_t *l;              // nullptr = search only
    float *lbub;
    uint32_t *barrier;       // zeroed before the launch
    const float *uniforms;
    uint64_t seed, offset;
    float s;                 // 2^n_bit
    int random;
};

// Optional in-kernel preparation, so that an encode needs no separate init launch: CTA 0 resets the
// per-tensor min/max keys and then publishes `id` (unique per launch) in *flag; every epilogue warp
// checks the flag once, before its first min/max update (about a tile's worth of work later, so
// it practically never waits).  The spare warps 2-3 of every CTA also run the attached small
// reduction (the identity tensors' copy into the record).  flag == nullptr: keys are ready.
struct TcInit {
    unsigned long long *flag;
    unsigned long long id;
    Rider rider;
};

template <int kEpiGroups, bool kDebug>
__global__ void __launch_bounds__(128 + 128 * kEpiGroups, 1)
hsq_search_tc_kernel(const __grid_constant__ CUtensorMap map_grad,
                     const float *__restrict__ codebook, int64_t n_chunks, uint8_t *__restrict__ codes,
                     float *__restrict__ u_out,
                     const int64_t *__restrict__ seg_start, int n_seg, uint32_t *__restrict__ minmax_keys,
                     float *__restrict__ dbg_scores, int dbg_tiles, int flags, const TcTail tail, const TcInit init)
{
    constexpr int kThreads = 128 + 128 * kEpiGroups;
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, still a shared pointer
    uint8_t *s_a = smem + kOffA;
    uint8_t *s_cb = smem + kOffCb;
    uint8_t *s_planes = smem + kOffPlanes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kOffBar);
    // barrier slots: full[kStages], empty[kStages], tmem_full[kStages], tmem_empty[kStages], cb_full
    const uint32_t bar_full = smem_u32(bars);
    const uint32_t bar_empty = bar_full + 8 * kStages;
    const uint32_t bar_tfull = bar_empty + 8 * kStages;
    const uint32_t bar_tempty = bar_tfull + 8 * kStages;
    const uint32_t bar_cb = bar_tempty + 8 * kStages;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + kOffBar + 8 * (4 * kStages + 2));
    float *s_cn2 = reinterpret_cast<float *>(smem + kOffBar + 8 * (4 * kStages + 2) + 8);   // [8] per-warp max ||c_k||^2
    static_assert(8 * (4 * kStages + 2) + 8 + 32 <= 256, "barrier region");

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // debug builds: CTA 0 time-stamps pipeline events, trace[event * 128 + it] (events: 0 TMA issued,
    // 1 MMA issued, 2 accumulators seen by the epilogue, 3 TMEM released, 4 stage released, 5 tile done)
    long long *trace = nullptr;
    if (kDebug && dbg_scores != nullptr && blockIdx.x == 0)
        trace = reinterpret_cast<long long *>(dbg_scores + (size_t)dbg_tiles * kTileM * kK + n_chunks * 24);
#define GQ_TRACE(ev, it_) do { if (kDebug && trace != nullptr && (it_) < 128) trace[(ev) * 128 + (it_)] = clock64(); } while (0)
    // every CTA owns a contiguous range of tiles (balanced to within one tile)
    const int64_t n_tiles = (n_chunks + kTileM - 1) / kTileM;
    const int64_t tq = n_tiles / gridDim.x, trem = n_tiles % gridDim.x;
    const int64_t tile0 = (int64_t)blockIdx.x * tq + min((int64_t)blockIdx.x, trem);
    const int my_tiles = (int)(tq + (((int64_t)blockIdx.x < trem) ? 1 : 0));

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 4);   // one arrive per epilogue warp of the tile
        }
        for (int b = 0; b < kStages; ++b) {
            mbar_init(bar_tfull + 8 * b, 1);
            mbar_init(bar_tempty + 8 * b, 4);
        }
        mbar_init(bar_cb, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // rescoring planes (see exact_score): 4 planes x 256 codewords x 2 halves x 4 units, exact fp32;
    // and the MMA's B operand: the codebook rounded to nearest TF32 (see kMargin), written in the
    // K-major SWIZZLE_64B layout the descriptor expects (row k = 64 bytes, 16-byte unit u at
    // u ^ ((k >> 1) & 3)) -- by these threads, not by TMA, since it has to be rounded anyway
    for (int i = threadIdx.x; i < kPlanes * kK * 8; i += kThreads) {
        const int u = i & 3, h = (i >> 2) & 1, k = (i >> 3) & (kK - 1), r = i >> 11;
        const float4 val = __ldg(reinterpret_cast<const float4 *>(codebook) + k * 4 + u);
        *reinterpret_cast<float4 *>(s_planes + r * kPlaneBytes + k * 128 + 16 * (4 * h + ((u + r) & 3))) = val;
        if (r == 0 && h == 0) {
            uint4 t;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t.x) : "f"(val.x));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t.y) : "f"(val.y));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t.z) : "f"(val.z));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t.w) : "f"(val.w));
            *reinterpret_cast<uint4 *>(s_cb + k * 64 + ((u ^ ((k >> 1) & 3)) << 4)) = t;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // operand writes -> visible to the tensor core
    // largest codeword norm: the error bound (hence the margin) scales with it, so the search stays
    // exact for a codebook that is not unit-norm; a non-finite codebook disables the filter
    if (threadIdx.x < kK) {
        float c2 = 0.0f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(codebook) + threadIdx.x * 4 + u);
            c2 = fmaf(t.x, t.x, c2); c2 = fmaf(t.y, t.y, c2); c2 = fmaf(t.z, t.z, c2); c2 = fmaf(t.w, t.w, c2);
        }
        if (!(c2 < 3.0e38f)) c2 = __int_as_float(0x7f800000);   // NaN -> +inf so that the max keeps it
        c2 = warp_max(c2);
        if (lane == 0) s_cn2[warp] = c2;
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    float cn2 = s_cn2[0];
#pragma unroll
    for (int w = 1; w < kK / 32; ++w) cn2 = fmaxf(cn2, s_cn2[w]);
    // (a non-finite codebook gives margin = +inf: threshold -inf, every group is rescored)
    const float margin = kMargin * (1.0f + 1.0e-5f) * sqrtf(cn2);
    // everything above (barrier init, codebook planes, TMEM allocation) reads only the static
    // codebook and overlaps the tail of the previous kernel; the gradient, the min/max keys and
    // the output buffers are touched only after this point
    pdl_wait();

    if (warp == 0) {
        // ------------------------------------------------------ TMA producer ---
        if (lane == 0) {
            for (int it = 0; it < my_tiles; ++it) {
                const int s = it % kStages;
                const int64_t tile = tile0 + it;
                mbar_wait(bar_empty + 8 * s, ((it / kStages) & 1) ^ 1, flags & 1);
                mbar_expect_tx(bar_full + 8 * s, kTileBytes);
                tma_load_2d(smem_u32(s_a + s * kTileBytes), &map_grad, bar_full + 8 * s, 0, (int)(tile * kTileM));
                GQ_TRACE(0, it);
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer ---
        if (lane == 0) {
            const uint64_t bdesc = make_desc(smem_u32(s_cb));
            for (int it = 0; it < my_tiles; ++it) {
                const int s = it % kStages;
                const int b = it & 1;
                if (it >= 2)   // the epilogue of tile it-2 drained this TMEM buffer
                    mbar_wait(bar_tempty + 8 * ((it - 2) % kStages), ((it - 2) / kStages) & 1, flags & 2);
                mbar_wait(bar_full + 8 * s, (it / kStages) & 1, flags & 2);      // TMA landed this tile
                tc_fence_after();
                const uint64_t adesc = make_desc(smem_u32(s_a + s * kTileBytes));
                const uint32_t taddr = tmem_base + (uint32_t)(b * kK);
                mma_tf32(taddr, adesc, bdesc, 0u, kIdesc);              // k = 0..7   (bytes  0..31 of each row)
                mma_tf32(taddr, adesc + 2, bdesc + 2, 1u, kIdesc);      // k = 8..15  (bytes 32..63): +32 B = +2 units
                mma_commit(bar_tfull + 8 * s);
                GQ_TRACE(1, it);
            }
        }
    } else if (warp < 4) {
        // ------------------------------------- spare warps 2, 3: preparation ---
        if (init.flag != nullptr && blockIdx.x == 0 && warp == 3) {
            for (int i = lane; i < 2 * n_seg; i += 32) minmax_keys[i] = (i & 1) ? GQ_KEY_MAX_INIT : GQ_KEY_MIN_INIT;
            __threadfence();
            __syncwarp();
            if (lane == 0)
                asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(init.flag), "l"(init.id) : "memory");
        }
        rider_run(init.rider, (int64_t)blockIdx.x * 64 + (threadIdx.x - 64), (int64_t)gridDim.x * 64);
    } else if (warp >= 4) {
        // ----------------------------------------------------------- epilogue ---
        bool keys_ready = init.flag == nullptr;
        const int egroup = (warp - 4) >> 2;   // handles local tiles it = egroup (mod kEpiGroups)
        const int quad = warp & 3;            // TMEM lane quadrant this warp may read
        const int row = quad * 32 + lane;     // row of the tile == TMEM lane
        const int rot = (lane & 7) >> 1, hlf = lane & 1;
        const uint8_t *cw_base = s_planes + rot * kPlaneBytes + 64 * hlf;
        uint32_t uo[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) uo[u] = 16u * ((u + rot) & 3);
        SegCache segc;
        MinMaxAcc mm;
        for (int it = egroup; it < my_tiles; it += kEpiGroups) {
            const int s = it % kStages;
            const int b = it & 1;
            const int64_t tile = tile0 + it;
            const int64_t c = tile * kTileM + row;
            const bool valid = c < n_chunks;
            mbar_wait(bar_full + 8 * s, (it / kStages) & 1, flags & 4);   // TMA data visible to this thread
            mbar_wait(bar_tfull + 8 * s, (it / kStages) & 1, flags & 4);  // accumulators complete
            __syncwarp();                                      // converged before .sync.aligned TMEM loads
            tc_fence_after();
            if (quad == 0 && lane == 0) GQ_TRACE(2, it);

            // pass over the 256 approximate scores of this row: max |.| per group of 4 codewords
            float gm[kNumGroups];
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * kK);
            if (!kDebug || dbg_tiles == 0) {   // (debug builds keep it unless they dump the scores)
                // software-pipelined: the load of the next 16 columns is in flight while the
                // previous 16 are reduced (two 16-register buffers); 60 vs 64 us on B200
                uint32_t sa[16], sb[16];
                tmem_ld16(taddr, sa);
                tmem_ld_wait16(sa);
#pragma unroll
                for (int h = 0; h < kK / 32; ++h) {
                    tmem_ld16(taddr + h * 32 + 16, sb);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float m = fmaxf(fabsf(__uint_as_float(sa[4 * g])), fabsf(__uint_as_float(sa[4 * g + 1])));
                        gm[h * 8 + g] = absmax3(m, sa[4 * g + 2], sa[4 * g + 3]);
                    }
                    tmem_ld_wait16(sb);
                    if (h + 1 < kK / 32) tmem_ld16(taddr + h * 32 + 32, sa);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float m = fmaxf(fabsf(__uint_as_float(sb[4 * g])), fabsf(__uint_as_float(sb[4 * g + 1])));
                        gm[h * 8 + 4 + g] = absmax3(m, sb[4 * g + 2], sb[4 * g + 3]);
                    }
                    if (h + 1 < kK / 32) tmem_ld_wait16(sa);
                }
            } else {
#pragma unroll
            for (int blk = 0; blk < kK / 32; ++blk) {
                uint32_t sc[32];
                tmem_ld32(taddr + blk * 32, sc);
                tmem_ld_wait();
                if (kDebug && dbg_scores != nullptr && tile < dbg_tiles) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        dbg_scores[(tile * kTileM + row) * kK + blk * 32 + j] = __uint_as_float(sc[j]);
                }
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    float m = fmaxf(fabsf(__uint_as_float(sc[4 * g])), fabsf(__uint_as_float(sc[4 * g + 1])));
                    gm[blk * 8 + g] = absmax3(m, sc[4 * g + 2], sc[4 * g + 3]);
                }
            }
            }
            // TMEM buffer b may be overwritten by the MMA of local tile it + 2
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * s);
            if (quad == 0 && lane == 0) GQ_TRACE(3, it);

            // this row's chunk, from the (swizzled) smem tile
            float v[kD];
            {
                const uint8_t *arow = s_a + s * kTileBytes + row * 64;
                const int sw = (row >> 1) & 3;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 t = *reinterpret_cast<const float4 *>(arow + ((u ^ sw) << 4));
                    v[4 * u] = t.x; v[4 * u + 1] = t.y; v[4 * u + 2] = t.z; v[4 * u + 3] = t.w;
                }
            }
            float n2 = 0.0f;
#pragma unroll
            for (int j = 0; j < kD; ++j) n2 = fmaf(v[j], v[j], n2);
            // Release the smem stage to the TMA producer only after EVERY lane's loads of v have
            // returned: the ballot consumes n2 (hence all four LDS of each lane), and the arrive
            // address depends on the ballot.  Measured on B200: without this dependency the
            // mbarrier arrive can be performed before the warp's last LDS.128 has read the stage
            // (a plain __syncwarp() is elided by the compiler), the producer's next TMA then
            // overwrites the row under the read -- tests/tc_diag.py shows v[12..15] of tile it+6.
            const uint32_t all_loaded = __ballot_sync(0xffffffffu, !(n2 < 0.0f));   // always all ones
            if (lane == 0) mbar_arrive(bar_empty + 8 * s + ((all_loaded == 0u) ? 8u : 0u));
            if (quad == 0 && lane == 0) GQ_TRACE(4, it);

            // row maximum as a tree (four independent chains, then a 4-way merge) rather than one
            // 32-deep dependent chain
            float am[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                am[q] = fmaxf(gm[16 * q], gm[16 * q + 1]);
#pragma unroll
                for (int g = 2; g < 16; g += 2) am[q] = fmaxf(am[q], fmaxf(gm[16 * q + g], gm[16 * q + g + 1]));
            }
            const float amax = fmaxf(fmaxf(am[0], am[1]), fmaxf(am[2], am[3]));
            const float thr = amax - margin * sqrtf(n2);
            // candidate groups: gm[g] >= thr.  d = gm - thr on the FMA pipe, sign bits funnelled
            // into two 32-bit words (bit g of below[g >> 5] set = group g is below the threshold).
            uint32_t bl[4] = {0u, 0u, 0u, 0u};   // four independent chains of 16 groups each
#pragma unroll
            for (int g = 15; g >= 0; --g) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    bl[q] = __funnelshift_l(__float_as_uint(gm[16 * q + g] - thr), bl[q], 1);
            }
            uint64_t cand = ~(((uint64_t)(bl[0] & 0xffffu)) | ((uint64_t)(bl[1] & 0xffffu) << 16) |
                              ((uint64_t)(bl[2] & 0xffffu) << 32) | ((uint64_t)(bl[3] & 0xffffu) << 48));
            // non-finite or overflowing norm, NaN scores, or an empty set: rescore everything
            // the same for rows so small that operands or products may have been flushed to zero
            // in the tensor core (the error bound is relative to ||v||); an all-zero row scores
            // +-0 against every codeword, so codeword 0 wins and only group 0 is needed
            if (!(n2 < 3.0e38f) || !(amax < 3.0e38f) || cand == 0ull || n2 < 1.0e-30f) {
                uint32_t any = 0u;
#pragma unroll
                for (int j = 0; j < kD; ++j) any |= __float_as_uint(v[j]);
                cand = ((any << 1) == 0u) ? 1ull : ~0ull;
            }
            const uint32_t mask0 = (uint32_t)cand;

            int best_bits = -1, best_k = 0;
            float best_u = 0.0f;
            while (cand) {
                const int g = __ffsll((long long)cand) - 1;
                cand &= cand - 1;
                float p[kGroup];
#pragma unroll
                for (int i = 0; i < kGroup; ++i) p[i] = exact_score(cw_base, uo, g * kGroup + i, v);
                // group maximum first (3 FMNMX), one compare against the running best, and only a
                // winning group pays for locating its first maximal codeword
                const float gmax = fmaxf(fmaxf(fabsf(p[0]), fabsf(p[1])), fmaxf(fabsf(p[2]), fabsf(p[3])));
                const int gb = __float_as_int(gmax);
                const float psum = (p[0] + p[1]) + (p[2] + p[3]);   // NaN (or inf - inf) takes the slow path
                const bool has_nan = psum != psum;
                if (gb > best_bits || has_nan || (flags & 8)) {
#pragma unroll
                    for (int i = 0; i < kGroup; ++i) {
                        const int ab = __float_as_int(p[i]) & 0x7fffffff;
                        if (ab > best_bits) { best_bits = ab; best_k = g * kGroup + i; best_u = p[i]; }
                    }
                }
            }
            __syncwarp();
            if (kDebug && dbg_scores != nullptr && valid) {
                float *aux = dbg_scores + (size_t)dbg_tiles * kTileM * kK + c * 24;
                for (int j = 0; j < kD; ++j) aux[8 + j] = v[j];
                aux[0] = v[0];
                aux[1] = amax;
                aux[2] = (float)it;
                aux[3] = (float)blockIdx.x;
                float vs = 0.f;
                for (int j = 0; j < kD; ++j) vs += v[j];
                aux[4] = vs;
                float cs = 0.f;
                {
                    const uint8_t *rowp = s_cb + best_k * 64;
                    for (int j = 0; j < 16; ++j) cs += *reinterpret_cast<const float *>(rowp + j * 4);
                }
                aux[5] = cs;                       // order-independent of the swizzle: sum over the row
                aux[6] = __uint_as_float(mask0);
                aux[7] = thr;
            }
            if (valid) {
                codes[c] = (uint8_t)best_k;
                u_out[c] = best_u;
            }
            if (minmax_keys != nullptr) {
                if (!keys_ready) {   // CTA 0 has reset the keys (see TcInit); bounded wait, then trap
                    unsigned long long seen;
                    uint32_t spins = 0;
                    do {
                        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(init.flag) : "memory");
                        if (seen == init.id) break;
                        __nanosleep(64);
                    } while (++spins < (1u << 22));
                    if (seen != init.id) __trap();
                    keys_ready = true;
                }
                const int seg = valid ? cached_segment(segc, seg_start, n_seg, c) : -1;
                minmax_add_warp(mm, valid, seg, best_u, minmax_keys);
            }
            if (quad == 0 && lane == 0) GQ_TRACE(5, it);
        }
        if (minmax_keys != nullptr) minmax_flush_warp(mm, minmax_keys);
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
    if (tail.l != nullptr) {
        // grid barrier: every CTA's min/max atomics are performed before anyone reads lb/ub
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(tail.barrier, 1u);
            uint32_t spins = 0;
            while (*reinterpret_cast<volatile uint32_t *>(tail.barrier) < gridDim.x) {
                __nanosleep(64);
                if (++spins > (1u << 24)) __trap();
            }
            __threadfence();
        }
        __syncthreads();
        if (blockIdx.x == 0) {
            for (int i = threadIdx.x; i < 2 * n_seg; i += kThreads) tail.lbub[i] = key_to_float(__ldcg(minmax_keys + i));
        }
        const int64_t c_begin = tile0 * kTileM;
        const int64_t c_end = min((tile0 + my_tiles) * (int64_t)kTileM, n_chunks);
        quantize_range<uint8_t, true>(u_out, c_begin, c_end, n_chunks, (int)threadIdx.x, kThreads, seg_start, n_seg,
                                      tail.s, tail.random, tail.uniforms, tail.seed, tail.offset, tail.l,
                                      minmax_keys);
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

}  // namespace tc

bool hsq_tc_supported(int d, int K, int code_bytes)
{
    if (d != tc::kD || K != tc::kK || code_bytes != 1) return false;
    static int cc = -1;
    if (cc < 0) {
        int dev = 0, major = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess)
            cc = major;
        else
            cc = 0;
    }
    return cc == 10 && tc::encode_fn() != nullptr;
}

size_t hsq_tc_workspace_bytes(int64_t) { return 0; }

static int launch_tc(const float *grad, int64_t n_chunks, const float *codebook, void *codes, float *u_out,
                     const int64_t *seg_start, int n_seg, uint32_t *minmax_keys, float *dbg_scores,
                     int dbg_tiles, const tc::TcTail &tail, cudaStream_t st, const tc::TcInit &init = tc::TcInit{})
{
    GQ_REQUIRE(((uintptr_t)grad & 15) == 0 && ((uintptr_t)codebook & 15) == 0, "TMA needs 16-byte aligned bases");
    GQ_REQUIRE(n_chunks < ((int64_t)1 << 31) - 256, "n_chunks too large for one tensor map");
    CUtensorMap mg;
    int e = tc::make_map(&mg, grad, n_chunks, tc::kTileM);
    if (e) return e;
    const int64_t n_tiles = (n_chunks + tc::kTileM - 1) / tc::kTileM;
    int sms = sm_count();
    if (const char *g = getenv("GQ_TC_GRID")) {   // debugging aid: force few CTAs -> many tiles per CTA
        int v = atoi(g);
        if (v > 0 && v < sms) sms = v;
    }
    const int grid = (int)(n_tiles < sms ? n_tiles : sms);
    int flags = 0;
    if (const char *f = getenv("GQ_TC_FLAGS")) flags = atoi(f);
    int groups = 3;   // measured on B200: 3 epilogue groups 75 us, 2 groups 80 us (ResNet-50 gradient)
    if (const char *f = getenv("GQ_TC_GROUPS")) groups = atoi(f) == 2 ? 2 : 3;
    const bool dbg = dbg_scores != nullptr;
#define GQ_TC_LAUNCH(G, DBG)                                                                                   \
    do {                                                                                                       \
        auto kern = tc::hsq_search_tc_kernel<G, DBG>;                                                          \
        GQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes)); \
        GQ_CUDA(launch_pdl(kern, dim3(grid), dim3(128 + 128 * G), (size_t)tc::kSmemBytes, st, mg, codebook,         \
                           n_chunks, (uint8_t *)codes, u_out, seg_start, n_seg, minmax_keys, dbg_scores,        \
                           dbg_tiles, flags, tail, init));                                                     \
    } while (0)
    if (dbg) {
        if (groups == 3) GQ_TC_LAUNCH(3, true); else GQ_TC_LAUNCH(2, true);
    } else {
        if (groups == 3) GQ_TC_LAUNCH(3, false); else GQ_TC_LAUNCH(2, false);
    }
#undef GQ_TC_LAUNCH
    GQ_LAUNCH_CHECK("hsq_search_tc");
    return GQ_OK;
}

int hsq_search_tc_dbg(const float *grad, int64_t n_chunks, const float *codebook, void *codes, float *u_out,
                      const int64_t *seg_start, int n_seg, uint32_t *minmax_keys, float *dbg_scores,
                      int dbg_tiles, cudaStream_t st)
{
    tc::TcTail tail = {};
    return launch_tc(grad, n_chunks, codebook, codes, u_out, seg_start, n_seg, minmax_keys, dbg_scores, dbg_tiles,
                     tail, st);
}

int hsq_encode_tc_fused(const float *grad, int64_t n_chunks, const float *codebook, void *codes, float *u_out,
                        const int64_t *seg_start, int n_seg, uint32_t *minmax_keys, uint32_t *barrier,
                        int n_bit, int random, const float *uniforms, uint64_t seed, uint64_t offset,
                        uint8_t *l, float *lbub, cudaStream_t st)
{
    tc::TcTail tail = {};
    tail.l = l;
    tail.lbub = lbub;
    tail.barrier = barrier;
    tail.uniforms = uniforms;
    tail.seed = seed;
    tail.offset = offset;
    tail.s = (float)(1u << n_bit);
    tail.random = random;
    return launch_tc(grad, n_chunks, codebook, codes, u_out, seg_start, n_seg, minmax_keys, nullptr, 0, tail, st);
}

// search that also resets the min/max keys (no separate init launch) and carries `rider`
int hsq_search_tc_prepared(const float *grad, int64_t n_chunks, const float *codebook, void *codes, float *u_out,
                           const int64_t *seg_start, int n_seg, uint32_t *minmax_keys, uint64_t *flag,
                           const Rider &rider, cudaStream_t st)
{
    // launch ids never repeat within a process; the random upper half makes a stale or
    // uninitialised flag word equal to a live id with probability 2^-64
    static std::atomic<unsigned long long> counter{[] {
        unsigned long long seed = (unsigned long long)(uintptr_t)&seed ^ (unsigned long long)clock();
        seed = seed * 6364136223846793005ull + 1442695040888963407ull;
        return (seed >> 16) << 32;
    }()};
    tc::TcInit init;
    init.flag = reinterpret_cast<unsigned long long *>(flag);
    init.id = counter.fetch_add(1) + 1;
    init.rider = rider;
    tc::TcTail tail = {};
    return launch_tc(grad, n_chunks, codebook, codes, u_out, seg_start, n_seg, minmax_keys, nullptr, 0, tail, st, init);
}

int hsq_search_tc(const float *grad, int64_t n_chunks, int d, const float *codebook, int K, void *codes,
                  int code_bytes, float *u_out, const int64_t *seg_start, int n_seg, uint32_t *minmax_keys,
                  void *, size_t, cudaStream_t st)
{
    GQ_REQUIRE(hsq_tc_supported(d, K, code_bytes), "tcgen05 search: unsupported shape");
    return hsq_search_tc_dbg(grad, n_chunks, codebook, codes, u_out, seg_start, n_seg, minmax_keys, nullptr, 0, st);
}

}  // namespace gq

// Test hook: run the tcgen05 search and also dump the raw TF32 scores of the first
// dbg_tiles tiles ([dbg_tiles*128, 256] fp32), so tests can measure the approximation
// error the rescoring margin has to cover.
extern "C" int gq_hsq_tc_debug(const float *grad, int64_t n_chunks, const float *codebook, void *codes,
                               float *u_out, const int64_t *seg_start, int n_seg, float *dbg_scores,
                               int dbg_tiles, gq_stream_t stream)
{
    using namespace gq;
    GQ_REQUIRE(hsq_tc_supported(16, 256, 1), "tcgen05 search needs an sm_100 device");
    return hsq_search_tc_dbg(grad, n_chunks, codebook, codes, u_out, seg_start, n_seg, nullptr, dbg_scores,
                             dbg_tiles, as_stream(stream));
}
