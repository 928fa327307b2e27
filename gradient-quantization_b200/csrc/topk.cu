// topk.cu -- segmented top-k sparsification by radix select.
// Replaces TopKSparsificationCompressor.compress
// (compressors/topk_sparsification_compressor.py:18-23): per tensor keep the
// k largest |v| (ties at the cut resolved towards the lowest index -- torch.topk
// leaves the order unspecified), zero the rest.
//
// key = bit pattern of |v| (31 bits, monotone; NaN sorts above inf like
// torch.topk).  Three histogram passes (11 + 10 + 10 bits) find the k-th largest
// key T of every tensor at once, one counting pass gives each 1024-element tile
// its output offset, one write pass emits the dense masked tensor and/or the
// (index, value) wire form in ascending index order.  HBM-bound; the passes
// re-read the gradient (L2-resident for models up to ~100 MB).
#include "gq_internal.cuh"

namespace gq {

constexpr int kTile = 1024;      // elements per tile (256 threads x 4)
constexpr int kBins = 2048;

struct TopkState {
    uint32_t prefix;    // key bits decided so far (aligned to the top)
    uint32_t k_rem;     // how many still to take among keys matching the prefix
    uint32_t T;         // final threshold key (after pass 3)
    uint32_t count_eq;  // number of elements with key == T
};

__device__ __forceinline__ uint32_t key_of(float x) { return __float_as_uint(x) & 0x7fffffffu; }

// tile_prefix[s] = number of tiles of segments < s
__global__ void topk_setup_kernel(const int64_t *__restrict__ seg_start, const int64_t *__restrict__ k,
                                  int n_seg, int64_t *__restrict__ tile_prefix,
                                  TopkState *__restrict__ state)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int64_t acc = 0;
        for (int s = 0; s < n_seg; ++s) {
            tile_prefix[s] = acc;
            acc += (seg_start[s + 1] - seg_start[s] + kTile - 1) / kTile;
        }
        tile_prefix[n_seg] = acc;
    }
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_seg; s += gridDim.x * blockDim.x) {
        TopkState st;
        st.prefix = 0;
        int64_t len = seg_start[s + 1] - seg_start[s];
        int64_t kk = k[s] < len ? k[s] : len;
        st.k_rem = (uint32_t)(kk < 0 ? 0 : kk);
        st.T = 0xffffffffu;
        st.count_eq = 0;
        state[s] = st;
    }
}

__device__ __forceinline__ bool tile_range(const int64_t *__restrict__ seg_start,
                                           const int64_t *__restrict__ tile_prefix, int n_seg,
                                           int64_t b, int &s, int64_t &lo, int64_t &hi)
{
    if (b >= tile_prefix[n_seg]) return false;
    s = find_segment(tile_prefix, n_seg, b);
    // skip empty segments sharing the same tile_prefix value
    while (s + 1 < n_seg && tile_prefix[s + 1] <= b) ++s;
    lo = seg_start[s] + (b - tile_prefix[s]) * kTile;
    hi = min(lo + (int64_t)kTile, seg_start[s + 1]);
    return true;
}

// PASS 0: bits 30..20 (2048 bins); PASS 1: bits 19..10; PASS 2: bits 9..0 (1024 bins)
template <int PASS>
__global__ void __launch_bounds__(256)
topk_hist_kernel(const float *__restrict__ v, const int64_t *__restrict__ seg_start,
                 const int64_t *__restrict__ tile_prefix, int n_seg,
                 const TopkState *__restrict__ state, uint32_t *__restrict__ hist)
{
    __shared__ uint32_t s_hist[kBins];
    const int64_t total_tiles = tile_prefix[n_seg];
    for (int64_t b = blockIdx.x; b < total_tiles; b += gridDim.x) {
        int s; int64_t lo, hi;
        if (!tile_range(seg_start, tile_prefix, n_seg, b, s, lo, hi)) break;
        const TopkState st = state[s];
        if (st.k_rem == 0) continue;  // nothing to take (k == 0)
        for (int i = threadIdx.x; i < kBins; i += 256) s_hist[i] = 0;
        __syncthreads();
        for (int64_t i = lo + threadIdx.x; i < hi; i += 256) {
            const uint32_t key = key_of(v[i]);
            if (PASS == 0) {
                atomicAdd(&s_hist[key >> 20], 1u);
            } else if (PASS == 1) {
                if ((key >> 20) == st.prefix) atomicAdd(&s_hist[(key >> 10) & 1023u], 1u);
            } else {
                if ((key >> 10) == st.prefix) atomicAdd(&s_hist[key & 1023u], 1u);
            }
        }
        __syncthreads();
        const int bins = (PASS == 0) ? kBins : 1024;
        for (int i = threadIdx.x; i < bins; i += 256) {
            const uint32_t c = s_hist[i];
            if (c) atomicAdd(hist + (int64_t)s * kBins + i, c);
        }
        __syncthreads();
    }
}

// one block per segment: pick the bin holding the k_rem-th largest key, update
// the state, clear the histogram for the next pass.
template <int PASS>
__global__ void __launch_bounds__(256)
topk_scan_kernel(uint32_t *__restrict__ hist, TopkState *__restrict__ state)
{
    __shared__ uint32_t s_part[256];
    __shared__ uint32_t s_suffix[257];
    const int s = blockIdx.x;
    uint32_t *h = hist + (int64_t)s * kBins;
    TopkState st = state[s];
    constexpr int BINS = (PASS == 0) ? kBins : 1024;
    constexpr int PER = BINS / 256;
    if (st.k_rem == 0) return;
    uint32_t local[PER];
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) { local[j] = h[threadIdx.x * PER + j]; sum += local[j]; }
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        s_suffix[256] = 0;
        for (int t = 255; t >= 0; --t) { acc += s_part[t]; s_suffix[t] = acc; }
    }
    __syncthreads();
    // suffix count of bins strictly above this thread's bins
    uint32_t above = s_suffix[threadIdx.x + 1];
    if (above < st.k_rem && above + sum >= st.k_rem) {
        uint32_t acc = above;
        for (int j = PER - 1; j >= 0; --j) {
            if (acc + local[j] >= st.k_rem) {
                const uint32_t bin = threadIdx.x * PER + j;
                TopkState ns = st;
                ns.k_rem = st.k_rem - acc;
                if (PASS == 0) ns.prefix = bin;
                else if (PASS == 1) ns.prefix = (st.prefix << 10) | bin;
                else { ns.T = (st.prefix << 10) | bin; ns.count_eq = local[j]; }
                state[s] = ns;
                break;
            }
            acc += local[j];
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PER; ++j) h[threadIdx.x * PER + j] = 0;
}

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t x, uint32_t *s_warp, uint32_t &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    __syncthreads();
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        uint32_t c = s_warp[w];
        if (w < warp) base += c;
        tot += c;
    }
    total = tot;
    return base + inc - x;
}

// per tile: number of keys > T and == T
__global__ void __launch_bounds__(256)
topk_count_kernel(const float *__restrict__ v, const int64_t *__restrict__ seg_start,
                  const int64_t *__restrict__ tile_prefix, int n_seg,
                  const TopkState *__restrict__ state, uint32_t *__restrict__ tile_gt,
                  uint32_t *__restrict__ tile_eq)
{
    const int64_t total_tiles = tile_prefix[n_seg];
    for (int64_t b = blockIdx.x; b < total_tiles; b += gridDim.x) {
        int s; int64_t lo, hi;
        if (!tile_range(seg_start, tile_prefix, n_seg, b, s, lo, hi)) break;
        const uint32_t T = state[s].T;
        uint32_t gt = 0, eq = 0;
        for (int64_t i = lo + threadIdx.x; i < hi; i += 256) {
            const uint32_t key = key_of(v[i]);
            gt += key > T;
            eq += key == T;
        }
        gt = __reduce_add_sync(0xffffffffu, gt);
        eq = __reduce_add_sync(0xffffffffu, eq);
        __shared__ uint32_t s_gt[8], s_eq[8];
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { s_gt[threadIdx.x >> 5] = gt; s_eq[threadIdx.x >> 5] = eq; }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t a = 0, c = 0;
            for (int w = 0; w < 8; ++w) { a += s_gt[w]; c += s_eq[w]; }
            tile_gt[b] = a;
            tile_eq[b] = c;
        }
    }
}

// one block per segment: exclusive scan of its tiles' counts (in place)
__global__ void __launch_bounds__(256)
topk_tilescan_kernel(const int64_t *__restrict__ tile_prefix, uint32_t *__restrict__ tile_gt,
                     uint32_t *__restrict__ tile_eq)
{
    __shared__ uint32_t s_warp[8];
    const int s = blockIdx.x;
    const int64_t t0 = tile_prefix[s], t1 = tile_prefix[s + 1];
    uint32_t carry_gt = 0, carry_eq = 0;
    for (int64_t base = t0; base < t1; base += 256) {
        const int64_t t = base + threadIdx.x;
        uint32_t g = t < t1 ? tile_gt[t] : 0u;
        uint32_t e = t < t1 ? tile_eq[t] : 0u;
        uint32_t tg, te;
        uint32_t xg = block_exclusive_scan(g, s_warp, tg);
        uint32_t xe = block_exclusive_scan(e, s_warp, te);
        if (t < t1) { tile_gt[t] = carry_gt + xg; tile_eq[t] = carry_eq + xe; }
        carry_gt += tg;
        carry_eq += te;
    }
}

__global__ void __launch_bounds__(256)
topk_write_kernel(const float *__restrict__ v, const int64_t *__restrict__ seg_start,
                  const int64_t *__restrict__ k_prefix, const int64_t *__restrict__ tile_prefix,
                  int n_seg, const TopkState *__restrict__ state, const uint32_t *__restrict__ tile_gt,
                  const uint32_t *__restrict__ tile_eq, float *__restrict__ out_dense,
                  int32_t *__restrict__ out_idx, float *__restrict__ out_val)
{
    __shared__ uint32_t s_warp[8];
    const int64_t total_tiles = tile_prefix[n_seg];
    for (int64_t b = blockIdx.x; b < total_tiles; b += gridDim.x) {
        int s; int64_t lo, hi;
        if (!tile_range(seg_start, tile_prefix, n_seg, b, s, lo, hi)) break;
        const TopkState st = state[s];
        const uint32_t gt_before = tile_gt[b], eq_before = tile_eq[b];
        // thread owns 4 consecutive elements so ranks follow the index order
        const int64_t i0 = lo + (int64_t)threadIdx.x * 4;
        float x[4];
        uint32_t key[4];
        uint32_t eq_local = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const bool in = (i0 + t < hi);
            x[t] = in ? v[i0 + t] : 0.0f;
            key[t] = in ? key_of(x[t]) : 0u;
            if (in && key[t] == st.T) ++eq_local;
        }
        uint32_t tot;
        uint32_t eq_rank = eq_before + block_exclusive_scan(eq_local, s_warp, tot);
        bool sel[4];
        uint32_t sel_local = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const bool in = (i0 + t < hi);
            bool take = false;
            if (in && st.k_rem > 0) {
                if (key[t] > st.T) take = true;
                else if (key[t] == st.T) { take = (eq_rank < st.k_rem); ++eq_rank; }
            }
            sel[t] = take;
            sel_local += take;
        }
        if (out_dense) {
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (i0 + t < hi) out_dense[i0 + t] = __fmul_rn(x[t], sel[t] ? 1.0f : 0.0f);
        }
        if (out_idx || out_val) {
            uint32_t pos = block_exclusive_scan(sel_local, s_warp, tot);
            const uint32_t eq_taken_before = eq_before < st.k_rem ? eq_before : st.k_rem;
            int64_t o = k_prefix[s] + gt_before + eq_taken_before + pos;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (sel[t]) {
                    if (out_idx) out_idx[o] = (int32_t)(i0 + t);
                    if (out_val) out_val[o] = x[t];
                    ++o;
                }
            }
        }
    }
}

struct TopkWorkspace {
    int64_t *tile_prefix;
    TopkState *state;
    uint32_t *hist;
    uint32_t *tile_gt;
    uint32_t *tile_eq;
    size_t bytes;
};

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

static TopkWorkspace carve(void *base, int64_t n, int n_seg)
{
    TopkWorkspace w;
    const int64_t max_tiles = n / kTile + n_seg + 1;
    size_t p = 0;
    char *b = (char *)base;
    w.tile_prefix = (int64_t *)(b + p); p += align256((size_t)(n_seg + 1) * 8);
    w.state = (TopkState *)(b + p);     p += align256((size_t)n_seg * sizeof(TopkState));
    w.hist = (uint32_t *)(b + p);       p += align256((size_t)n_seg * kBins * 4);
    w.tile_gt = (uint32_t *)(b + p);    p += align256((size_t)max_tiles * 4);
    w.tile_eq = (uint32_t *)(b + p);    p += align256((size_t)max_tiles * 4);
    w.bytes = p;
    return w;
}

size_t topk_workspace_bytes(int64_t n, int n_seg) { return carve(nullptr, n, n_seg).bytes; }

int topk_select(const float *grad, int64_t n, const int64_t *seg_start, const int64_t *k,
                const int64_t *k_prefix, int n_seg, float *out_dense, int32_t *out_idx, float *out_val,
                void *workspace, cudaStream_t st)
{
    if (n == 0 || n_seg == 0) return GQ_OK;
    TopkWorkspace w = carve(workspace, n, n_seg);
    const int64_t max_tiles = n / kTile + n_seg + 1;
    int64_t cap = (int64_t)sm_count() * 8;
    const int grid = (int)(max_tiles < cap ? max_tiles : cap);
    GQ_CUDA(cudaMemsetAsync(w.hist, 0, (size_t)n_seg * kBins * 4, st));
    topk_setup_kernel<<<(n_seg + 127) / 128, 128, 0, st>>>(seg_start, k, n_seg, w.tile_prefix, w.state);
    GQ_LAUNCH_CHECK("topk_setup");
    topk_hist_kernel<0><<<grid, 256, 0, st>>>(grad, seg_start, w.tile_prefix, n_seg, w.state, w.hist);
    topk_scan_kernel<0><<<n_seg, 256, 0, st>>>(w.hist, w.state);
    topk_hist_kernel<1><<<grid, 256, 0, st>>>(grad, seg_start, w.tile_prefix, n_seg, w.state, w.hist);
    topk_scan_kernel<1><<<n_seg, 256, 0, st>>>(w.hist, w.state);
    topk_hist_kernel<2><<<grid, 256, 0, st>>>(grad, seg_start, w.tile_prefix, n_seg, w.state, w.hist);
    topk_scan_kernel<2><<<n_seg, 256, 0, st>>>(w.hist, w.state);
    GQ_LAUNCH_CHECK("topk_hist/scan");
    topk_count_kernel<<<grid, 256, 0, st>>>(grad, seg_start, w.tile_prefix, n_seg, w.state, w.tile_gt,
                                            w.tile_eq);
    topk_tilescan_kernel<<<n_seg, 256, 0, st>>>(w.tile_prefix, w.tile_gt, w.tile_eq);
    topk_write_kernel<<<grid, 256, 0, st>>>(grad, seg_start, k_prefix, w.tile_prefix, n_seg, w.state,
                                            w.tile_gt, w.tile_eq, out_dense, out_idx, out_val);
    GQ_LAUNCH_CHECK("topk_count/write");
    return GQ_OK;
}

// ------------------------------------------------------- scatter + reduce ---
__global__ void __launch_bounds__(256)
fill_kernel(float *__restrict__ out, int64_t n, float val)
{
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
        out[i] = val;
}
// one user's sparse entries: indices are unique inside a user, so plain RMW is race-free
__global__ void __launch_bounds__(256)
scatter_add_kernel(const int32_t *__restrict__ idx, const float *__restrict__ val, int64_t k,
                   float *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < k; i += (int64_t)gridDim.x * 256) {
        const int32_t j = idx[i];
        out[j] = __fadd_rn(out[j], val[i]);
    }
}
__global__ void __launch_bounds__(256)
scale_div_kernel(float *__restrict__ out, int64_t n, float denom)
{
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
        out[i] = __fdiv_rn(out[i], denom);
}

static int grid_for(int64_t n)
{
    int64_t blocks = (n + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 8;
    int64_t g = blocks < cap ? blocks : cap;
    return (int)(g < 1 ? 1 : g);
}

int topk_scatter_reduce(const int32_t *idx, const float *val, int64_t user_stride, int n_users,
                        int64_t k_total, int64_t n, int mean, int accumulate, float *out, cudaStream_t st)
{
    if (n == 0) return GQ_OK;
    // mean with accumulate: out + (sum)/U needs a scratch-free formulation only when U == 1
    if (!accumulate) {
        fill_kernel<<<grid_for(n), 256, 0, st>>>(out, n, 0.0f);
    }
    for (int u = 0; u < n_users; ++u) {
        const int32_t *iu = reinterpret_cast<const int32_t *>(reinterpret_cast<const char *>(idx) + u * user_stride);
        const float *vu = reinterpret_cast<const float *>(reinterpret_cast<const char *>(val) + u * user_stride);
        if (k_total > 0) scatter_add_kernel<<<grid_for(k_total), 256, 0, st>>>(iu, vu, k_total, out);
    }
    if (mean && n_users > 1) scale_div_kernel<<<grid_for(n), 256, 0, st>>>(out, n, (float)n_users);
    GQ_LAUNCH_CHECK("topk_scatter_reduce");
    return GQ_OK;
}

}  // namespace gq

using namespace gq;

extern "C" {

size_t gq_topk_workspace_bytes(int64_t n, int n_seg) { return topk_workspace_bytes(n, n_seg > 0 ? n_seg : 1); }

int gq_topk_select(const float *grad, int64_t n, const int64_t *seg_start, const int64_t *k,
                   const int64_t *k_prefix, int n_seg, float *out_dense, int32_t *out_idx, float *out_val,
                   void *workspace, size_t workspace_bytes, gq_stream_t stream)
{
    GQ_REQUIRE(n >= 0 && n_seg >= 1, "bad sizes");
    GQ_REQUIRE(n < ((int64_t)1 << 31), "top-k group larger than 2^31 elements");
    GQ_REQUIRE(grad && seg_start && k, "null pointer");
    GQ_REQUIRE(out_dense || (out_idx && out_val), "need dense or (idx, val) outputs");
    GQ_REQUIRE(!(out_idx || out_val) || k_prefix, "wire output needs k_prefix");
    if (!workspace || workspace_bytes < topk_workspace_bytes(n, n_seg)) {
        set_error("workspace too small: %zu < %zu", workspace_bytes, topk_workspace_bytes(n, n_seg));
        return GQ_ERR_WORKSPACE;
    }
    return topk_select(grad, n, seg_start, k, k_prefix, n_seg, out_dense, out_idx, out_val, workspace,
                       as_stream(stream));
}

int gq_topk_scatter_reduce(const int32_t *idx, const float *val, int64_t user_stride_bytes, int n_users,
                           int64_t k_total, int64_t n, int mean, int accumulate, float *out,
                           gq_stream_t stream)
{
    GQ_REQUIRE(n >= 0 && k_total >= 0 && n_users >= 1, "bad sizes");
    GQ_REQUIRE(!(mean && accumulate && n_users > 1), "mean + accumulate is not defined for top-k scatter");
    GQ_REQUIRE(k_total == 0 || (idx && val), "null pointer");
    return topk_scatter_reduce(idx, val, user_stride_bytes, n_users, k_total, n, mean, accumulate, out,
                               as_stream(stream));
}

}  // extern "C"
