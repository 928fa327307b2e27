// mbar_latency.cu -- how long after `mbarrier.arrive` (issued by one warp) does a spinning waiter in another
// warp of the same CTA see the phase flip?  Compared with a plain shared-memory flag (st.volatile / ld.volatile
// spin).  Measured idle and with the other warps of the CTA saturating the issue slots / with global stores in
// flight before the arrive (the release semantics of the arrive).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o mbar_latency tools/mbar_latency.cu && ./mbar_latency
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>   // 0: mbarrier arrive / test_wait, 1: mbarrier arrive / try_wait, 2: st.shared flag / ld.shared spin
__global__ void __launch_bounds__(512, 1) lat_kernel(long long *out, float *sink, int load_warps, int stores_before, int iters, int blockers)
{
    __shared__ uint64_t bar[3];
    __shared__ volatile uint32_t flag[2];
    __shared__ long long t_arrive[64], t_seen[64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[2])));
        flag[0] = 0; flag[1] = 0;
    }
    __syncthreads();
    if (warp == 0) {          // arriver
        for (int it = 0; it < iters; ++it) {
            // wait until the waiter has acknowledged the previous round (bar[1] / flag[1])
            if (it > 0) {
                if (MODE == 2) { while (flag[1] != (uint32_t)it) { } }
                else {
                    uint32_t ok = 0;
                    while (!ok) asm volatile("{.reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(smem_u32(&bar[1])), "r"((uint32_t)((it - 1) & 1)) : "memory");
                }
            }
            // some idle time so that the waiter is surely spinning
            long long t = clock64();
            while (clock64() - t < 2000) { }
            for (int k = 0; k < stores_before; ++k) sink[(blockIdx.x * 512 + threadIdx.x) * 8 + k + it * 4096 * 8] = (float)it;
            __syncwarp();
            if (lane == 0) {
                if (MODE == 2) { flag[0] = (uint32_t)(it + 1); }
                else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
                t_arrive[it & 63] = clock64();
            }
            __syncwarp();
        }
    } else if (warp == 1) {   // waiter
        for (int it = 0; it < iters; ++it) {
            if (lane == 0) {
                if (MODE == 2) { while (flag[0] != (uint32_t)(it + 1)) { } }
                else if (MODE == 0) {
                    uint32_t ok = 0;
                    while (!ok) asm volatile("{.reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(smem_u32(&bar[0])), "r"((uint32_t)(it & 1)) : "memory");
                } else {
                    uint32_t ok = 0;
                    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(smem_u32(&bar[0])), "r"((uint32_t)(it & 1)) : "memory");
                }
                t_seen[it & 63] = clock64();
                if (MODE == 2) flag[1] = (uint32_t)(it + 1);
                else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
            }
            __syncwarp();
        }
        if (lane == 0) out[1] = 1;
    } else if (warp >= 16 - blockers) {   // warps parked in a blocking try_wait (suspend-time hint) on a barrier that never completes
        volatile long long *done = out + 1;
        if (lane == 0) {
            for (int r = 0; r < 4000 && !*done; ++r) {
                uint32_t ok;
                asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(smem_u32(&bar[2])), "r"(0u), "r"(1000000u) : "memory");
            }
        }
        __syncwarp();
    } else if (warp - 2 < load_warps) {   // issue-slot hogs: independent FMAs until the waiter is done
        float a = threadIdx.x, b = 1.0001f, c = 0.5f, d = 0.25f, e = 2.0f;
        volatile long long *done = out + 1;
        for (int r = 0; r < 200000; ++r) {
#pragma unroll
            for (int k = 0; k < 64; ++k) { a = fmaf(a, b, c); d = fmaf(d, b, e); c = fmaf(c, b, a); e = fmaf(e, b, d); }
            if ((r & 15) == 0 && *done) break;
        }
        sink[threadIdx.x] = a + c + d + e;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long s = 0, mx = 0;
        const int n = iters < 64 ? iters : 64;
        for (int i = 0; i < n; ++i) { long long d = t_seen[i] - t_arrive[i]; s += d; if (d > mx) mx = d; }
        out[2] = s / n;
        out[3] = mx;
    }
}

int main()
{
    long long *out; float *sink;
    cudaMalloc(&out, 64); cudaMalloc(&sink, 512 * 8 * 4 * 4096 + 4096 * 8 * 4 * 64);
    const char *names[3] = {"mbarrier arrive -> test_wait spin", "mbarrier arrive -> try_wait spin ", "st.shared flag  -> ld.shared spin "};
    for (int mode = 0; mode < 3; ++mode)
        for (int load = 0; load <= 8; load += 8)
            for (int st = 0; st <= 4; st += 4)
                for (int blk = 0; blk <= 4; blk += 2) {
                    cudaMemset(out, 0, 64);
                    if (mode == 0) lat_kernel<0><<<1, 512>>>(out, sink, load, st, 64, blk);
                    if (mode == 1) lat_kernel<1><<<1, 512>>>(out, sink, load, st, 64, blk);
                    if (mode == 2) lat_kernel<2><<<1, 512>>>(out, sink, load, st, 64, blk);
                    cudaError_t e = cudaDeviceSynchronize();
                    long long h[8]; cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost);
                    printf("%s  busy warps %2d  stores before %d  warps parked in try_wait(hint) %d : mean %lld cycles  max %lld  (%s)\n", names[mode], load, st, blk, h[2], h[3], cudaGetErrorString(e));
                }
    return 0;
}
