// tc_ptx.cuh -- thin inline-PTX wrappers shared by the tcgen05 kernels (sm_100a): mbarrier, TMA,
// tcgen05.mma / commit / ld / fences, the K-major SWIZZLE_64B shared-memory operand descriptor.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace gq {
namespace tcptx {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ uint32_t mbar_test_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
// bounded wait: a protocol bug traps (sticky error, process exits) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, bool backoff = false)
{
    // back off between polls so that waiting warps do not steal issue/ALU slots from the
    // epilogue warps sharing their scheduler
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (backoff) __nanosleep(40);
        if (++spins > (1u << 22)) __trap();
    }
}
// same, with a suspend-time hint: the warp sleeps inside try_wait (up to ~hint_ns) instead of
// spinning through the issue slots of the warps that share its scheduler; still bounded -> trap
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity)
{
    unsigned long long t0 = 0ull;
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
        if (ok) break;
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0ull) t0 = now;
        else if (now - t0 > 4000000000ull) __trap();   // 4 s: a protocol bug, not a slow tile
    }
}

// Latency-critical waits (TMEM buffer turnaround): plain try_wait, which blocks in hardware for a
// short, implementation-defined time and wakes ~60 cycles after the completing arrive -- the
// suspend-hint form above sleeps through NANOSLEEP.SYNCS and was measured to resume 250-750 cycles
// after the arrive (tests/tc2_trace.py).  Bounded by wall clock -> trap.
__device__ __forceinline__ void mbar_wait_hw(uint32_t bar, uint32_t parity)
{
    uint32_t spins = 0;
    unsigned long long t0 = 0ull;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 1023u) == 0u) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0ull) t0 = now;
            else if (now - t0 > 4000000000ull) __trap();
        }
    }
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_64B operand: rows of 64 B, 8-row groups 512 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;           // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(512 >> 4) << 32;  // stride byte offset
    d |= (uint64_t)1 << 46;           // descriptor version (Blackwell)
    d |= (uint64_t)4 << 61;           // SWIZZLE_64B
    return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t accumulate,
                                         uint32_t idesc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(taddr), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// wait for all outstanding TMEM loads; the registers of the load being waited for pass through
// the asm, so that none of their uses can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                   "+r"(r[15])
                 :: "memory");
}

__device__ __forceinline__ float absmax3(float m, uint32_t a, uint32_t b)
{
    return fmaxf(m, fmaxf(fabsf(__uint_as_float(a)), fabsf(__uint_as_float(b))));
}

}  // namespace tcptx
}  // namespace gq
