// Throughput microbenchmark of the ALU-pipe instructions the HSQ epilogue is made of (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_microbench alu_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(float *out, int iters, float seed)
{
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed * (threadIdx.x + i + 1);
    float t = seed;
    unsigned m = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (OP == 0) a[i] = fmaxf(fabsf(a[i]), fabsf(t));                                   // FMNMX |.|
            if (OP == 1) a[i] = fmaxf(fmaxf(fabsf(a[i]), fabsf(a[(i + 1) & 15])), fabsf(t));    // FMNMX3
            if (OP == 2) a[i] = __uint_as_float(__funnelshift_l(__float_as_uint(a[i]), __float_as_uint(t), 1));  // SHF
            if (OP == 3) a[i] = a[i] + t;                                                       // FADD
            if (OP == 4) a[i] = __uint_as_float((__float_as_uint(a[i]) & 0x7fffff00u) | (unsigned)i);  // LOP3
            if (OP == 5) a[i] = fmaf(a[i], t, a[(i + 1) & 15]);                                 // FFMA
        }
        t += 1e-9f;
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (threadIdx.x == 0 && blockIdx.x == 0) out[OP] = (float)(t1 - t0) / ((float)iters * 16.0f);
    if (s == 123.456f) out[100] = s + m;
}

int main()
{
    float *d;
    cudaMalloc(&d, 1024);
    const char *names[] = {"FMNMX |a|,|b|", "FMNMX3", "SHF funnel", "FADD", "LOP3", "FFMA"};
    for (int warps = 1; warps <= 16; warps *= 2) {
        printf("warps/SM-block = %d (one block on one SM; cycles per warp-instruction per warp)\n", warps);
        k<0><<<1, 32 * warps>>>(d, 2000, 1.0001f);
        k<1><<<1, 32 * warps>>>(d, 2000, 1.0001f);
        k<2><<<1, 32 * warps>>>(d, 2000, 1.0001f);
        k<3><<<1, 32 * warps>>>(d, 2000, 1.0001f);
        k<4><<<1, 32 * warps>>>(d, 2000, 1.0001f);
        k<5><<<1, 32 * warps>>>(d, 2000, 1.0001f);
        cudaDeviceSynchronize();
        float h[8];
        cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        for (int i = 0; i < 6; ++i)
            printf("   %-14s %.2f clk/instr/warp  -> %.2f warp-instr/clk/SMSP\n", names[i], h[i],
                   (warps >= 4 ? warps / 4.0f : 1.0f) / h[i]);
    }
    return 0;
}
