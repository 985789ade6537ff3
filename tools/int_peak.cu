// int_peak.cu -- INT32 throughput microbenchmark for the DP roofline:
// dependent-free chains of the instructions the NW kernel is made of, 8 independent
// chains per thread, full-chip launch.  Prints warp-instruction rates in Gop/s
// (lane-ops/s) per instruction class.   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CHAINS 8
#define UNROLL 16

template <int OP>
__device__ __forceinline__ void step(int &a, int b, int c) {
    if (OP == 0) asm volatile("add.s32 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == 1) asm volatile("max.s32 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == 2) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == 3) asm volatile("xor.b32 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == 4) asm volatile("{ .reg .pred p; setp.gt.s32 p, %0, %1; selp.s32 %0, %2, %0, p; }" : "+r"(a) : "r"(b), "r"(c));
    if (OP == 5) a = __viaddmax_s32(a, b, c);
    if (OP == 6) a = __vimax3_s32(a, b, c);
    // float forms of the same recurrence pieces (scores are small integers, exact in fp32)
    if (OP == 7) asm volatile("{ .reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %2, %0, p; }" : "+f"(*(float *)&a) : "f"(*(float *)&b), "f"(*(float *)&c));
    if (OP == 8) asm volatile("max.f32 %0, %0, %1;" : "+f"(*(float *)&a) : "f"(*(float *)&b));
    if (OP == 9) asm volatile("add.f32 %0, %0, %1;" : "+f"(*(float *)&a) : "f"(*(float *)&b));
}

template <int OP, int OP2>
__global__ void __launch_bounds__(256) bench(int *out, int iters, int b, int c) {
    int a[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; k++) a[k] = threadIdx.x + k;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
#pragma unroll
            for (int k = 0; k < CHAINS; k++) {
                // operands rotate across chains so that ptxas cannot fold consecutive steps
                if (k & 1) step<OP2>(a[k], a[(k + 1) % CHAINS], a[(k + 3) % CHAINS]);
                else step<OP>(a[k], a[(k + 1) % CHAINS], a[(k + 3) % CHAINS]);
            }
        }
    }
    int s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) s += a[k];
    if (s == 0x7fffffff) out[0] = s;
}

template <int OP, int OP2>
static void run(const char *name, int instr_per_step, int *d, int sm) {
    const int iters = 2000, grid = sm * 8, threads = 256;
    bench<OP, OP2><<<grid, threads>>>(d, 10, 3, 5);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        bench<OP, OP2><<<grid, threads>>>(d, iters, 3, 5);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double ops = (double)grid * threads * iters * UNROLL * CHAINS * instr_per_step;
    printf("{\"op\": \"%s\", \"gops\": %.1f, \"ms\": %.3f, \"lane_ops_per_clk_per_sm_at_1965MHz\": %.1f}\n", name,
           ops / best * 1e-6, best, ops / (best * 1e-3) / 1.965e9 / sm);
}

int main() {
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    int *d; cudaMalloc(&d, 4);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, p.clockRate);
    run<0, 0>("iadd", 1, d, p.multiProcessorCount);
    run<1, 1>("imnmx", 1, d, p.multiProcessorCount);
    run<2, 2>("imad", 1, d, p.multiProcessorCount);
    run<3, 3>("lop3", 1, d, p.multiProcessorCount);
    run<4, 4>("setp+selp", 2, d, p.multiProcessorCount);
    run<5, 5>("viaddmax", 1, d, p.multiProcessorCount);
    run<6, 6>("vimax3", 1, d, p.multiProcessorCount);
    run<7, 7>("fsetp+fsel", 2, d, p.multiProcessorCount);
    run<8, 8>("fmnmx", 1, d, p.multiProcessorCount);
    run<9, 9>("fadd", 1, d, p.multiProcessorCount);
    run<7, 4>("fsetp+fsel | isetp+sel 1:1", 2, d, p.multiProcessorCount);
    run<0, 2>("iadd+imad 1:1", 1, d, p.multiProcessorCount);
    run<1, 2>("imnmx+imad 1:1", 1, d, p.multiProcessorCount);
    run<0, 1>("iadd+imnmx 1:1", 1, d, p.multiProcessorCount);
    return 0;
}
