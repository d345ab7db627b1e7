// Micro-benchmark (developer tool): issue / pipe rates of packed f32x2 arithmetic vs scalar fp32 on sm_100a, alone and
// mixed with 128-bit shared-memory traffic.  Answers: how many cycles does an FFMA2 / FADD2 occupy the FMA pipe of an SM
// sub-partition, and do LDS.128 / STS.128 overlap with it?   nvcc -arch=sm_100a -O3 -o ubench_fp2 ubench_fp2.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ILP 8
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }

template <int MODE>   // 0: fma.f32x2   1: add.f32x2   2: scalar fma (2 per packed op)   3: scalar add
__global__ void __launch_bounds__(512, 1) k_arith(float* out, int iters, float seed) {
    unsigned long long x[ILP], a = pk(seed, seed * 0.5f), b = pk(0.25f, 0.125f);
    float xs[2 * ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = pk(threadIdx.x + i, i); xs[2 * i] = threadIdx.x + i; xs[2 * i + 1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(a), "l"(b));
            else if (MODE == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x[i]) : "l"(b));
            else if (MODE == 2) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(xs[2 * i]) : "f"(seed), "f"(0.25f)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(xs[2 * i + 1]) : "f"(seed), "f"(0.25f)); }
            else { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(xs[2 * i]) : "f"(0.25f)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(xs[2 * i + 1]) : "f"(0.25f)); }
        }
    }
    float acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); acc += lo + hi + xs[2 * i] + xs[2 * i + 1]; }
    if (acc == 12345.678f) out[0] = acc;
}

// per iteration: NLS LDS.128 + NLS STS.128 (conflict-free, 16 B per thread) and NFP packed FMAs on independent registers
template <int NLS, int NFP>
__global__ void __launch_bounds__(512, 1) k_mix(float* out, int iters, float seed) {
    extern __shared__ float4 sm[];
    unsigned long long x[ILP], a = pk(seed, seed * 0.5f), b = pk(0.25f, 0.125f);
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = pk(threadIdx.x + i, i);
    float4 v[NLS > 0 ? NLS : 1];
    for (int i = 0; i < (NLS > 0 ? NLS : 1); ++i) v[i] = make_float4(threadIdx.x, i, 1, 2);
    for (int i = threadIdx.x; i < 8 * 512; i += 512) sm[i] = make_float4(i, 0, 0, 0);
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NLS; ++i) v[i] = sm[threadIdx.x + 512 * ((i + it) & 7)];
#pragma unroll
        for (int k = 0; k < NFP / ILP; ++k) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(a), "l"(b));
        }
#pragma unroll
        for (int i = 0; i < NLS; ++i) { v[i].x += 1.f; sm[threadIdx.x + 512 * ((i + it + 3) & 7)] = v[i]; }
    }
    float acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); acc += lo + hi; }
    for (int i = 0; i < (NLS > 0 ? NLS : 1); ++i) acc += v[i].x;
    if (acc == 12345.678f) out[0] = acc;
}

template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    float* out; cudaMalloc(&out, 4);
    int dev = 0, sms = 0, khz = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const double clk = khz * 1e3;
    const int iters = 20000;
    printf("SMs %d, clock %.0f MHz (nominal max; rates below assume it)\n", sms, clk / 1e6);
    const char* names[4] = {"fma.f32x2", "add.f32x2", "fma.f32 x2", "add.f32 x2"};
    for (int nthr : {128, 256, 512}) {
        for (int mode = 0; mode < 4; ++mode) {
            float ms = 0;
            if (mode == 0) ms = timeit([&] { k_arith<0><<<sms, nthr>>>(out, iters, 1.0001f); });
            if (mode == 1) ms = timeit([&] { k_arith<1><<<sms, nthr>>>(out, iters, 1.0001f); });
            if (mode == 2) ms = timeit([&] { k_arith<2><<<sms, nthr>>>(out, iters, 1.0001f); });
            if (mode == 3) ms = timeit([&] { k_arith<3><<<sms, nthr>>>(out, iters, 1.0001f); });
            const double cyc = ms * 1e-3 * clk;
            const double warp_ops = (double)iters * ILP * (nthr / 32);                 // packed-op equivalents per SM
            printf("threads/SM %3d  %-11s: %.3f ms  -> %.2f cycles per packed-op-equivalent per SM sub-partition (4 per SM)\n", nthr, names[mode], ms,
                   cyc / (warp_ops / 4));
        }
    }
    const size_t smem = 8 * 512 * 16;
    auto run_mix = [&](const char* nm, auto kern, int nls, int nfp) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        float ms = timeit([&] { kern<<<sms, 512, smem>>>(out, 4000, 1.0001f); });
        const double cyc = ms * 1e-3 * clk / 4000;
        printf("mix %-22s: %.1f cycles / iteration / SM   (FP2 alone would need %.0f at 2 cyc/op/SMSP, smem alone %.0f at 128 B/clk)\n", nm, cyc,
               (double)nfp * 16 / 4 * 2, (double)nls * 2 * 16 * 4);
    };
    run_mix("16 LDS+16 STS, 0 FP2", k_mix<16, 0>, 16, 0);
    run_mix("0 LDS/STS, 224 FP2", k_mix<0, 224>, 0, 224);
    run_mix("16 LDS+16 STS, 224 FP2", k_mix<16, 224>, 16, 224);
    run_mix("16 LDS+16 STS, 112 FP2", k_mix<16, 112>, 16, 112);
    run_mix("8 LDS+8 STS, 224 FP2", k_mix<8, 224>, 8, 224);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
