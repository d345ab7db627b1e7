// Micro-benchmark 2 (developer tool): the column kernel's per-thread section structure in isolation.
//   section = 16 x LDS.128 -> 224 packed FMAs that DEPEND on the loaded registers (butterfly-like network) -> 16 x STS.128
// Variants: plain | half of the warps delayed at start | two half-size items per thread, software-pipelined
// (loads of item B in flight while item A computes, stores of A in flight while B computes) | __syncwarp between sections.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void fma2(u64& d, u64 a, u64 b, u64 c) { asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); }

struct __align__(16) L4 { u64 a, b; };

// NP packed registers per item (NP/2 LDS.128), ROUNDS rounds of NP dependent packed FMAs
template <int NP, int ROUNDS>
__device__ __forceinline__ void compute(u64 (&p)[NP], u64 c) {
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        u64 q[NP];
#pragma unroll
        for (int k = 0; k < NP; ++k) fma2(q[k], p[k], c, p[k ^ (1 << (r % 4))]);
#pragma unroll
        for (int k = 0; k < NP; ++k) p[k] = q[k];
    }
}

// MODE 0: one item of 32 packed regs (16 LDS, 7 rounds = 224 FMAs, 16 STS)
// MODE 1: two items of 16 packed regs each (8 LDS, 7 rounds = 112 FMAs, 8 STS), software-pipelined
template <int MODE, bool SYNCW>
__global__ void __launch_bounds__(512, 1) k_sec(float* out, int iters, float seed, int delay_mask, int delay_ns) {
    extern __shared__ L4 sm[];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < 16 * nt; i += nt) { sm[i].a = pk(i, 1); sm[i].b = pk(2, i); }
    __syncthreads();
    const u64 c = pk(seed, seed);
    if (((tid >> 5) & delay_mask) && delay_ns) __nanosleep(delay_ns);
    if (MODE == 0) {
        u64 p[32];
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { const L4 v = sm[tid + nt * i]; p[2 * i] = v.a; p[2 * i + 1] = v.b; }
            compute<32, 7>(p, c);
#pragma unroll
            for (int i = 0; i < 16; ++i) { L4 v; v.a = p[2 * i]; v.b = p[2 * i + 1]; sm[tid + nt * i] = v; }
            if (SYNCW) __syncwarp();
        }
        if (p[0] == 12345ull) out[0] = 1.f;
    } else {
        u64 pa[16], pb[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) { const L4 v = sm[tid + nt * i]; pa[2 * i] = v.a; pa[2 * i + 1] = v.b; }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { const L4 v = sm[tid + nt * (8 + i)]; pb[2 * i] = v.a; pb[2 * i + 1] = v.b; }   // B's loads
            compute<16, 7>(pa, c);                                                                                           // A computes
#pragma unroll
            for (int i = 0; i < 8; ++i) { L4 v; v.a = pa[2 * i]; v.b = pa[2 * i + 1]; sm[tid + nt * i] = v; }               // A's stores
            if (SYNCW) __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) { const L4 v = sm[tid + nt * i]; pa[2 * i] = v.a; pa[2 * i + 1] = v.b; }           // A's next loads
            compute<16, 7>(pb, c);                                                                                           // B computes
#pragma unroll
            for (int i = 0; i < 8; ++i) { L4 v; v.a = pb[2 * i]; v.b = pb[2 * i + 1]; sm[tid + nt * (8 + i)] = v; }         // B's stores
            if (SYNCW) __syncwarp();
        }
        if (pa[0] == 12345ull || pb[0] == 12345ull) out[0] = 1.f;
    }
}

template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    float* out; cudaMalloc(&out, 4);
    int sms = 0, khz = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double clk = khz * 1e3; const int iters = 4000;
    auto run = [&](const char* nm, auto kern, int nthr, int mask, int ns) {
        const size_t smem = (size_t)16 * nthr * 16;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        float ms = timeit([&] { kern<<<sms, nthr, smem>>>(out, iters, 1.0001f, mask, ns); });
        const double cyc = ms * 1e-3 * clk / iters;
        const double fp_floor = 224.0 * (nthr / 32) / 4 * 2.05;
        printf("%-58s threads %3d: %7.1f cycles / section / SM  (FP2 floor %.0f, ratio %.2f)\n", nm, nthr, cyc, fp_floor, cyc / fp_floor);
    };
    for (int nthr : {256, 512}) {
        run("1 item (16 LDS -> 224 dependent FP2 -> 16 STS)", k_sec<0, false>, nthr, 0, 0);
        run("1 item + __syncwarp", k_sec<0, true>, nthr, 0, 0);
        run("1 item, warps 8-15 delayed 400 ns", k_sec<0, false>, nthr, 8, 400);
        run("1 item, warps 4-7,12-15 delayed 400 ns", k_sec<0, false>, nthr, 4, 400);
        run("2 half items, software-pipelined", k_sec<1, false>, nthr, 0, 0);
        run("2 half items, software-pipelined + __syncwarp", k_sec<1, true>, nthr, 0, 0);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
