// cuda_emu.h -- TEST-ONLY host emulation of the small CUDA subset the kernels in this directory use.
//
// Purpose: the build container has nvcc but no GPU.  Compiling the very same kernel sources with
// g++ -DHIPGP_EMU lets `tests/` (marker "not gpu") execute every kernel's index arithmetic, barriers and
// reductions on the CPU and compare with the oracle before GPU minutes are spent.  One OS thread per
// CUDA thread, blocks run one after another, __syncthreads() is a real barrier.
//
// This is NOT a product path: `hipgp_b200/_lib.py` only ever loads the nvcc-built libhipgp_b200.so and
// raises if it is missing; the emulation library is built into tests/_emu/ by tests/emu_build.py and is
// loaded by tests only.
#pragma once
#ifdef HIPGP_EMU
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)

struct uint3e { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct double2 { double x, y; };
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline double2 make_double2(double a, double b) { return double2{a, b}; }

typedef void* cudaStream_t;
typedef void* cudaEvent_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };

namespace emu {
struct Barrier {
    std::mutex mu;
    std::condition_variable cv;
    int expected = 0, arrived = 0;
    unsigned long gen = 0;
    void reset(int n) { expected = n; arrived = 0; }
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        unsigned long g = gen;
        if (++arrived >= expected) { arrived = 0; ++gen; cv.notify_all(); return; }
        cv.wait(lk, [&] { return gen != g; });
    }
    void drop() {   // a thread that returned from the kernel no longer takes part in __syncthreads
        std::unique_lock<std::mutex> lk(mu);
        --expected;
        if (expected > 0 && arrived >= expected) { arrived = 0; ++gen; cv.notify_all(); }
    }
};
// named barriers (bar.sync id, count / bar.arrive id, count): a counting barrier whose arrivals need not wait
struct NamedBarrier {
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0;
    unsigned long gen = 0;
    void reset() { arrived = 0; }
    void arrive(int expected) {
        std::unique_lock<std::mutex> lk(mu);
        if (++arrived >= expected) { arrived = 0; ++gen; cv.notify_all(); }
    }
    void sync(int expected) {
        std::unique_lock<std::mutex> lk(mu);
        const unsigned long g = gen;
        if (++arrived >= expected) { arrived = 0; ++gen; cv.notify_all(); return; }
        cv.wait(lk, [&] { return gen != g; });
    }
};
extern NamedBarrier g_named_barrier[16];
extern thread_local dim3 t_threadIdx, t_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern Barrier g_block_barrier;
extern Barrier g_warp_barrier[64];
extern unsigned char* g_dyn_smem;
extern double g_shfl_scratch[64][32][2];

template <class F>
void launch(dim3 grid, dim3 block, size_t smem, F&& body) {
    const int T = (int)(block.x * block.y * block.z);
    g_blockDim = block; g_gridDim = grid;
    std::vector<unsigned char> dyn(smem + 64);
    g_dyn_smem = dyn.data() + (64 - ((uintptr_t)dyn.data() & 63)) % 64;
    Barrier start, finish;
    const long nblocks = (long)grid.x * grid.y * grid.z;
    start.reset(T); finish.reset(T);
    auto worker = [&](int t) {
        t_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
        for (long b = 0; b < nblocks; ++b) {
            t_blockIdx = dim3((unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((long)grid.x * grid.y)));
            if (t == 0) {
                g_block_barrier.reset(T);
                for (int w = 0; w < 64; ++w) g_warp_barrier[w].reset(32);
                for (int w = 0; w < 16; ++w) g_named_barrier[w].reset();
            }
            start.wait();
            body();
            g_block_barrier.drop();
            finish.wait();
        }
    };
    std::vector<std::thread> th;
    th.reserve(T);
    for (int t = 1; t < T; ++t) th.emplace_back(worker, t);
    worker(0);
    for (auto& x : th) x.join();
    g_dyn_smem = nullptr;
}
}  // namespace emu

#define threadIdx emu::t_threadIdx
#define blockIdx emu::t_blockIdx
#define blockDim emu::g_blockDim
#define gridDim emu::g_gridDim
static inline void __syncthreads() { emu::g_block_barrier.wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int lanemask) {
    const int tid = (int)(threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z));
    const int w = tid >> 5, l = tid & 31;
    static_assert(sizeof(T) <= 16, "shfl payload");
    std::memcpy(emu::g_shfl_scratch[w][l], &v, sizeof(T));
    emu::g_warp_barrier[w].wait();
    T r;
    std::memcpy(&r, emu::g_shfl_scratch[w][l ^ lanemask], sizeof(T));
    emu::g_warp_barrier[w].wait();
    return r;
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, int delta) {
    const int tid = (int)(threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z));
    const int w = tid >> 5, l = tid & 31;
    std::memcpy(emu::g_shfl_scratch[w][l], &v, sizeof(T));
    emu::g_warp_barrier[w].wait();
    T r;
    std::memcpy(&r, emu::g_shfl_scratch[w][(l + delta) < 32 ? l + delta : l], sizeof(T));
    emu::g_warp_barrier[w].wait();
    return r;
}
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicInc_emu(unsigned* p) { return __atomic_fetch_add(p, 1u, __ATOMIC_SEQ_CST); }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int __popc(int x) { return __builtin_popcount((unsigned)x); }
static inline void sincospi(double x, double* s, double* c) { *s = std::sin(M_PI * x); *c = std::cos(M_PI * x); }
static inline double cospi(double x) { return std::cos(M_PI * x); }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }

// ---- tiny runtime ----
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }

#define HIPGP_DYN_SMEM(name) unsigned char* name = emu::g_dyn_smem
#define HIPGP_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#define HIPGP_SET_MAX_SMEM(kernel, bytes) ((void)0)

#else  // ---------------------------------------------------------------- real CUDA
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <vector>

#define HIPGP_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]   /* 128: TMA tensor boxes land in it */
#define HIPGP_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define HIPGP_SET_MAX_SMEM(kernel, bytes) \
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
#endif
