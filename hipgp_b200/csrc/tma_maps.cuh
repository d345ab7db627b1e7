// tma_maps.cuh -- tiled TMA (cp.async.bulk.tensor) for the column pass: tensor-map construction on the host and the 3-D box
// load on the device.  The driver entry point cuTensorMapEncodeTiled is looked up at run time (cudaGetDriverEntryPoint), so
// the library links against the runtime only.  In the CPU emulation (HIPGP_EMU) a tensor map is a plain descriptor and the
// load is a synchronous strided copy with zero fill outside the tensor -- same coordinates, same box, same layout.
#pragma once
#include "lane_fft.cuh"
#ifndef HIPGP_EMU
#include <cuda.h>
#endif

namespace hipgp {

#ifdef HIPGP_EMU
struct alignas(64) TensorMap3 {
    const unsigned char* base; unsigned long long dim[3]; unsigned long long stride[2]; unsigned box[3]; unsigned esize; int valid;
};
#define HIPGP_GRID_CONSTANT
#else
typedef CUtensorMap TensorMap3;
#define HIPGP_GRID_CONSTANT __grid_constant__
#endif

struct alignas(64) ColsTmaMaps { TensorMap3 in; TensorMap3 spec; };

// rank-3 tiled map over `esize`-byte elements: dims (fastest first), byte strides of dims 1 and 2, box extents
static inline bool encode_map3(TensorMap3* out, const void* base, int esize, const unsigned long long dim[3],
                               const unsigned long long stride_bytes[2], const unsigned box[3]) {
#ifdef HIPGP_EMU
    out->base = (const unsigned char*)base; out->esize = (unsigned)esize; out->valid = 1;
    for (int i = 0; i < 3; ++i) { out->dim[i] = dim[i]; out->box[i] = box[i]; }
    out->stride[0] = stride_bytes[0]; out->stride[1] = stride_bytes[1];
    return true;
#else
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    static bool looked = false;
    if (!looked) {
        looked = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeFn)p;
    }
    if (!fn) return false;
    if (((uintptr_t)base % 16) || (stride_bytes[0] % 16) || (stride_bytes[1] % 16)) return false;
    cuuint64_t gd[3] = {dim[0], dim[1], dim[2]};
    cuuint64_t gs[2] = {stride_bytes[0], stride_bytes[1]};
    cuuint32_t bx[3] = {box[0], box[1], box[2]};
    cuuint32_t es[3] = {1, 1, 1};
    const CUtensorMapDataType dt = esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    const CUresult r = fn(out, dt, 3, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
#endif
}

// one box: elements (c0.., c1.., c2..) of the tensor -> dense [box2][box1][box0] block at smem_dst; completion on `bar`
__device__ __forceinline__ void tma_load_box3(void* smem_dst, const TensorMap3* map, int c0, int c1, int c2, unsigned long long* bar) {
#ifdef HIPGP_EMU
    (void)bar;
    unsigned char* d = (unsigned char*)smem_dst;
    for (unsigned z = 0; z < map->box[2]; ++z)
        for (unsigned y = 0; y < map->box[1]; ++y)
            for (unsigned x = 0; x < map->box[0]; ++x) {
                const long long gx = c0 + (long long)x, gy = c1 + (long long)y, gz = c2 + (long long)z;
                unsigned char* dst = d + (((size_t)z * map->box[1] + y) * map->box[0] + x) * map->esize;
                if (gx >= 0 && gy >= 0 && gz >= 0 && (unsigned long long)gx < map->dim[0] && (unsigned long long)gy < map->dim[1] && (unsigned long long)gz < map->dim[2])
                    std::memcpy(dst, map->base + (size_t)gz * map->stride[1] + (size_t)gy * map->stride[0] + (size_t)gx * map->esize, map->esize);
                else
                    std::memset(dst, 0, map->esize);
            }
#else
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2),
                   "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
#endif
}

// mbarrier helpers usable from both builds (the emulation turns a wait into a CTA-wide rendezvous: every thread of the CTA
// waits exactly once per phase in the kernels that use them, and the issuing thread copies synchronously before it waits)
__device__ __forceinline__ void tma_bar_init(unsigned long long* bar) {
#ifndef HIPGP_EMU
    mbar_init(bar, 1);
#else
    (void)bar;
#endif
}
__device__ __forceinline__ void tma_bar_expect(unsigned long long* bar, unsigned bytes) {
#ifndef HIPGP_EMU
    mbar_arrive_expect_tx(bar, bytes);
#else
    (void)bar; (void)bytes;
#endif
}
__device__ __forceinline__ void tma_bar_wait(unsigned long long* bar, unsigned parity, int emu_id, int emu_count) {
#ifndef HIPGP_EMU
    (void)emu_id; (void)emu_count;
    mbar_wait(bar, parity);
#else
    (void)bar; (void)parity;
    emu::g_named_barrier[emu_id].sync(emu_count);
#endif
}
// plain arrival-count barrier between one lane per warp and ONE waiting thread (which is itself one of the arriving lanes):
// the other lanes arrive without waiting.  Emulation: named barrier `emu_id`, arrive / sync of `count` participants.
__device__ __forceinline__ void warps_bar_init(unsigned long long* bar, unsigned count) {
#ifndef HIPGP_EMU
    mbar_init(bar, count);
#else
    (void)bar; (void)count;
#endif
}
__device__ __forceinline__ void warps_bar_arrive(unsigned long long* bar, int emu_id, int count) {
#ifndef HIPGP_EMU
    (void)emu_id; (void)count;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
#else
    (void)bar;
    emu::g_named_barrier[emu_id].arrive(count);
#endif
}
__device__ __forceinline__ void warps_bar_arrive_and_wait(unsigned long long* bar, unsigned parity, int emu_id, int count) {
#ifndef HIPGP_EMU
    (void)emu_id; (void)count;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
    mbar_wait(bar, parity);
#else
    (void)bar; (void)parity;
    emu::g_named_barrier[emu_id].sync(count);
#endif
}
__device__ __forceinline__ void tma_fence_before_issue() {
#ifndef HIPGP_EMU
    fence_proxy_async();
#endif
}

}  // namespace hipgp
