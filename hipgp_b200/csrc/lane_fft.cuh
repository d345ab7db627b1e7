// lane_fft.cuh -- the register-blocked line-FFT engine behind the specialised kernel family (fast_kernels.cuh).
//
// Unit of work: a LANE = 16 bytes = the same position of two neighbouring fp32 lines, or of one fp64 line.
//   Lane<float>  = { (re_line0, re_line1), (im_line0, im_line1) }     Lane<double> = { re, im }
// fp32 arithmetic runs on Blackwell's packed FADD2 / FMUL2 / FFMA2 ACROSS the two lines of a lane, so a twiddle
// (one complex scalar, shared by both lines) is a broadcast operand and a complex multiply is 4 packed instructions
// for two lines with no register shuffling; multiplication by -+i is register renaming plus an operand negation.
//
// Shared memory holds a tile as  s[slot(pos) * NL + lane]  (NL lanes of one position are contiguous, 16 B each), so
//   * every shared access is one LDS.128 / STS.128,
//   * a thread owns (butterfly, lane); the NL lanes of a butterfly position are neighbouring threads,
//   * slot(pos) = pos + pos / R_last  (only when NL < 8) keeps the stride-R_last butterflies of the last stage on
//     distinct banks; every butterfly leg is then [thread base + compile-time immediate].
// A transform is the same DIF-forward / DIT-inverse pair as fft_engine.cuh (digit-reversed spectra, no permutation
// pass), with a compile-time radix list: first radix in {2,3,4,5,8,16}, later radices powers of two.
#pragma once
#include "conv_kernels.cuh"

namespace hipgp {

// ---------------------------------------------------------------------------------------------------------
#ifdef HIPGP_EMU
static inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
static inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y)}; }
#endif

template <class T> struct Lane;
template <> struct __align__(16) Lane<float> { float2 re, im; };
template <> struct __align__(16) Lane<double> { double re, im; };
template <class T> struct LaneInfo { static constexpr int LPT = sizeof(T) == 4 ? 2 : 1; };   // lines per lane

__device__ __forceinline__ float2 bc2(float s) { return make_float2(s, s); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }

template <class T> __device__ __forceinline__ Lane<T> lzero();
template <> __device__ __forceinline__ Lane<float> lzero<float>() { Lane<float> r; r.re = bc2(0.f); r.im = bc2(0.f); return r; }
template <> __device__ __forceinline__ Lane<double> lzero<double>() { Lane<double> r; r.re = 0.0; r.im = 0.0; return r; }

// ---- fp32 lanes: packed across the two lines ----
__device__ __forceinline__ Lane<float> operator+(Lane<float> a, Lane<float> b) { Lane<float> r; r.re = __fadd2_rn(a.re, b.re); r.im = __fadd2_rn(a.im, b.im); return r; }
__device__ __forceinline__ Lane<float> operator-(Lane<float> a, Lane<float> b) { Lane<float> r; r.re = __fadd2_rn(a.re, neg2(b.re)); r.im = __fadd2_rn(a.im, neg2(b.im)); return r; }
// a * w
__device__ __forceinline__ Lane<float> lmul(Lane<float> a, cplx<float> w) {
    Lane<float> r;
    r.re = __ffma2_rn(a.re, bc2(w.x), __fmul2_rn(a.im, bc2(-w.y)));
    r.im = __ffma2_rn(a.re, bc2(w.y), __fmul2_rn(a.im, bc2(w.x)));
    return r;
}
// a * conj(w)
__device__ __forceinline__ Lane<float> lmulc(Lane<float> a, cplx<float> w) {
    Lane<float> r;
    r.re = __ffma2_rn(a.re, bc2(w.x), __fmul2_rn(a.im, bc2(w.y)));
    r.im = __ffma2_rn(a.im, bc2(w.x), __fmul2_rn(a.re, bc2(-w.y)));
    return r;
}
__device__ __forceinline__ Lane<float> lscale(Lane<float> a, float s) { Lane<float> r; r.re = __fmul2_rn(a.re, bc2(s)); r.im = __fmul2_rn(a.im, bc2(s)); return r; }
// c + s * a
__device__ __forceinline__ Lane<float> lfma(Lane<float> a, float s, Lane<float> c) { Lane<float> r; r.re = __ffma2_rn(a.re, bc2(s), c.re); r.im = __ffma2_rn(a.im, bc2(s), c.im); return r; }
__device__ __forceinline__ Lane<float> lconj(Lane<float> a) { a.im = neg2(a.im); return a; }
template <bool INV> __device__ __forceinline__ Lane<float> lmi(Lane<float> a) {   // * (-i) forward, * (+i) inverse
    Lane<float> r;
    if (INV) { r.re = neg2(a.im); r.im = a.re; } else { r.re = a.im; r.im = neg2(a.re); }
    return r;
}
// per-line real / complex factors (spectrum multiply): s0 for line 0, s1 for line 1
__device__ __forceinline__ Lane<float> lmul_real2(Lane<float> a, float s0, float s1) {
    Lane<float> r; const float2 s = make_float2(s0, s1); r.re = __fmul2_rn(a.re, s); r.im = __fmul2_rn(a.im, s); return r;
}
template <bool CONJ> __device__ __forceinline__ Lane<float> lmul_cplx2(Lane<float> a, cplx<float> w0, cplx<float> w1) {
    const float2 wx = make_float2(w0.x, w1.x), wy = make_float2(CONJ ? -w0.y : w0.y, CONJ ? -w1.y : w1.y);
    Lane<float> r;
    r.re = __ffma2_rn(a.re, wx, __fmul2_rn(a.im, neg2(wy)));
    r.im = __ffma2_rn(a.re, wy, __fmul2_rn(a.im, wx));
    return r;
}

// ---- fp64 lanes: one line ----
__device__ __forceinline__ Lane<double> operator+(Lane<double> a, Lane<double> b) { Lane<double> r; r.re = a.re + b.re; r.im = a.im + b.im; return r; }
__device__ __forceinline__ Lane<double> operator-(Lane<double> a, Lane<double> b) { Lane<double> r; r.re = a.re - b.re; r.im = a.im - b.im; return r; }
__device__ __forceinline__ Lane<double> lmul(Lane<double> a, cplx<double> w) { Lane<double> r; r.re = a.re * w.x - a.im * w.y; r.im = a.re * w.y + a.im * w.x; return r; }
__device__ __forceinline__ Lane<double> lmulc(Lane<double> a, cplx<double> w) { Lane<double> r; r.re = a.re * w.x + a.im * w.y; r.im = a.im * w.x - a.re * w.y; return r; }
__device__ __forceinline__ Lane<double> lscale(Lane<double> a, double s) { Lane<double> r; r.re = a.re * s; r.im = a.im * s; return r; }
__device__ __forceinline__ Lane<double> lfma(Lane<double> a, double s, Lane<double> c) { Lane<double> r; r.re = c.re + s * a.re; r.im = c.im + s * a.im; return r; }
__device__ __forceinline__ Lane<double> lconj(Lane<double> a) { a.im = -a.im; return a; }
template <bool INV> __device__ __forceinline__ Lane<double> lmi(Lane<double> a) {
    Lane<double> r;
    if (INV) { r.re = -a.im; r.im = a.re; } else { r.re = a.im; r.im = -a.re; }
    return r;
}

// ---- line <-> lane conversion: line l (0 .. LPT-1) of a lane as an ordinary complex number ----
__device__ __forceinline__ cplx<float> lane_get(const Lane<float>& a, int l) { return l ? mk<float>(a.re.y, a.im.y) : mk<float>(a.re.x, a.im.x); }
__device__ __forceinline__ cplx<double> lane_get(const Lane<double>& a, int) { return mk<double>(a.re, a.im); }
__device__ __forceinline__ void lane_set(Lane<float>& a, int l, cplx<float> v) { if (l) { a.re.y = v.x; a.im.y = v.y; } else { a.re.x = v.x; a.im.x = v.y; } }
__device__ __forceinline__ void lane_set(Lane<double>& a, int, cplx<double> v) { a.re = v.x; a.im = v.y; }

// ---- frequency-workspace layout ------------------------------------------------------------------------------
// The specialised kernels keep the frequency workspace W in LANE layout: for fp32 the two neighbouring bins (2c, 2c+1)
// of a row are stored as (re_2c, re_2c+1, im_2c, im_2c+1), i.e. exactly one Lane<float>, so the column pass moves
// lanes with plain 16-byte accesses and no register shuffling; for fp64 a lane is one interleaved complex number.
// (The generic kernels of conv_kernels.cuh use interleaved complex for both, so an fp32 pipeline is either all
// specialised or all generic -- see run_pipeline.)
struct __align__(16) Raw16f { float a, b, c, d; };
// Streaming data (vectors, the frequency workspace, spectrum tiles) is moved with L2-only cache policy (ld/st.global.cg)
// so that it does not evict the small twiddle / pairing tables, which are read through L1 by every CTA of an SM.
// (.cg loads are coherent at L2: fine for the in-place column pass.)
#ifdef HIPGP_EMU
template <class V> __device__ __forceinline__ V ld_stream(const V* p) { return *p; }
template <class V> __device__ __forceinline__ void st_stream(V* p, V v) { *p = v; }
#else
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcg(p); }
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcg(p); }
__device__ __forceinline__ cplx<float> ld_stream(const cplx<float>* p) { const float2 t = __ldcg(reinterpret_cast<const float2*>(p)); return mk<float>(t.x, t.y); }
__device__ __forceinline__ cplx<double> ld_stream(const cplx<double>* p) { const double2 t = __ldcg(reinterpret_cast<const double2*>(p)); return mk<double>(t.x, t.y); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcg(p, v); }
__device__ __forceinline__ void st_stream(double* p, double v) { __stcg(p, v); }
__device__ __forceinline__ void st_stream(cplx<float>* p, cplx<float> v) { __stcg(reinterpret_cast<float2*>(p), make_float2(v.x, v.y)); }
__device__ __forceinline__ void st_stream(cplx<double>* p, cplx<double> v) { __stcg(reinterpret_cast<double2*>(p), make_double2(v.x, v.y)); }
#endif
// 16 bytes of reals (vector rows are streamed in 16-byte chunks by the row kernels)
template <class T> struct Vec16;
template <> struct __align__(16) Vec16<float> { float v[4]; };
template <> struct __align__(16) Vec16<double> { double v[2]; };
#ifdef HIPGP_EMU
template <class T> __device__ __forceinline__ Vec16<T> ldv_stream(const T* p) { return *reinterpret_cast<const Vec16<T>*>(p); }
template <class T> __device__ __forceinline__ void stv_stream(T* p, Vec16<T> v) { *reinterpret_cast<Vec16<T>*>(p) = v; }
#else
__device__ __forceinline__ Vec16<float> ldv_stream(const float* p) { const float4 t = __ldcg(reinterpret_cast<const float4*>(p)); Vec16<float> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r; }
__device__ __forceinline__ Vec16<double> ldv_stream(const double* p) { const double2 t = __ldcg(reinterpret_cast<const double2*>(p)); Vec16<double> r; r.v[0] = t.x; r.v[1] = t.y; return r; }
__device__ __forceinline__ void stv_stream(float* p, Vec16<float> v) { __stcg(reinterpret_cast<float4*>(p), make_float4(v.v[0], v.v[1], v.v[2], v.v[3])); }
__device__ __forceinline__ void stv_stream(double* p, Vec16<double> v) { __stcg(reinterpret_cast<double2*>(p), make_double2(v.v[0], v.v[1])); }
#endif
__device__ __forceinline__ Lane<float> lane_from_global(const cplx<float>* p) {
#ifdef HIPGP_EMU
    return *reinterpret_cast<const Lane<float>*>(p);
#else
    const float4 t = __ldcg(reinterpret_cast<const float4*>(p));
    Lane<float> r; r.re = make_float2(t.x, t.y); r.im = make_float2(t.z, t.w); return r;
#endif
}
__device__ __forceinline__ Lane<double> lane_from_global(const cplx<double>* p) {
#ifdef HIPGP_EMU
    return *reinterpret_cast<const Lane<double>*>(p);
#else
    const double2 t = __ldcg(reinterpret_cast<const double2*>(p));
    Lane<double> r; r.re = t.x; r.im = t.y; return r;
#endif
}
__device__ __forceinline__ void lane_to_global(cplx<float>* p, Lane<float> v) {
#ifdef HIPGP_EMU
    *reinterpret_cast<Lane<float>*>(p) = v;
#else
    __stcg(reinterpret_cast<float4*>(p), make_float4(v.re.x, v.re.y, v.im.x, v.im.y));
#endif
}
__device__ __forceinline__ void lane_to_global(cplx<double>* p, Lane<double> v) {
#ifdef HIPGP_EMU
    *reinterpret_cast<Lane<double>*>(p) = v;
#else
    __stcg(reinterpret_cast<double2*>(p), make_double2(v.re, v.im));
#endif
}

// one pairing of two scalars into a packed register pair (opaque to the optimiser, so it happens once per value)
__device__ __forceinline__ float2 pack2(float a, float b) {
#ifdef HIPGP_EMU
    return make_float2(a, b);
#else
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    float2 o;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
    return o;
#endif
}
// a lane from the complex values of its LPT lines
__device__ __forceinline__ Lane<float> lane_make(cplx<float> v0, cplx<float> v1) { Lane<float> r; r.re = pack2(v0.x, v1.x); r.im = pack2(v0.y, v1.y); return r; }
__device__ __forceinline__ Lane<double> lane_make(cplx<double> v0, cplx<double>) { Lane<double> r; r.re = v0.x; r.im = v0.y; return r; }

// row accessors of W for the row kernels (q = bin index along the row)
template <class T> struct WRow;
template <> struct WRow<float> {
    static __device__ __forceinline__ void store1(cplx<float>* row, int q, cplx<float> v) {
        float* f = reinterpret_cast<float*>(row) + 4 * (q >> 1) + (q & 1); st_stream(f, v.x); st_stream(f + 2, v.y);
    }
    static __device__ __forceinline__ cplx<float> load1(const cplx<float>* row, int q) {
        const float* f = reinterpret_cast<const float*>(row) + 4 * (q >> 1) + (q & 1); return mk<float>(ld_stream(f), ld_stream(f + 2));
    }
    // bins (q, q + 1), q even
    static __device__ __forceinline__ void store2(cplx<float>* row, int q, cplx<float> v0, cplx<float> v1) {
        Lane<float> t; t.re = make_float2(v0.x, v1.x); t.im = make_float2(v0.y, v1.y);
        lane_to_global(row + q, t);
    }
    static __device__ __forceinline__ void load2(const cplx<float>* row, int q, cplx<float>& v0, cplx<float>& v1) {
        const Lane<float> t = lane_from_global(row + q);
        v0 = mk<float>(t.re.x, t.im.x); v1 = mk<float>(t.re.y, t.im.y);
    }
};
template <> struct WRow<double> {
    static __device__ __forceinline__ void store1(cplx<double>* row, int q, cplx<double> v) { st_stream(row + q, v); }
    static __device__ __forceinline__ cplx<double> load1(const cplx<double>* row, int q) { return ld_stream(row + q); }
    static __device__ __forceinline__ void store2(cplx<double>* row, int q, cplx<double> v0, cplx<double> v1) { st_stream(row + q, v0); st_stream(row + q + 1, v1); }
    static __device__ __forceinline__ void load2(const cplx<double>* row, int q, cplx<double>& v0, cplx<double>& v1) { v0 = ld_stream(row + q); v1 = ld_stream(row + q + 1); }
};

// ---- asynchronous global -> shared copies (LDGSTS): no register staging, all requests in flight at once ----
template <int BYTES> __device__ __forceinline__ void cp_async(void* smem_dst, const void* gmem_src) {
#ifdef HIPGP_EMU
    std::memcpy(smem_dst, gmem_src, BYTES);
#else
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src));
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem_src));
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef HIPGP_EMU
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
__device__ __forceinline__ void cp_async_wait_all() {
#ifndef HIPGP_EMU
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

// ---- bulk asynchronous copies (TMA, 1-D): one instruction moves a whole contiguous block of rows between global and
// shared memory; completion of a load is signalled on an mbarrier (transaction bytes), of a store by its bulk group ----
#ifndef HIPGP_EMU
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "HIPGP_MBAR_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra HIPGP_MBAR_DONE;\n\t"
        "bra HIPGP_MBAR_WAIT;\n\t"
        "HIPGP_MBAR_DONE:\n\t}" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"((unsigned)__cvta_generic_to_shared(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// the same in two halves: the issuing thread may do other work (and other threads may READ the source) before it waits
__device__ __forceinline__ void bulk_s2g_issue(void* gmem_dst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"((unsigned)__cvta_generic_to_shared(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_s2g_wait() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
#endif

// ---------------------------------------------------------------------------------------------------------
// butterflies on lanes:  v[q] = sum_r v[r] w_R^{qr},  w_R = exp(-+ 2 pi i / R)
template <bool INV, class E> __device__ __forceinline__ void lb2(E* v) { const E a = v[0], b = v[1]; v[0] = a + b; v[1] = a - b; }
template <bool INV, class E> __device__ __forceinline__ void lb4(E* v) {
    const E a = v[0] + v[2], b = v[0] - v[2], c = v[1] + v[3], d = lmi<INV>(v[1] - v[3]);
    v[0] = a + c; v[1] = b + d; v[2] = a - c; v[3] = b - d;
}
template <bool INV, class T, class E> __device__ __forceinline__ void lb8(E* v) {
    const T h = (T)0.70710678118654752440;
    E e[4] = {v[0] + v[4], v[1] + v[5], v[2] + v[6], v[3] + v[7]};
    E o[4] = {v[0] - v[4], v[1] - v[5], v[2] - v[6], v[3] - v[7]};
    o[1] = lmul(o[1], mk<T>(h, INV ? h : -h));
    o[2] = lmi<INV>(o[2]);
    o[3] = lmul(o[3], mk<T>(-h, INV ? h : -h));
    lb4<INV>(e); lb4<INV>(o);
    v[0] = e[0]; v[2] = e[1]; v[4] = e[2]; v[6] = e[3];
    v[1] = o[0]; v[3] = o[1]; v[5] = o[2]; v[7] = o[3];
}
template <bool INV, class T, class E> __device__ __forceinline__ void lb16_tail(E (&t)[4][4], E* v) {
    const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173, h = (T)0.70710678118654752440;
    const cplx<T> w1 = mk<T>(c1, INV ? s1 : -s1), w2 = mk<T>(h, INV ? h : -h), w3 = mk<T>(s1, INV ? c1 : -c1);
    const cplx<T> w6 = mk<T>(-h, INV ? h : -h), w9 = mk<T>(-c1, INV ? -s1 : s1);
    t[1][1] = lmul(t[1][1], w1); t[1][2] = lmul(t[1][2], w2); t[1][3] = lmul(t[1][3], w3);
    t[2][1] = lmul(t[2][1], w2); t[2][2] = lmi<INV>(t[2][2]); t[2][3] = lmul(t[2][3], w6);
    t[3][1] = lmul(t[3][1], w3); t[3][2] = lmul(t[3][2], w6); t[3][3] = lmul(t[3][3], w9);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        E u[4] = {t[0][c], t[1][c], t[2][c], t[3][c]};
        lb4<INV>(u);
#pragma unroll
        for (int d = 0; d < 4; ++d) v[c + 4 * d] = u[d];
    }
}
template <bool INV, class T, class E> __device__ __forceinline__ void lb16(E* v) {
    E t[4][4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        E u[4] = {v[b], v[b + 4], v[b + 8], v[b + 12]};
        lb4<INV>(u);
#pragma unroll
        for (int c = 0; c < 4; ++c) t[b][c] = u[c];
    }
    lb16_tail<INV, T>(t, v);
}
template <bool INV, class T, class E> __device__ __forceinline__ void lb3(E* v) {
    const T s = (T)0.86602540378443864676;
    const E t = v[1] + v[2];
    const E m = lfma(t, (T)-0.5, v[0]);
    const E j = lmi<INV>(lscale(v[1] - v[2], s));
    v[0] = v[0] + t; v[1] = m + j; v[2] = m - j;
}
template <bool INV, class T, class E> __device__ __forceinline__ void lb5(E* v) {
    const T c1 = (T)0.30901699437494742410, c2 = (T)-0.80901699437494742410;
    const T s1 = (T)0.95105651629515357212, s2 = (T)0.58778525229247312917;
    const E a1 = v[1] + v[4], a2 = v[2] + v[3], b1 = v[1] - v[4], b2 = v[2] - v[3];
    const E m1 = lfma(a2, c2, lfma(a1, c1, v[0]));
    const E m2 = lfma(a2, c1, lfma(a1, c2, v[0]));
    const E j1 = lmi<INV>(lfma(b2, s2, lscale(b1, s1)));
    const E j2 = lmi<INV>(lfma(b2, -s1, lscale(b1, s2)));
    v[0] = v[0] + a1 + a2;
    v[1] = m1 + j1; v[4] = m1 - j1; v[2] = m2 + j2; v[3] = m2 - j2;
}
template <int R, bool INV, class T, class E> __device__ __forceinline__ void lbfly(E* v) {
    if (R == 2) lb2<INV>(v);
    else if (R == 3) lb3<INV, T>(v);
    else if (R == 4) lb4<INV>(v);
    else if (R == 5) lb5<INV, T>(v);
    else if (R == 8) lb8<INV, T>(v);
    else lb16<INV, T>(v);
}
// forward butterflies whose inputs v[R/2 .. R) are known to be zero (pruned zero padding); R in {2,4,8,16}
template <int R, class T, class E> __device__ __forceinline__ void lbfly_zero_hi(E* v) {
    if (R == 2) { v[1] = v[0]; }
    else if (R == 4) {
        const E a = v[0], c = v[1], d = lmi<false>(v[1]);
        v[0] = a + c; v[1] = a + d; v[2] = a - c; v[3] = a - d;
    } else if (R == 8) {
        const T h = (T)0.70710678118654752440;
        E e[4] = {v[0], v[1], v[2], v[3]};
        E o[4] = {v[0], lmul(v[1], mk<T>(h, -h)), lmi<false>(v[2]), lmul(v[3], mk<T>(-h, -h))};
        lb4<false>(e); lb4<false>(o);
        v[0] = e[0]; v[2] = e[1]; v[4] = e[2]; v[6] = e[3];
        v[1] = o[0]; v[3] = o[1]; v[5] = o[2]; v[7] = o[3];
    } else {
        E t[4][4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const E a = v[b], c = v[b + 4], d = lmi<false>(c);
            t[b][0] = a + c; t[b][1] = a + d; t[b][2] = a - c; t[b][3] = a - d;
        }
        lb16_tail<false, T>(t, v);
    }
}

// ---------------------------------------------------------------------------------------------------------
// compile-time radix lists
template <int... Rs> struct RL {};
template <class L> struct RLInfo;
template <> struct RLInfo<RL<>> { static constexpr int N = 1; static constexpr int count = 0; };
template <int R0, int... Rs> struct RLInfo<RL<R0, Rs...>> {
    static constexpr int N = R0 * RLInfo<RL<Rs...>>::N;
    static constexpr int count = 1 + RLInfo<RL<Rs...>>::count;
};
template <class L> struct RLLast;
template <int R0> struct RLLast<RL<R0>> { static constexpr int value = R0; };
template <int R0, int R1, int... Rs> struct RLLast<RL<R0, R1, Rs...>> { static constexpr int value = RLLast<RL<R1, Rs...>>::value; };
template <class L> struct RLFirst;
template <int R0, int... Rs> struct RLFirst<RL<R0, Rs...>> { static constexpr int value = R0; };
template <class L> struct RLMax;
template <int R0> struct RLMax<RL<R0>> { static constexpr int value = R0; };
template <int R0, int R1, int... Rs> struct RLMax<RL<R0, R1, Rs...>> {
    static constexpr int rest = RLMax<RL<R1, Rs...>>::value;
    static constexpr int value = R0 > rest ? R0 : rest;
};
__host__ __device__ constexpr bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }
__host__ __device__ constexpr int ilog2(int x) { return x <= 1 ? 0 : 1 + ilog2(x >> 1); }

// Tile geometry shared by the row and column kernels of one (length, lane count).
template <class T, int NL, int... Rs>
struct TileGeo {
    using List = RL<Rs...>;
    static constexpr int Ln = RLInfo<List>::N;
    static constexpr int NST = RLInfo<List>::count;
    static constexpr int RLAST = RLLast<List>::value;
    static constexpr int LOGRL = ilog2(RLAST);
    static constexpr bool PAD = (NL < 8) && (NST > 1) && is_pow2(RLAST);
    __host__ __device__ static constexpr int slot(int p) { return PAD ? p + (p >> LOGRL) : p; }
    static constexpr int SLOTS = PAD ? Ln + Ln / RLAST : Ln;
    __host__ __device__ static constexpr size_t smem_bytes(int extra_slots = 0) { return sizeof(Lane<T>) * (size_t)(SLOTS + extra_slots) * NL; }
    // lane-slots between the legs of a butterfly of stride S (S a multiple of RLAST, or S == 1 for the last stage)
    __host__ __device__ static constexpr int leg(int S) { return (S == 1 ? 1 : (PAD ? S + (S >> LOGRL) : S)) * NL; }
};

// per-stage twiddles w^{j r}, r = 1..R-1, from the table of the stage (fft_engine.cuh: paired in fp32, [r-1][j] in fp64)
template <int R, int S, class T>
__device__ __forceinline__ void lane_twiddles(cplx<T>* w, const cplx<T>* __restrict__ tab, int j) {
    if (TwLayout<T>::paired) {
#pragma unroll
        for (int p = 0; p < TwPairs<R>::n; ++p) {
            const TwPair<T> e = tw_pair_global<R, S, T>(tab, p, j);
            w[2 * p + 1] = e.a;
            if (2 * p + 2 < R) w[2 * p + 2] = e.b;
        }
    } else {
#pragma unroll
        for (int r = 1; r < R; ++r) w[r] = ldg_c(tab + (r - 1) * S + j);
    }
}

// one in-place shared-memory stage of sub-transform length Nt, radix R, over the whole tile
template <class G, class T, int NL, int Nt, int R, bool INV>
__device__ __forceinline__ void lane_stage_item(Lane<T>* s, const cplx<T>* __restrict__ tab, int it) {
    constexpr int S = Nt / R, LEG = G::leg(S);
    const int lane = it % NL, bf = it / NL;
    const int blk = bf / S, j = bf - blk * S;
    Lane<T>* base = s + (G::slot(blk * Nt + j) * NL + lane);
    cplx<T> w[R];
    if (S > 1) lane_twiddles<R, S>(w, tab, j);
    Lane<T> v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = base[r * LEG];
    if (INV) {
        if (S > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = lmulc(v[r], w[r]);
        }
        lbfly<R, true, T>(v);
    } else {
        lbfly<R, false, T>(v);
        if (S > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = lmul(v[r], w[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) base[r * LEG] = v[r];
}
template <class G, class T, int NL, int NT, int Nt, int R, bool INV>
__device__ __forceinline__ void lane_stage(Lane<T>* s, const cplx<T>* __restrict__ tab, int tid) {
    constexpr int ITEMS = (G::Ln / R) * NL, NIT = (ITEMS + NT - 1) / NT;
    if constexpr (ITEMS % NT == 0 && NIT * R <= 16) {
        // few small butterflies per thread: unrolled, so that the table and shared-memory loads of all of them overlap
#pragma unroll
        for (int k = 0; k < NIT; ++k) lane_stage_item<G, T, NL, Nt, R, INV>(s, tab, tid + k * NT);
    } else {
#pragma unroll 1
        for (int it = tid; it < ITEMS; it += NT) lane_stage_item<G, T, NL, Nt, R, INV>(s, tab, it);
    }
}

// middle stages (all but the first and the last of the list), each followed by a barrier
template <class G, class T, int NL, int NT, int Nt, int STG, int... Rs> struct LaneMidFwd;
template <class G, class T, int NL, int NT, int Nt, int STG, int R0> struct LaneMidFwd<G, T, NL, NT, Nt, STG, R0> {
    static __device__ __forceinline__ void run(Lane<T>*, const LineFft<T>&, int) {}
};
template <class G, class T, int NL, int NT, int Nt, int STG, int R0, int R1, int... Rs> struct LaneMidFwd<G, T, NL, NT, Nt, STG, R0, R1, Rs...> {
    static __device__ __forceinline__ void run(Lane<T>* s, const LineFft<T>& f, int tid) {
        lane_stage<G, T, NL, NT, Nt, R0, false>(s, f.twst + f.twoff[STG], tid);
        __syncthreads();
        LaneMidFwd<G, T, NL, NT, Nt / R0, STG + 1, R1, Rs...>::run(s, f, tid);
    }
};
template <class G, class T, int NL, int NT, int Nt, int STG, int... Rs> struct LaneMidInv;
template <class G, class T, int NL, int NT, int Nt, int STG, int R0> struct LaneMidInv<G, T, NL, NT, Nt, STG, R0> {
    static __device__ __forceinline__ void run(Lane<T>*, const LineFft<T>&, int) {}
};
template <class G, class T, int NL, int NT, int Nt, int STG, int R0, int R1, int... Rs> struct LaneMidInv<G, T, NL, NT, Nt, STG, R0, R1, Rs...> {
    static __device__ __forceinline__ void run(Lane<T>* s, const LineFft<T>& f, int tid) {
        LaneMidInv<G, T, NL, NT, Nt / R0, STG + 1, R1, Rs...>::run(s, f, tid);
        lane_stage<G, T, NL, NT, Nt, R0, true>(s, f.twst + f.twoff[STG], tid);
        __syncthreads();
    }
};

}  // namespace hipgp
