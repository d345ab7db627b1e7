// fft_engine.cuh -- shared-memory FFT building blocks.
//
// Design (see DESIGN.md "FFT engine"):
//  * every transform in the matvec pipeline is a batch of LINES held in shared memory as
//    s[pos * pos_stride + line] (lines along the fast index => consecutive threads touch consecutive
//    words for every stage, whatever the butterfly span);
//  * forward transforms are in-place decimation-in-frequency (natural order in, digit-reversed order
//    out); inverse transforms are the exact mirror, decimation-in-time (digit-reversed in, natural out).
//    Frequency-domain data therefore stays in digit-reversed order everywhere between the two, and the
//    spectrum is produced by the same forward kernels at set-up, so no permutation pass ever runs;
//  * radix-R butterflies (R in {2,3,4,5,8}) live in registers; stages exchange through shared memory;
//  * twiddles come from a per-length table computed in fp64 at plan time (TW[i] = exp(-2 pi i * i / Ln)).
#pragma once
#include "cuda_emu.h"

namespace hipgp {

template <class T> struct __align__(2 * sizeof(T)) cplx { T x, y; };

template <class T> __host__ __device__ __forceinline__ cplx<T> mk(T a, T b) { cplx<T> r; r.x = a; r.y = b; return r; }
template <class T> __host__ __device__ __forceinline__ cplx<T> operator+(cplx<T> a, cplx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <class T> __host__ __device__ __forceinline__ cplx<T> operator-(cplx<T> a, cplx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
template <class T> __host__ __device__ __forceinline__ cplx<T> operator*(cplx<T> a, cplx<T> b) { return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <class T> __host__ __device__ __forceinline__ cplx<T> operator*(cplx<T> a, T s) { return mk<T>(a.x * s, a.y * s); }
template <class T> __host__ __device__ __forceinline__ cplx<T> conj(cplx<T> a) { return mk<T>(a.x, -a.y); }
// a * conj(b)
template <class T> __host__ __device__ __forceinline__ cplx<T> mulc(cplx<T> a, cplx<T> b) { return mk<T>(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
// multiply by -i (forward) or +i (inverse)
template <bool INV, class T> __host__ __device__ __forceinline__ cplx<T> mul_mi(cplx<T> a) { return INV ? mk<T>(-a.y, a.x) : mk<T>(a.y, -a.x); }

#ifndef HIPGP_EMU
// Blackwell packed fp32 pairs (FADD2 / FMUL2 / FFMA2, sm_100+): one instruction per complex add, two per complex
// multiply.  (re, im) of a cplx<float> is exactly one f32x2 operand.
__device__ __forceinline__ float2 f2(cplx<float> a) { return make_float2(a.x, a.y); }
__device__ __forceinline__ cplx<float> c2(float2 a) { return mk<float>(a.x, a.y); }
__device__ __forceinline__ cplx<float> operator+(cplx<float> a, cplx<float> b) { return c2(__fadd2_rn(f2(a), f2(b))); }
__device__ __forceinline__ cplx<float> operator-(cplx<float> a, cplx<float> b) { return c2(__fadd2_rn(f2(a), make_float2(-b.x, -b.y))); }
__device__ __forceinline__ cplx<float> operator*(cplx<float> a, cplx<float> b) {
    return c2(__ffma2_rn(make_float2(a.x, a.x), f2(b), __fmul2_rn(make_float2(a.y, a.y), make_float2(-b.y, b.x))));
}
__device__ __forceinline__ cplx<float> operator*(cplx<float> a, float s) { return c2(__fmul2_rn(f2(a), make_float2(s, s))); }
__device__ __forceinline__ cplx<float> mulc(cplx<float> a, cplx<float> b) {
    return c2(__ffma2_rn(make_float2(a.x, a.x), make_float2(b.x, -b.y), __fmul2_rn(make_float2(a.y, a.y), make_float2(b.y, b.x))));
}
#endif

// read-only (non-coherent) load of a complex table entry
#ifdef HIPGP_EMU
template <class T> __device__ __forceinline__ cplx<T> ldg_c(const cplx<T>* p) { return *p; }
#else
__device__ __forceinline__ cplx<float> ldg_c(const cplx<float>* p) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p)); return mk<float>(v.x, v.y);
}
__device__ __forceinline__ cplx<double> ldg_c(const cplx<double>* p) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(p)); return mk<double>(v.x, v.y);
}
#endif

// fp32 stage twiddle tables hold PAIRS: entry [p][j] = (w^{j (2p+1)}, w^{j (2p+2)}), p < R / 2 -- one 16-byte load brings two
// twiddles; an unused second half (R even) is 1.  fp64 tables stay [r-1][j]: a pair would be 32 bytes, i.e. two loads that each use
// half of every sector (measured: +4 % on the fp64 matvec).
template <class T> struct alignas(16) TwPair { cplx<T> a, b; };
template <int R> struct TwPairs { static constexpr int n = R / 2; };            // = ceil((R - 1) / 2)
template <class T> struct TwLayout { static constexpr bool paired = sizeof(T) == 4; };
#ifdef HIPGP_EMU
template <class T> __device__ __forceinline__ TwPair<T> ldg_pair(const TwPair<T>* p) { return *p; }
#else
__device__ __forceinline__ TwPair<float> ldg_pair(const TwPair<float>* p) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    TwPair<float> e; e.a = mk<float>(v.x, v.y); e.b = mk<float>(v.z, v.w); return e;
}
__device__ __forceinline__ TwPair<double> ldg_pair(const TwPair<double>* p) {
    const double2 u = __ldg(reinterpret_cast<const double2*>(p)), v = __ldg(reinterpret_cast<const double2*>(p) + 1);
    TwPair<double> e; e.a = mk<double>(u.x, u.y); e.b = mk<double>(v.x, v.y); return e;
}
#endif
// pair p of butterfly position j from the GLOBAL table of a stage, whatever its layout
template <int R, int S, class T>
__device__ __forceinline__ TwPair<T> tw_pair_global(const cplx<T>* tab, int p, int j) {
    if (TwLayout<T>::paired) return ldg_pair(reinterpret_cast<const TwPair<T>*>(tab) + p * S + j);
    TwPair<T> e;
    e.a = ldg_c(tab + (2 * p) * S + j);
    e.b = 2 * p + 2 < R ? ldg_c(tab + (2 * p + 1) * S + j) : mk<T>(1, 0);
    return e;
}

constexpr int kMaxStages = 24;

// Device-side description of one line-FFT length.
template <class T>
struct LineFft {
    int Ln;                 // complex length
    int nst;                // number of stages
    int radix[kMaxStages];  // DIF order (stage 0 first)
    const cplx<T>* tw;      // Ln entries, exp(-2 pi i k / Ln)
    const int* rev;         // Ln entries: rev[p] = frequency index k stored at position p after DIF
    const int* pos;         // Ln entries: pos[k] = position p (inverse permutation)
    const cplx<T>* twst;    // per-stage twiddles, stage s at twst + twoff[s]: fp32 TwPair [p * S + j] = (w_Nt^{j (2p+1)}, w_Nt^{j (2p+2)}); fp64 [(r-1) * S + j]
    int twoff[kMaxStages];
};

// ---- radix butterflies: v[q] = sum_r v[r] w_R^{qr}  (w_R = exp(-+2 pi i / R)) ----------------------
template <bool INV, class T> __device__ __forceinline__ void bfly2(cplx<T>* v) {
    cplx<T> a = v[0], b = v[1];
    v[0] = a + b; v[1] = a - b;
}
template <bool INV, class T> __device__ __forceinline__ void bfly4(cplx<T>* v) {
    cplx<T> a = v[0] + v[2], b = v[0] - v[2], c = v[1] + v[3], d = mul_mi<INV>(v[1] - v[3]);
    v[0] = a + c; v[1] = b + d; v[2] = a - c; v[3] = b - d;
}
template <bool INV, class T> __device__ __forceinline__ void bfly8(cplx<T>* v) {
    const T h = (T)0.70710678118654752440;
    // radix-2 split: even/odd halves
    cplx<T> e[4] = {v[0] + v[4], v[1] + v[5], v[2] + v[6], v[3] + v[7]};
    cplx<T> o[4] = {v[0] - v[4], v[1] - v[5], v[2] - v[6], v[3] - v[7]};
    // twiddle odd part by w_8^r
    cplx<T> t1 = o[1], t3 = o[3];
    o[1] = INV ? mk<T>((t1.x - t1.y) * h, (t1.x + t1.y) * h) : mk<T>((t1.x + t1.y) * h, (t1.y - t1.x) * h);
    o[2] = mul_mi<INV>(o[2]);
    o[3] = INV ? mk<T>((-t3.x - t3.y) * h, (t3.x - t3.y) * h) : mk<T>((t3.y - t3.x) * h, (-t3.x - t3.y) * h);
    bfly4<INV>(e); bfly4<INV>(o);
    v[0] = e[0]; v[2] = e[1]; v[4] = e[2]; v[6] = e[3];
    v[1] = o[0]; v[3] = o[1]; v[5] = o[2]; v[7] = o[3];
}
template <bool INV, class T> __device__ __forceinline__ void bfly3(cplx<T>* v) {
    const T s = (T)0.86602540378443864676;
    cplx<T> t = v[1] + v[2];
    cplx<T> m = mk<T>(v[0].x - (T)0.5 * t.x, v[0].y - (T)0.5 * t.y);
    cplx<T> d = v[1] - v[2];
    cplx<T> j = INV ? mk<T>(-s * d.y, s * d.x) : mk<T>(s * d.y, -s * d.x);   // -+ i s d
    v[0] = v[0] + t; v[1] = m + j; v[2] = m - j;
}
template <bool INV, class T> __device__ __forceinline__ void bfly5(cplx<T>* v) {
    const T c1 = (T)0.30901699437494742410, c2 = (T)-0.80901699437494742410;
    const T s1 = (T)0.95105651629515357212, s2 = (T)0.58778525229247312917;
    cplx<T> a1 = v[1] + v[4], a2 = v[2] + v[3], b1 = v[1] - v[4], b2 = v[2] - v[3];
    cplx<T> m1 = mk<T>(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
    cplx<T> m2 = mk<T>(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
    cplx<T> n1 = mk<T>(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
    cplx<T> n2 = mk<T>(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
    cplx<T> j1 = INV ? mk<T>(-n1.y, n1.x) : mk<T>(n1.y, -n1.x);   // -+ i n1
    cplx<T> j2 = INV ? mk<T>(-n2.y, n2.x) : mk<T>(n2.y, -n2.x);
    v[0] = v[0] + a1 + a2;
    v[1] = m1 + j1; v[4] = m1 - j1; v[2] = m2 + j2; v[3] = m2 - j2;
}
// radix-16 as 4 x 4:  r = b + 4a, q = c + 4d:  X[c+4d] = sum_b w4^{db} ( w16^{cb} sum_a v[b+4a] w4^{ca} )
template <bool INV, class T> __device__ __forceinline__ void bfly16(cplx<T>* v) {
    const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173, h = (T)0.70710678118654752440;
    cplx<T> t[4][4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        cplx<T> u[4] = {v[b], v[b + 4], v[b + 8], v[b + 12]};
        bfly4<INV>(u);
#pragma unroll
        for (int c = 0; c < 4; ++c) t[b][c] = u[c];
    }
    const cplx<T> w1 = mk<T>(c1, INV ? s1 : -s1), w2 = mk<T>(h, INV ? h : -h), w3 = mk<T>(s1, INV ? c1 : -c1);
    const cplx<T> w6 = mk<T>(-h, INV ? h : -h), w9 = mk<T>(-c1, INV ? -s1 : s1);
    t[1][1] = t[1][1] * w1; t[1][2] = t[1][2] * w2; t[1][3] = t[1][3] * w3;
    t[2][1] = t[2][1] * w2; t[2][2] = mul_mi<INV>(t[2][2]); t[2][3] = t[2][3] * w6;
    t[3][1] = t[3][1] * w3; t[3][2] = t[3][2] * w6; t[3][3] = t[3][3] * w9;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        cplx<T> u[4] = {t[0][c], t[1][c], t[2][c], t[3][c]};
        bfly4<INV>(u);
#pragma unroll
        for (int d = 0; d < 4; ++d) v[c + 4 * d] = u[d];
    }
}
template <int R, bool INV, class T> __device__ __forceinline__ void bfly(cplx<T>* v) {
    if (R == 2) bfly2<INV>(v);
    else if (R == 3) bfly3<INV>(v);
    else if (R == 4) bfly4<INV>(v);
    else if (R == 5) bfly5<INV>(v);
    else if (R == 8) bfly8<INV>(v);
    else bfly16<INV>(v);
}

// One stage over `nlines` lines.  Nt = current sub-transform length, S = Nt / R, twmul = Ln / Nt.
// Work item w -> (line = w % nlines, butterfly = w / nlines) so that a warp walks the fast index.
template <int R, bool INV, class T>
__device__ __forceinline__ void fft_stage(cplx<T>* s, int pos_stride, int nlines, int Ln, int Nt,
                                          const cplx<T>* __restrict__ tw, int tid, int nthreads) {
    const int S = Nt / R;
    const int twmul = Ln / Nt;
    const int items = (Ln / R) * nlines;
    for (int w = tid; w < items; w += nthreads) {
        const int line = w % nlines;
        const int bf = w / nlines;
        const int blk = bf / S;
        const int j = bf - blk * S;
        cplx<T>* base = s + (size_t)(blk * Nt + j) * pos_stride + line;
        cplx<T> v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = base[(size_t)r * S * pos_stride];
        if (INV) {   // DIT: conj twiddle first, then butterfly
            if (j != 0) {
#pragma unroll
                for (int r = 1; r < R; ++r) v[r] = mulc(v[r], ldg_c(tw + j * r * twmul));
            }
            bfly<R, true>(v);
        } else {     // DIF: butterfly, then twiddle
            bfly<R, false>(v);
            if (j != 0) {
#pragma unroll
                for (int r = 1; r < R; ++r) v[r] = v[r] * ldg_c(tw + j * r * twmul);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) base[(size_t)r * S * pos_stride] = v[r];
    }
}

template <bool INV, class T>
__device__ __forceinline__ void fft_stage_dyn(int R, cplx<T>* s, int pos_stride, int nlines, int Ln, int Nt,
                                              const cplx<T>* __restrict__ tw, int tid, int nthreads) {
    switch (R) {
        case 16: fft_stage<16, INV>(s, pos_stride, nlines, Ln, Nt, tw, tid, nthreads); break;
        case 8: fft_stage<8, INV>(s, pos_stride, nlines, Ln, Nt, tw, tid, nthreads); break;
        case 4: fft_stage<4, INV>(s, pos_stride, nlines, Ln, Nt, tw, tid, nthreads); break;
        case 2: fft_stage<2, INV>(s, pos_stride, nlines, Ln, Nt, tw, tid, nthreads); break;
        case 3: fft_stage<3, INV>(s, pos_stride, nlines, Ln, Nt, tw, tid, nthreads); break;
        default: fft_stage<5, INV>(s, pos_stride, nlines, Ln, Nt, tw, tid, nthreads); break;
    }
}

// Forward DIF over all lines (natural in -> digit-reversed out).  Ends with a __syncthreads().
template <class T>
__device__ __forceinline__ void fft_forward(cplx<T>* s, int pos_stride, int nlines, const LineFft<T>& f,
                                            int tid, int nthreads) {
    int Nt = f.Ln;
    for (int st = 0; st < f.nst; ++st) {
        const int R = f.radix[st];
        fft_stage_dyn<false>(R, s, pos_stride, nlines, f.Ln, Nt, f.tw, tid, nthreads);
        Nt /= R;
        __syncthreads();
    }
}
// Inverse DIT (digit-reversed in -> natural out), unnormalised.  Ends with a __syncthreads().
template <class T>
__device__ __forceinline__ void fft_inverse(cplx<T>* s, int pos_stride, int nlines, const LineFft<T>& f,
                                            int tid, int nthreads) {
    int Nt = 1;
    for (int st = f.nst - 1; st >= 0; --st) {
        const int R = f.radix[st];
        Nt *= R;
        fft_stage_dyn<true>(R, s, pos_stride, nlines, f.Ln, Nt, f.tw, tid, nthreads);
        __syncthreads();
    }
}

}  // namespace hipgp
