// fast_kernels.cuh -- compile-time specialised versions of the three matvec passes on the lane engine (lane_fft.cuh).
//
// Same math and the same digit-reversed layouts as conv_kernels.cuh (the generic, runtime-radix kernels stay as the
// fallback for lengths without an instantiation).  What is different:
//   * a thread owns (butterfly, lane): 16-byte lanes (two fp32 lines / one fp64 line), LDS.128 / STS.128 only, packed
//     f32x2 arithmetic across the two lines of a lane, all butterfly legs at [base + immediate];
//   * the first forward stage reads its operands straight from global memory (zero padding = skipped loads and a
//     pruned butterfly when the upper half of the inputs is padding; PCG vector updates fused into the row loads) and
//     the last inverse stage writes straight to global memory (crop = skipped stores, dot products fused);
//   * column pass: last forward stage, spectrum multiply and first inverse stage run back to back in registers;
//   * row passes: the r2c split / c2r merge works on (k, H-k) PAIRS (one thread produces both members from one pair of
//     shared-memory reads) straight to / from global memory;
//   * tiles are small (64-72 KB) so that two or three CTAs are resident per SM and one CTA's global traffic overlaps
//     another's butterflies.
#pragma once
#include "lane_fft.cuh"

namespace hipgp {

template <class T> __device__ __forceinline__ void ld2(const T* p, bool vec, T& a, T& b, bool ok0, bool ok1) {
    if (vec && ok1) { const cplx<T> t = *reinterpret_cast<const cplx<T>*>(p); a = t.x; b = t.y; }
    else { a = ok0 ? p[0] : (T)0; b = ok1 ? p[1] : (T)0; }
}
template <class T> __device__ __forceinline__ void st2(T* p, bool vec, T a, T b, bool ok0, bool ok1) {
    if (vec && ok1) { *reinterpret_cast<cplx<T>*>(p) = mk<T>(a, b); }
    else { if (ok0) p[0] = a; if (ok1) p[1] = b; }
}

// deterministic per-row reduction of per-thread partials in smem scratch: warp `row` sums scratch[row*NI .. +NI)
__device__ __forceinline__ void rows_reduce_partials(const double* scratch, int NI, int nl, long g0, double* partial, int tid, int nthreads) {
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
    for (int row = warp; row < nl; row += nwarps) {
        double a = 0.0;
        for (int i = lane; i < NI; i += 32) a += scratch[row * NI + i];
        a = warp_sum(a);
        if (lane == 0) partial[g0 + row] = a;
    }
}

// spectrum factor(s) of one lane: `idx` = index of the lane's first line in the spectrum array
__device__ __forceinline__ Lane<float> lane_spec(Lane<float> v, const void* spec, int kind, size_t idx) {
    if (kind == SPEC_REAL) {
        const cplx<float> sv = ldg_c(reinterpret_cast<const cplx<float>*>(reinterpret_cast<const float*>(spec) + idx));   // two reals
        return lmul_real2(v, sv.x, sv.y);
    }
    const cplx<float>* sp = reinterpret_cast<const cplx<float>*>(spec) + idx;
    const cplx<float> w0 = ldg_c(sp), w1 = ldg_c(sp + 1);
    return kind == SPEC_CPLX ? lmul_cplx2<false>(v, w0, w1) : lmul_cplx2<true>(v, w0, w1);
}
__device__ __forceinline__ Lane<double> lane_spec(Lane<double> v, const void* spec, int kind, size_t idx) {
    if (kind == SPEC_REAL) return lscale(v, __ldg(reinterpret_cast<const double*>(spec) + idx));
    const cplx<double> w = ldg_c(reinterpret_cast<const cplx<double>*>(spec) + idx);
    return kind == SPEC_CPLX ? lmul(v, w) : lmulc(v, w);
}

// =====================================================================================================
// Column pass.  A CTA owns NL lanes (= NL * LPT neighbouring lines) of one (outer, batch) slice.
// =====================================================================================================
template <class T, int NL, int NT, int MINB, int R0, int... Rs>
__global__ void __launch_bounds__(NT, MINB) cols_fast_kernel(ColsParams<T> P) {
    using G = TileGeo<T, NL, R0, Rs...>;
    constexpr int Ln = G::Ln, NST = G::NST, RLAST = G::RLAST, LPT = LaneInfo<T>::LPT, TBL = NL * LPT, S0 = Ln / R0;
    HIPGP_DYN_SMEM(smem_raw);
    Lane<T>* s = reinterpret_cast<Lane<T>*>(smem_raw);
    const int tid = threadIdx.x;
    if (P.done_flag && *P.done_flag) return;
    const long c0 = (long)blockIdx.x * TBL;
    const long nvalid = P.inner - c0;                       // lines of this tile that exist
    const cplx<T>* in = P.in + (size_t)blockIdx.y * P.in_ostride + (size_t)blockIdx.z * P.in_bstride + c0;
    cplx<T>* out = P.out + (size_t)blockIdx.y * P.out_ostride + (size_t)blockIdx.z * P.out_bstride + c0;
    const int mode = P.mode;
    const long pitch = P.pitch;
    const size_t spitch = P.spec_pitch ? (size_t)P.spec_pitch : (size_t)P.pitch;
    auto in_off = [&](int i) { return cols_rowoff<T>(i, pitch, P.in_split_len, P.in_split_stride); };
    auto out_off = [&](int i) { return cols_rowoff<T>(i, pitch, P.out_split_len, P.out_split_stride); };

    if constexpr (NST == 1) {
        // the whole line lives in one thread's registers
        for (int lane = tid; lane < NL; lane += NT) {
            if ((long)lane * LPT >= nvalid) continue;
            const int rows_in = mode == CM_INV ? Ln : P.n_in;
            const int rows_out = mode == CM_FWD ? Ln : P.n_out;
            Lane<T> v[R0];
#pragma unroll
            for (int r = 0; r < R0; ++r) v[r] = r < rows_in ? lane_from_global(in + in_off(r) + lane * LPT) : lzero<T>();
            if (mode != CM_INV) lbfly<R0, false, T>(v);
            if (mode == CM_FUSED) {
#pragma unroll
                for (int r = 0; r < R0; ++r) v[r] = lane_spec(v[r], P.spec, P.spec_kind, (size_t)r * spitch + c0 + lane * LPT);
            }
            if (mode != CM_FWD) lbfly<R0, true, T>(v);
#pragma unroll
            for (int r = 0; r < R0; ++r) if (r < rows_out) lane_to_global(out + out_off(r) + lane * LPT, v[r]);
        }
        return;
    } else {
        constexpr int LEG0 = G::leg(S0);
        const cplx<T>* tw0 = P.f.twst + P.f.twoff[0];

        // ---- first forward stage, operands straight from global memory (zero padding = skipped loads) ----
        if (mode != CM_INV) {
            const int n_in = P.n_in;
            const bool zero_hi = is_pow2(R0) && n_in <= Ln / 2;
#pragma unroll 1
            for (int it = tid; it < S0 * NL; it += NT) {
                const int lane = it % NL, j = it / NL;
                const bool ok = (long)lane * LPT < nvalid;
                const cplx<T>* gp = in + lane * LPT;
                cplx<T> w[R0];
                lane_twiddles<R0, S0>(w, tw0, j);
                Lane<T> v[R0];
                if (zero_hi) {
#pragma unroll
                    for (int r = 0; r < R0 / 2; ++r) {
                        const int i = j + r * S0;
                        v[r] = (ok && i < n_in) ? lane_from_global(gp + in_off(i)) : lzero<T>();
                    }
                    lbfly_zero_hi<R0, T>(v);
                } else {
#pragma unroll
                    for (int r = 0; r < R0; ++r) {
                        const int i = j + r * S0;
                        v[r] = (ok && i < n_in) ? lane_from_global(gp + in_off(i)) : lzero<T>();
                    }
                    lbfly<R0, false, T>(v);
                }
#pragma unroll
                for (int r = 1; r < R0; ++r) v[r] = lmul(v[r], w[r]);
                Lane<T>* base = s + (G::slot(j) * NL + lane);
#pragma unroll
                for (int r = 0; r < R0; ++r) base[r * LEG0] = v[r];
            }
            __syncthreads();
            LaneMidFwd<G, T, NL, NT, Ln / R0, 1, Rs...>::run(s, P.f, tid);
        }

        // ---- last forward stage + spectrum + first inverse stage: RLAST neighbouring positions, in registers ----
        {
#pragma unroll 1
            for (int it = tid; it < (Ln / RLAST) * NL; it += NT) {
                const int lane = it % NL, bf = it / NL;
                const bool ok = (long)lane * LPT < nvalid;
                const int p0 = bf * RLAST;
                Lane<T>* base = s + (G::slot(p0) * NL + lane);
                Lane<T> v[RLAST];
                if (mode == CM_INV) {
#pragma unroll
                    for (int r = 0; r < RLAST; ++r) v[r] = ok ? lane_from_global(in + in_off(p0 + r) + lane * LPT) : lzero<T>();
                } else {
#pragma unroll
                    for (int r = 0; r < RLAST; ++r) v[r] = base[r * NL];
                    lbfly<RLAST, false, T>(v);
                }
                if (mode == CM_FUSED && ok) {
                    const size_t sidx = (size_t)p0 * spitch + c0 + lane * LPT;
#pragma unroll
                    for (int r = 0; r < RLAST; ++r) v[r] = lane_spec(v[r], P.spec, P.spec_kind, sidx + (size_t)r * spitch);
                }
                if (mode == CM_FWD) {
                    if (ok) {
#pragma unroll
                        for (int r = 0; r < RLAST; ++r) lane_to_global(out + out_off(p0 + r) + lane * LPT, v[r]);
                    }
                } else {
                    lbfly<RLAST, true, T>(v);
#pragma unroll
                    for (int r = 0; r < RLAST; ++r) base[r * NL] = v[r];
                }
            }
            if (mode == CM_FWD) return;
            __syncthreads();
        }

        // ---- inverse middle stages, then the last inverse stage straight to global memory (crop = skipped stores) ----
        LaneMidInv<G, T, NL, NT, Ln / R0, 1, Rs...>::run(s, P.f, tid);
        {
            const int n_out = P.n_out;
            const bool out_lo = is_pow2(R0) && n_out <= Ln / 2;
#pragma unroll 1
            for (int it = tid; it < S0 * NL; it += NT) {
                const int lane = it % NL, j = it / NL;
                const bool ok = (long)lane * LPT < nvalid;
                cplx<T> w[R0];
                lane_twiddles<R0, S0>(w, tw0, j);
                const Lane<T>* base = s + (G::slot(j) * NL + lane);
                Lane<T> v[R0];
#pragma unroll
                for (int r = 0; r < R0; ++r) v[r] = base[r * LEG0];
#pragma unroll
                for (int r = 1; r < R0; ++r) v[r] = lmulc(v[r], w[r]);
                cplx<T>* gp = out + lane * LPT;
                if (out_lo) {          // outputs R0/2 .. R0-1 are cropped: the compiler drops their arithmetic
                    lbfly<R0, true, T>(v);
                    if (ok) {
#pragma unroll
                        for (int r = 0; r < R0 / 2; ++r) { const int i = j + r * S0; if (i < n_out) lane_to_global(gp + out_off(i), v[r]); }
                    }
                } else {
                    lbfly<R0, true, T>(v);
                    if (ok) {
#pragma unroll
                        for (int r = 0; r < R0; ++r) { const int i = j + r * S0; if (i < n_out) lane_to_global(gp + out_off(i), v[r]); }
                    }
                }
            }
        }
    }
}

// =====================================================================================================
// Row passes.  H = product of the radix list; a CTA owns NL lanes = NL * LPT rows.
// =====================================================================================================
template <class T, int NL, int NT, int MINB, int R0, int... Rs>
__global__ void __launch_bounds__(NT, MINB) rows_fwd_fast_kernel(RowsParams<T> P) {
    using G = TileGeo<T, NL, R0, Rs...>;
    constexpr int H = G::Ln, NST = G::NST, RLAST = G::RLAST, LPT = LaneInfo<T>::LPT, NROW = NL * LPT, S0 = H / R0;
    constexpr int LEG0 = G::leg(NST > 1 ? S0 : 1);
    HIPGP_DYN_SMEM(smem_raw);
    Lane<T>* s = reinterpret_cast<Lane<T>*>(smem_raw);
    double* scratch = reinterpret_cast<double*>(smem_raw + G::smem_bytes());
    const int tid = threadIdx.x;
    if (P.mode != RF_PLAIN && P.st.flags[0]) return;
    const long g0 = (long)blockIdx.x * NROW;
    const long g1 = g0 + NROW < P.total_rows ? g0 + NROW : P.total_rows;
    const int nl = (int)(g1 - g0);
    const int n = P.n_real;
    const int mode = P.mode;
    const bool vec = ((n & 1) == 0) && P.vec_ok;             // rows start on even offsets => 2-element accesses are aligned
    const bool first_it = (mode == RF_PUPDATE) && (P.st.flags[2] != 0);
    const bool want_dot = (mode == RF_XRUPDATE || mode == RF_SELFDOT);
    // per-row scalars, computed once (the 64-bit divisions stay out of the element loops)
    __shared__ T s_coef[32];
    __shared__ long s_wbase[32];
    if (tid < NROW && tid < nl) {
        const long gr = g0 + tid;
        const long b = gr / P.nrows;
        s_wbase[tid] = (b * P.W_rows + (gr - b * P.nrows)) * P.W_pitch;
        T coef = 0;
        if (mode == RF_PUPDATE) coef = first_it ? (T)0 : (T)(P.st.zr[b] / P.st.zr_prev[b]);
        else if (mode == RF_XRUPDATE) coef = (T)(P.st.zr[b] / P.st.pAp[b]);
        s_coef[tid] = coef;
    }
    __syncthreads();

    // ---- first DIF stage fused with the load (and the PCG vector update) ----
    {
        const bool zero_hi = is_pow2(R0) && R0 > 1 && (n + 1) / 2 <= H / 2;
        const cplx<T>* tw0 = P.f.twst + P.f.twoff[0];
        // element loader: packed complex e = x[2e] + i x[2e+1] of row `row` (tile-local), with the fused vector update
        auto load_elem = [&](int row, int e, double& accd) -> cplx<T> {
            const int i = 2 * e;
            const bool ok0 = i < n, ok1 = i + 1 < n;
            T a = 0, b2 = 0;
            if (row < nl && ok0) {
                const size_t off = (size_t)(g0 + row) * n + i;
                const T coef = s_coef[row];
                if (mode == RF_PLAIN) {
                    ld2(P.in + off, vec, a, b2, ok0, ok1);
                } else if (mode == RF_PUPDATE) {
                    T z0, z1; ld2(P.in + off, vec, z0, z1, ok0, ok1);
                    if (first_it) { a = z0; b2 = z1; }
                    else { T p0, p1; ld2((const T*)P.v0 + off, vec, p0, p1, ok0, ok1); a = z0 + coef * p0; b2 = z1 + coef * p1; }
                    st2(P.v0 + off, vec, a, b2, ok0, ok1);
                } else if (mode == RF_SELFDOT) {
                    ld2(P.in + off, vec, a, b2, ok0, ok1);
                    accd += (double)(a * a) + (double)(b2 * b2);
                } else {
                    T p0, p1, x0, x1, r0, r1, q0, q1;
                    ld2(P.v2 + off, vec, p0, p1, ok0, ok1);
                    ld2((const T*)P.v1 + off, vec, x0, x1, ok0, ok1);
                    ld2((const T*)P.v0 + off, vec, r0, r1, ok0, ok1);
                    ld2(P.in + off, vec, q0, q1, ok0, ok1);
                    st2(P.v1 + off, vec, x0 + coef * p0, x1 + coef * p1, ok0, ok1);
                    a = r0 - coef * q0; b2 = ok1 ? r1 - coef * q1 : (T)0;
                    st2(P.v0 + off, vec, a, b2, ok0, ok1);
                    accd += (double)(a * a) + (double)(b2 * b2);
                }
            }
            return mk<T>(a, b2);
        };
#pragma unroll 1
        for (int it = tid; it < S0 * NL; it += NT) {
            const int lane = it % NL, j = it / NL;
            cplx<T> w[R0];
            if (NST > 1) lane_twiddles<R0, S0>(w, tw0, j);
            double accd[LPT];
#pragma unroll
            for (int l = 0; l < LPT; ++l) accd[l] = 0.0;
            Lane<T> v[R0];
            if (zero_hi) {
#pragma unroll
                for (int r = 0; r < R0 / 2; ++r) {
#pragma unroll
                    for (int l = 0; l < LPT; ++l) lane_set(v[r], l, load_elem(lane * LPT + l, j + r * S0, accd[l]));
                }
                lbfly_zero_hi<R0, T>(v);
            } else {
#pragma unroll
                for (int r = 0; r < R0; ++r) {
#pragma unroll
                    for (int l = 0; l < LPT; ++l) lane_set(v[r], l, load_elem(lane * LPT + l, j + r * S0, accd[l]));
                }
                lbfly<R0, false, T>(v);
            }
            if (want_dot) {
#pragma unroll
                for (int l = 0; l < LPT; ++l) scratch[(lane * LPT + l) * S0 + j] = accd[l];
            }
            if (NST > 1) {
#pragma unroll
                for (int r = 1; r < R0; ++r) v[r] = lmul(v[r], w[r]);
            }
            Lane<T>* base = s + (G::slot(j) * NL + lane);
#pragma unroll
            for (int r = 0; r < R0; ++r) base[r * LEG0] = v[r];
        }
    }
    __syncthreads();
    if (want_dot) {
        rows_reduce_partials(scratch, S0, nl, g0, P.st.partial, tid, NT);
        pcg_finalize(P.st, mode == RF_XRUPDATE ? DOT_RR : DOT_ZR, g0, g1, P.nrows, tid, NT);
    }

    if constexpr (NST > 1) {
        LaneMidFwd<G, T, NL, NT, H / R0, 1, Rs...>::run(s, P.f, tid);
        lane_stage<G, T, NL, NT, RLAST, RLAST, false>(s, P.f.twst, tid);
        __syncthreads();
    }

    // ---- r2c split on (k, H-k) pairs, straight to global memory ----
    {
#pragma unroll 1
        for (int it = tid; it < (H / 2 + 1) * NL; it += NT) {
            const int lane = it % NL, pi = it / NL;
            const int q = P.pairq[pi];
            Lane<T> Xa, Xb;
            int qb;
            bool two;
            if (q == 0) {
                const Lane<T> z = s[lane];
                Lane<T> t; t.re = z.im; t.im = z.im;                 // (im, im)
                Lane<T> u; u.re = z.re; u.im = z.re;                 // (re, re)
                Xa = lscale(u + t, (T)2); Xb = lscale(u - t, (T)2);  // re parts are the values; im parts are zeroed below
                Xa.im = lzero<T>().im; Xb.im = lzero<T>().im;
                qb = H; two = true;
            } else {
                const int q2 = P.part[q];
                const Lane<T> a = s[G::slot(q) * NL + lane], c = lconj(s[G::slot(q2) * NL + lane]);
                const Lane<T> E = a + c;
                const Lane<T> t = lmul(lmi<false>(a - c), P.twLp[q]);
                Xa = E + t; Xb = lconj(E - t);
                qb = q2; two = q2 != q;
            }
#pragma unroll
            for (int l = 0; l < LPT; ++l) {
                const int row = lane * LPT + l;
                if (row < nl) {
                    cplx<T>* dst = P.W + s_wbase[row];
                    dst[q] = lane_get(Xa, l);
                    if (two) dst[qb] = lane_get(Xb, l);
                }
            }
        }
    }
}

template <class T, int NL, int NT, int MINB, int R0, int... Rs>
__global__ void __launch_bounds__(NT, MINB) rows_inv_fast_kernel(RowsParams<T> P) {
    using G = TileGeo<T, NL, R0, Rs...>;
    constexpr int H = G::Ln, NST = G::NST, RLAST = G::RLAST, LPT = LaneInfo<T>::LPT, NROW = NL * LPT, S0 = H / R0;
    constexpr int LEG0 = G::leg(NST > 1 ? S0 : 1);
    HIPGP_DYN_SMEM(smem_raw);
    Lane<T>* s = reinterpret_cast<Lane<T>*>(smem_raw);
    double* scratch = reinterpret_cast<double*>(smem_raw + G::smem_bytes());
    const int tid = threadIdx.x;
    if (P.mode != RI_PLAIN && P.st.flags[0]) return;
    const long g0 = (long)blockIdx.x * NROW;
    const long g1 = g0 + NROW < P.total_rows ? g0 + NROW : P.total_rows;
    const int nl = (int)(g1 - g0);
    const int n = P.n_real;

    __shared__ long s_wbase[32];
    if (tid < NROW && tid < nl) {
        const long gr = g0 + tid;
        const long b = gr / P.nrows;
        s_wbase[tid] = (b * P.W_rows + (gr - b * P.nrows)) * P.W_pitch;
    }
    __syncthreads();

    // ---- c2r merge on (k, H-k) pairs, straight from global memory into shared memory ----
    {
        const int spec_kind = P.spec_kind;
#pragma unroll 1
        for (int it = tid; it < (H / 2 + 1) * NL; it += NT) {
            const int lane = it % NL, pi = it / NL;
            const int q = P.pairq[pi];
            const int q2 = q == 0 ? H : P.part[q];
            Lane<T> a = lzero<T>(), c = lzero<T>();
#pragma unroll
            for (int l = 0; l < LPT; ++l) {
                const int row = lane * LPT + l;
                if (row < nl) {
                    const cplx<T>* src = P.W + s_wbase[row];
                    cplx<T> ya = src[q], yc = src[q2];
                    if (spec_kind != SPEC_NONE) { ya = apply_spec(ya, P.spec, spec_kind, (size_t)q); yc = apply_spec(yc, P.spec, spec_kind, (size_t)q2); }
                    lane_set(a, l, ya); lane_set(c, l, yc);
                }
            }
            if (q == 0) {
                // Z[0] = (Y0 + YH) + i (Y0 - YH), real parts only
                Lane<T> z; z.re = (a + c).re; z.im = (a - c).re;
                s[lane] = z;
            } else {
                // bin q (frequency k):  E + i O with O = conj(w^k)(Y[k] - conj Y[k']);  bin q2 is conj(E - i O)
                c = lconj(c);
                const Lane<T> E = a + c;
                const Lane<T> iO = lmi<true>(lmulc(a - c, P.twLp[q]));
                s[G::slot(q) * NL + lane] = E + iO;
                if (q2 != q) s[G::slot(q2) * NL + lane] = lconj(E - iO);
            }
        }
    }
    __syncthreads();

    if constexpr (NST > 1) {
        lane_stage<G, T, NL, NT, RLAST, RLAST, true>(s, P.f.twst, tid);
        __syncthreads();
        LaneMidInv<G, T, NL, NT, H / R0, 1, Rs...>::run(s, P.f, tid);
    }

    // ---- last inverse stage fused with the store (crop) and the dot product ----
    const bool vec = ((n & 1) == 0) && P.vec_ok;
    const bool want_dot = P.mode == RI_DOT;
    {
        const bool out_lo = is_pow2(R0) && R0 > 1 && (n + 1) / 2 <= H / 2;
        const cplx<T>* tw0 = P.f.twst + P.f.twoff[0];
        auto store_elem = [&](int row, int e, cplx<T> val, double& accd) {
            const int i = 2 * e;
            const bool ok0 = i < n, ok1 = i + 1 < n;
            if (row < nl && ok0) {
                const size_t off = (size_t)(g0 + row) * n + i;
                st2(P.out + off, vec, val.x, val.y, ok0, ok1);
                if (want_dot) {
                    T o0, o1; ld2((const T*)P.v0 + off, vec, o0, o1, ok0, ok1);
                    accd += (double)(val.x * o0) + (ok1 ? (double)(val.y * o1) : 0.0);
                }
            }
        };
#pragma unroll 1
        for (int it = tid; it < S0 * NL; it += NT) {
            const int lane = it % NL, j = it / NL;
            cplx<T> w[R0];
            if (NST > 1) lane_twiddles<R0, S0>(w, tw0, j);
            const Lane<T>* base = s + (G::slot(j) * NL + lane);
            Lane<T> v[R0];
#pragma unroll
            for (int r = 0; r < R0; ++r) v[r] = base[r * LEG0];
            if (NST > 1) {
#pragma unroll
                for (int r = 1; r < R0; ++r) v[r] = lmulc(v[r], w[r]);
            }
            double accd[LPT];
#pragma unroll
            for (int l = 0; l < LPT; ++l) accd[l] = 0.0;
            if (out_lo) {
                lbfly<R0, true, T>(v);
#pragma unroll
                for (int r = 0; r < R0 / 2; ++r) {
#pragma unroll
                    for (int l = 0; l < LPT; ++l) store_elem(lane * LPT + l, j + r * S0, lane_get(v[r], l), accd[l]);
                }
            } else {
                lbfly<R0, true, T>(v);
#pragma unroll
                for (int r = 0; r < R0; ++r) {
#pragma unroll
                    for (int l = 0; l < LPT; ++l) store_elem(lane * LPT + l, j + r * S0, lane_get(v[r], l), accd[l]);
                }
            }
            if (want_dot) {
#pragma unroll
                for (int l = 0; l < LPT; ++l) scratch[(lane * LPT + l) * S0 + j] = accd[l];
            }
        }
    }
    if (want_dot) {
        __syncthreads();
        rows_reduce_partials(scratch, S0, nl, g0, P.st.partial, tid, NT);
        pcg_finalize(P.st, P.dot_kind, g0, g1, P.nrows, tid, NT);
    }
}

}  // namespace hipgp
