// fast_kernels.cuh -- compile-time specialised versions of the three matvec passes on the lane engine (lane_fft.cuh).
//
// Same math and the same digit-reversed ordering as conv_kernels.cuh (the generic, runtime-radix kernels stay as the
// fallback for lengths without an instantiation).  What is different:
//   * a thread owns (butterfly, lane): 16-byte lanes (two fp32 lines / one fp64 line), LDS.128 / STS.128 only, packed
//     f32x2 arithmetic across the two lines of a lane, all butterfly legs at [base + immediate];
//   * the frequency workspace is kept in LANE layout (lane_fft.cuh), so the column pass moves lanes between global
//     memory and registers with plain 16-byte accesses;
//   * the first forward stage reads its operands straight from global memory (zero padding = skipped loads and a
//     pruned butterfly when the upper half of the inputs is padding; PCG vector updates fused into the row loads) and
//     the last inverse stage writes straight to global memory (crop = skipped stores, dot products fused);
//   * column pass: last forward stage, spectrum multiply and first inverse stage run back to back in registers; the
//     spectrum tile is fetched with cp.async into shared memory at kernel start, so its latency hides behind the
//     forward stages;
//   * row passes: the r2c split / c2r merge works on QUADS -- two neighbouring bins and their two mirror bins -- so
//     one thread turns four shared-memory lanes into 16-byte global accesses (the few bins of digit group 0, whose
//     mirrors are irregular, go through a scalar path);
//   * tiles are small (64-110 KB) so that two CTAs are resident per SM and one CTA's global traffic overlaps the
//     other's butterflies.
#pragma once
#include "lane_fft.cuh"

namespace hipgp {

template <class T> __device__ __forceinline__ void ld2(const T* p, bool vec, T& a, T& b, bool ok0, bool ok1) {
    if (vec && ok1) { const cplx<T> t = ld_stream(reinterpret_cast<const cplx<T>*>(p)); a = t.x; b = t.y; }
    else { a = ok0 ? ld_stream(p) : (T)0; b = ok1 ? ld_stream(p + 1) : (T)0; }
}
template <class T> __device__ __forceinline__ void st2(T* p, bool vec, T a, T b, bool ok0, bool ok1) {
    if (vec && ok1) { st_stream(reinterpret_cast<cplx<T>*>(p), mk<T>(a, b)); }
    else { if (ok0) st_stream(p, a); if (ok1) st_stream(p + 1, b); }
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

// spectrum factor(s) of one lane, read from global memory: `idx` = index of the lane's first line in the spectrum
__device__ __forceinline__ Lane<float> lane_spec(Lane<float> v, const void* spec, int kind, size_t idx) {
    if (kind == SPEC_REAL) {
        const cplx<float> sv = ld_stream(reinterpret_cast<const cplx<float>*>(reinterpret_cast<const float*>(spec) + idx));   // two reals
        return lmul_real2(v, sv.x, sv.y);
    }
    const cplx<float>* sp = reinterpret_cast<const cplx<float>*>(spec) + idx;
    const cplx<float> w0 = ld_stream(sp), w1 = ld_stream(sp + 1);
    return kind == SPEC_CPLX ? lmul_cplx2<false>(v, w0, w1) : lmul_cplx2<true>(v, w0, w1);
}
__device__ __forceinline__ Lane<double> lane_spec(Lane<double> v, const void* spec, int kind, size_t idx) {
    if (kind == SPEC_REAL) return lscale(v, ld_stream(reinterpret_cast<const double*>(spec) + idx));
    const cplx<double> w = ld_stream(reinterpret_cast<const cplx<double>*>(spec) + idx);
    return kind == SPEC_CPLX ? lmul(v, w) : lmulc(v, w);
}
// real spectrum factor(s) of one lane staged in shared memory (8 bytes per lane)
__device__ __forceinline__ Lane<float> lane_spec_smem(Lane<float> v, const void* p) { const cplx<float> sv = *reinterpret_cast<const cplx<float>*>(p); return lmul_real2(v, sv.x, sv.y); }
__device__ __forceinline__ Lane<double> lane_spec_smem(Lane<double> v, const void* p) { return lscale(v, *reinterpret_cast<const double*>(p)); }

// =====================================================================================================
// Column pass.  A CTA owns NL lanes (= NL * LPT neighbouring lines) of one (outer, batch) slice.
// Dynamic shared memory: the tile, then (P.spec_stage) NL * 8 bytes per padded position for the real spectrum tile.
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
    const size_t pitch = (size_t)P.pitch;
    const size_t spitch = P.spec_pitch ? (size_t)P.spec_pitch : (size_t)P.pitch;

    if constexpr (NST == 1) {
        // the whole line lives in one thread's registers
        for (int lane = tid; lane < NL; lane += NT) {
            if ((long)lane * LPT >= nvalid) continue;
            const int rows_in = mode == CM_INV ? Ln : P.n_in;
            const int rows_out = mode == CM_FWD ? Ln : P.n_out;
            Lane<T> v[R0];
#pragma unroll
            for (int r = 0; r < R0; ++r)
                v[r] = r < rows_in ? lane_from_global(in + cols_rowoff<T>(r, P.pitch, P.in_split_len, P.in_split_stride) + lane * LPT) : lzero<T>();
            if (mode != CM_INV) lbfly<R0, false, T>(v);
            if (mode == CM_FUSED) {
#pragma unroll
                for (int r = 0; r < R0; ++r) v[r] = lane_spec(v[r], P.spec, P.spec_kind, (size_t)r * spitch + c0 + lane * LPT);
            }
            if (mode != CM_FWD) lbfly<R0, true, T>(v);
#pragma unroll
            for (int r = 0; r < R0; ++r)
                if (r < rows_out) lane_to_global(out + cols_rowoff<T>(r, P.pitch, P.out_split_len, P.out_split_stride) + lane * LPT, v[r]);
        }
        return;
    } else {
        constexpr int LEG0 = G::leg(S0);
        constexpr int SPEC_LANE = 8;                                  // bytes of real spectrum per lane (2 x fp32 or 1 x fp64)
        unsigned char* sspec = smem_raw + G::smem_bytes();
        auto sslot = [](int p) { return p + (p >> G::LOGRL); };       // padded position of the staged spectrum
        const cplx<T>* tw0 = P.f.twst + P.f.twoff[0];
        const bool spec_smem = P.spec_stage != 0;

        // ---- spectrum tile -> shared memory, asynchronously (consumed after the forward stages) ----
        if (spec_smem) {
            const unsigned char* sp = reinterpret_cast<const unsigned char*>(P.spec) + (size_t)c0 * sizeof(T);
            constexpr int ROWB = NL * SPEC_LANE;                      // bytes per position
            constexpr int CH = ROWB >= 16 ? 16 : 8, NCH = ROWB / CH;
            constexpr int NIT = (Ln * NCH + NT - 1) / NT;
#pragma unroll
            for (int k = 0; k < NIT; ++k) {
                const int w = tid + k * NT;
                const int p = w / NCH, c = w - p * NCH;
                if (w < Ln * NCH) cp_async<CH>(sspec + (size_t)sslot(p) * ROWB + c * CH, sp + (size_t)p * spitch * sizeof(T) + c * CH);
            }
            cp_async_commit();
        }

        // ---- first forward stage, operands straight from global memory (zero padding = skipped loads) ----
        if (mode != CM_INV) {
            const int n_in = P.n_in;
            const bool zero_hi = is_pow2(R0) && n_in <= Ln / 2;
            const size_t rstep = (size_t)S0 * pitch;
#pragma unroll 1
            for (int it = tid; it < S0 * NL; it += NT) {
                const int lane = it % NL, j = it / NL;
                const bool ok = (long)lane * LPT < nvalid;
                const cplx<T>* gp = in + (size_t)j * pitch + lane * LPT;
                cplx<T> w[R0];
                lane_twiddles<R0, S0>(w, tw0, j);
                Lane<T> v[R0];
                if (zero_hi) {
#pragma unroll
                    for (int r = 0; r < R0 / 2; ++r) v[r] = (ok && j + r * S0 < n_in && !(P.dbg & 2)) ? lane_from_global(gp + r * rstep) : lzero<T>();
                    lbfly_zero_hi<R0, T>(v);
                } else {
#pragma unroll
                    for (int r = 0; r < R0; ++r) v[r] = (ok && j + r * S0 < n_in) ? lane_from_global(gp + r * rstep) : lzero<T>();
                    lbfly<R0, false, T>(v);
                }
#pragma unroll
                for (int r = 1; r < R0; ++r) v[r] = lmul(v[r], w[r]);
                Lane<T>* base = s + (G::slot(j) * NL + lane);
#pragma unroll
                for (int r = 0; r < R0; ++r) base[r * LEG0] = v[r];
            }
            __syncthreads();
            if (!(P.dbg & 1)) LaneMidFwd<G, T, NL, NT, Ln / R0, 1, Rs...>::run(s, P.f, tid);
        }
        if (spec_smem) { cp_async_wait_all(); __syncthreads(); }

        // ---- last forward stage + spectrum + first inverse stage: RLAST neighbouring positions, in registers ----
        if (!(P.dbg & 1)) {
#pragma unroll 1
            for (int it = tid; it < (Ln / RLAST) * NL; it += NT) {
                const int lane = it % NL, bf = it / NL;
                const bool ok = (long)lane * LPT < nvalid;
                const int p0 = bf * RLAST;
                Lane<T>* base = s + (G::slot(p0) * NL + lane);
                Lane<T> v[RLAST];
                if (mode == CM_INV) {
                    // (slab grids: rows may be gathered in blocks of in_split_len positions, a multiple of RLAST)
                    const cplx<T>* gp = in + cols_rowoff<T>(p0, P.pitch, P.in_split_len, P.in_split_stride) + lane * LPT;
#pragma unroll
                    for (int r = 0; r < RLAST; ++r) v[r] = ok ? lane_from_global(gp + r * pitch) : lzero<T>();
                } else {
#pragma unroll
                    for (int r = 0; r < RLAST; ++r) v[r] = base[r * NL];
                    lbfly<RLAST, false, T>(v);
                }
                if (mode == CM_FUSED) {
                    if (spec_smem) {
                        const unsigned char* sp = sspec + ((size_t)sslot(p0) * NL + lane) * SPEC_LANE;
#pragma unroll
                        for (int r = 0; r < RLAST; ++r) v[r] = lane_spec_smem(v[r], sp + r * NL * SPEC_LANE);
                    } else if (ok) {
                        const size_t sidx = (size_t)p0 * spitch + c0 + lane * LPT;
#pragma unroll
                        for (int r = 0; r < RLAST; ++r) v[r] = lane_spec(v[r], P.spec, P.spec_kind, sidx + (size_t)r * spitch);
                    }
                }
                if (mode == CM_FWD) {
                    if (ok) {
                        cplx<T>* gp = out + cols_rowoff<T>(p0, P.pitch, P.out_split_len, P.out_split_stride) + lane * LPT;
#pragma unroll
                        for (int r = 0; r < RLAST; ++r) lane_to_global(gp + r * pitch, v[r]);
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
        if (!(P.dbg & 1)) LaneMidInv<G, T, NL, NT, Ln / R0, 1, Rs...>::run(s, P.f, tid);
        {
            const int n_out = P.n_out;
            const bool out_lo = is_pow2(R0) && n_out <= Ln / 2;
            const size_t rstep = (size_t)S0 * pitch;
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
                cplx<T>* gp = out + (size_t)j * pitch + lane * LPT;
                if (out_lo) {          // outputs R0/2 .. R0-1 are cropped: the compiler drops their arithmetic
                    lbfly<R0, true, T>(v);
                    if (ok && !(P.dbg & 4)) {
#pragma unroll
                        for (int r = 0; r < R0 / 2; ++r) if (j + r * S0 < n_out) lane_to_global(gp + r * rstep, v[r]);
                    }
                } else {
                    lbfly<R0, true, T>(v);
                    if (ok) {
#pragma unroll
                        for (int r = 0; r < R0; ++r) if (j + r * S0 < n_out) lane_to_global(gp + r * rstep, v[r]);
                    }
                }
            }
        }
    }
}

// =====================================================================================================
// Row passes.  H = product of the radix list; a CTA owns NL lanes = NL * LPT rows.
// =====================================================================================================
struct __align__(8) QuadIdx { int a, b; };     // even positions of a quad: bins (a, a+1) mirror bins (b+1, b)

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
#pragma unroll 1
        for (int it = tid; it < S0 * NL; it += NT) {
            const int lane = it % NL, j = it / NL;
            cplx<T> w[R0];
            if (NST > 1) lane_twiddles<R0, S0>(w, tw0, j);
            double accd[LPT];
            size_t off[LPT];
            bool rok[LPT];
            T coef[LPT];
#pragma unroll
            for (int l = 0; l < LPT; ++l) {
                const int row = lane * LPT + l;
                accd[l] = 0.0; rok[l] = row < nl; off[l] = (size_t)(g0 + row) * n + 2 * j; coef[l] = s_coef[row];
            }
            // element loaders: packed complex e = j + r*S0  ->  x[2e] + i x[2e+1] of line l, with the fused vector update
            auto ld_plain = [&](int l, int r) -> cplx<T> {
                const int i = 2 * (j + r * S0); const bool ok0 = rok[l] && i < n, ok1 = rok[l] && i + 1 < n;
                T a = 0, b2 = 0;
                if (ok0) ld2(P.in + off[l] + 2 * r * S0, vec, a, b2, ok0, ok1);
                return mk<T>(a, b2);
            };
            auto ld_pupdate = [&](int l, int r) -> cplx<T> {
                const int i = 2 * (j + r * S0); const bool ok0 = rok[l] && i < n, ok1 = rok[l] && i + 1 < n;
                T a = 0, b2 = 0;
                if (ok0) {
                    const size_t o = off[l] + 2 * r * S0;
                    ld2(P.in + o, vec, a, b2, ok0, ok1);
                    if (!first_it) { T p0, p1; ld2((const T*)P.v0 + o, vec, p0, p1, ok0, ok1); a = a + coef[l] * p0; b2 = b2 + coef[l] * p1; }
                    st2(P.v0 + o, vec, a, b2, ok0, ok1);
                }
                return mk<T>(a, b2);
            };
            auto ld_selfdot = [&](int l, int r) -> cplx<T> {
                const int i = 2 * (j + r * S0); const bool ok0 = rok[l] && i < n, ok1 = rok[l] && i + 1 < n;
                T a = 0, b2 = 0;
                if (ok0) { ld2(P.in + off[l] + 2 * r * S0, vec, a, b2, ok0, ok1); accd[l] += (double)(a * a) + (double)(b2 * b2); }
                return mk<T>(a, b2);
            };
            auto ld_xrupdate = [&](int l, int r) -> cplx<T> {
                const int i = 2 * (j + r * S0); const bool ok0 = rok[l] && i < n, ok1 = rok[l] && i + 1 < n;
                T a = 0, b2 = 0;
                if (ok0) {
                    const size_t o = off[l] + 2 * r * S0;
                    T p0, p1, x0, x1, r0, r1, q0, q1;
                    ld2(P.v2 + o, vec, p0, p1, ok0, ok1);
                    ld2((const T*)P.v1 + o, vec, x0, x1, ok0, ok1);
                    ld2((const T*)P.v0 + o, vec, r0, r1, ok0, ok1);
                    ld2(P.in + o, vec, q0, q1, ok0, ok1);
                    st2(P.v1 + o, vec, x0 + coef[l] * p0, x1 + coef[l] * p1, ok0, ok1);
                    a = r0 - coef[l] * q0; b2 = ok1 ? r1 - coef[l] * q1 : (T)0;
                    st2(P.v0 + o, vec, a, b2, ok0, ok1);
                    accd[l] += (double)(a * a) + (double)(b2 * b2);
                }
                return mk<T>(a, b2);
            };
            Lane<T> v[R0];
            // (each line's loader runs exactly once per element: the fused updates have side effects)
            auto fill1 = [&](auto& ld, int r) -> Lane<T> {
                const cplx<T> e0 = ld(0, r);
                if constexpr (LPT == 2) { const cplx<T> e1 = ld(1, r); return lane_make(e0, e1); }
                else return lane_make(e0, e0);
            };
#define HIPGP_FILL(RC, LD)                                                                                  \
            _Pragma("unroll") for (int r = 0; r < (RC); ++r) v[r] = fill1(LD, r);
            if (zero_hi) {
                if (mode == RF_PLAIN) { HIPGP_FILL(R0 / 2, ld_plain) }
                else if (mode == RF_PUPDATE) { HIPGP_FILL(R0 / 2, ld_pupdate) }
                else if (mode == RF_XRUPDATE) { HIPGP_FILL(R0 / 2, ld_xrupdate) }
                else { HIPGP_FILL(R0 / 2, ld_selfdot) }
                lbfly_zero_hi<R0, T>(v);
            } else {
                auto ld_any = [&](int l, int r) -> cplx<T> {
                    return mode == RF_PLAIN ? ld_plain(l, r) : (mode == RF_PUPDATE ? ld_pupdate(l, r) : (mode == RF_XRUPDATE ? ld_xrupdate(l, r) : ld_selfdot(l, r)));
                };
                HIPGP_FILL(R0, ld_any)
                lbfly<R0, false, T>(v);
            }
#undef HIPGP_FILL
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

    // ---- r2c split, straight to global memory ----
    // (a) pairs (k, H-k) of digit group 0 (all pairs when the list has a single stage): scalar W accesses
    constexpr int NPAIR0 = NST > 1 ? RLAST / 2 + 1 : H / 2 + 1;
    constexpr int NQUAD = NST > 1 ? (H - RLAST) / 4 : 0;
#pragma unroll 1
    for (int it = tid; it < NPAIR0 * NL; it += NT) {
        const int lane = it % NL, pi = it / NL;
        const QuadIdx pq = reinterpret_cast<const QuadIdx*>(P.pairs)[pi];
        const int q = pq.a, q2 = pq.b;
        Lane<T> Xa, Xb;
        bool two;
        if (q == 0) {
            const Lane<T> z = s[lane];
            Lane<T> zs; zs.re = z.im; zs.im = z.re;
            Xa = lscale(z + zs, (T)2); Xb = lscale(z - zs, (T)2);       // .re = 2 (re +- im); imaginary parts are zero
            Xa.im = lzero<T>().im; Xb.im = lzero<T>().im;
            two = true;
        } else {
            const cplx<T> wq = ldg_c(P.pairw + pi);
            const Lane<T> a = s[G::slot(q) * NL + lane], c = lconj(s[G::slot(q2) * NL + lane]);
            const Lane<T> E = a + c;
            const Lane<T> t = lmul(lmi<false>(a - c), wq);
            Xa = E + t; Xb = lconj(E - t);
            two = q2 != q;
        }
#pragma unroll
        for (int l = 0; l < LPT; ++l) {
            const int row = lane * LPT + l;
            if (row < nl) {
                cplx<T>* dst = P.W + s_wbase[row];
                WRow<T>::store1(dst, q, lane_get(Xa, l));
                if (two) WRow<T>::store1(dst, q2, lane_get(Xb, l));
            }
        }
    }
    // (b) quads: bins (a, a+1) and their mirrors (b+1, b): four lanes in, 16-byte stores out.  The table loads of
    //     all of a thread's quads are issued up front.
    if constexpr (NQUAD > 0) {
        constexpr int QIT = (NQUAD * NL + NT - 1) / NT;
        QuadIdx qq[QIT]; cplx<T> wa[QIT], wa1[QIT];
#pragma unroll
        for (int k = 0; k < QIT; ++k) {
            const int it = tid + k * NT, qi = it / NL;
            if (it < NQUAD * NL) { qq[k] = reinterpret_cast<const QuadIdx*>(P.quadq)[qi]; wa[k] = ldg_c(P.quadw + 2 * qi); wa1[k] = ldg_c(P.quadw + 2 * qi + 1); }
        }
#pragma unroll
        for (int k = 0; k < QIT; ++k) {
            const int it = tid + k * NT, lane = it % NL;
            if (it < NQUAD * NL) {
                const Lane<T>* pa = s + (G::slot(qq[k].a) * NL + lane);
                const Lane<T>* pb = s + (G::slot(qq[k].b) * NL + lane);
                const Lane<T> za = pa[0], za1 = pa[NL], zb = pb[0], zb1 = pb[NL];
                // bin a with mirror b+1
                Lane<T> c = lconj(zb1);
                Lane<T> E = za + c;
                Lane<T> t = lmul(lmi<false>(za - c), wa[k]);
                const Lane<T> Xa = E + t, Xb1 = lconj(E - t);
                // bin a+1 with mirror b
                c = lconj(zb);
                E = za1 + c;
                t = lmul(lmi<false>(za1 - c), wa1[k]);
                const Lane<T> Xa1 = E + t, Xb = lconj(E - t);
                const bool self = qq[k].a == qq[k].b;    // the quad mirrors onto itself: bins (a, a+1) are each other's mirror
#pragma unroll
                for (int l = 0; l < LPT; ++l) {
                    const int row = lane * LPT + l;
                    if (row < nl) {
                        cplx<T>* dst = P.W + s_wbase[row];
                        if (self) WRow<T>::store2(dst, qq[k].a, lane_get(Xa, l), lane_get(Xb1, l));
                        else { WRow<T>::store2(dst, qq[k].a, lane_get(Xa, l), lane_get(Xa1, l)); WRow<T>::store2(dst, qq[k].b, lane_get(Xb, l), lane_get(Xb1, l)); }
                    }
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

    // ---- c2r merge, straight from global memory into shared memory ----
    const int spec_kind = P.spec_kind;
    constexpr int NPAIR0 = NST > 1 ? RLAST / 2 + 1 : H / 2 + 1;
    constexpr int NQUAD = NST > 1 ? (H - RLAST) / 4 : 0;
    // (a) pairs of digit group 0 (all pairs when the list has a single stage)
#pragma unroll 1
    for (int it = tid; it < NPAIR0 * NL; it += NT) {
        const int lane = it % NL, pi = it / NL;
        const QuadIdx pq = reinterpret_cast<const QuadIdx*>(P.pairs)[pi];
        const int q = pq.a, q2 = pq.b;
        const cplx<T> wq = ldg_c(P.pairw + pi);
        cplx<T> ya[LPT], yc[LPT];
#pragma unroll
        for (int l = 0; l < LPT; ++l) {
            const int row = lane * LPT + l;
            ya[l] = mk<T>(0, 0); yc[l] = mk<T>(0, 0);
            if (row < nl) {
                const cplx<T>* src = P.W + s_wbase[row];
                ya[l] = WRow<T>::load1(src, q); yc[l] = WRow<T>::load1(src, q2);
                if (spec_kind != SPEC_NONE) { ya[l] = apply_spec(ya[l], P.spec, spec_kind, (size_t)q); yc[l] = apply_spec(yc[l], P.spec, spec_kind, (size_t)q2); }
            }
        }
        const Lane<T> a = lane_make(ya[0], ya[LPT - 1]);
        Lane<T> c = lane_make(yc[0], yc[LPT - 1]);
        if (q == 0) {
            // Z[0] = (Y0 + YH) + i (Y0 - YH), real parts only
            Lane<T> z; z.re = (a + c).re; z.im = (a - c).re;
            s[lane] = z;
        } else {
            // bin q (frequency k):  E + i O with O = conj(w^k)(Y[k] - conj Y[k']);  bin q2 is conj(E - i O)
            c = lconj(c);
            const Lane<T> E = a + c;
            const Lane<T> iO = lmi<true>(lmulc(a - c, wq));
            s[G::slot(q) * NL + lane] = E + iO;
            if (q2 != q) s[G::slot(q2) * NL + lane] = lconj(E - iO);
        }
    }
    // (b) quads: table and workspace loads of all of a thread's quads are issued up front
    if constexpr (NQUAD > 0) {
        constexpr int QIT = (NQUAD * NL + NT - 1) / NT;
        QuadIdx qq[QIT]; cplx<T> wa[QIT], wa1[QIT];
#pragma unroll
        for (int k = 0; k < QIT; ++k) {
            const int it = tid + k * NT, qi = it / NL;
            if (it < NQUAD * NL) { qq[k] = reinterpret_cast<const QuadIdx*>(P.quadq)[qi]; wa[k] = ldg_c(P.quadw + 2 * qi); wa1[k] = ldg_c(P.quadw + 2 * qi + 1); }
        }
#pragma unroll
        for (int k = 0; k < QIT; ++k) {
            const int it = tid + k * NT, lane = it % NL;
            if (it < NQUAD * NL) {
                const bool self = qq[k].a == qq[k].b;
                cplx<T> ya[LPT], ya1[LPT], yb[LPT], yb1[LPT];
#pragma unroll
                for (int l = 0; l < LPT; ++l) {
                    const int row = lane * LPT + l;
                    ya[l] = ya1[l] = yb[l] = yb1[l] = mk<T>(0, 0);
                    if (row < nl) {
                        const cplx<T>* src = P.W + s_wbase[row];
                        WRow<T>::load2(src, qq[k].a, ya[l], ya1[l]);
                        if (self) { yb[l] = ya[l]; yb1[l] = ya1[l]; } else WRow<T>::load2(src, qq[k].b, yb[l], yb1[l]);
                        if (spec_kind != SPEC_NONE) {
                            ya[l] = apply_spec(ya[l], P.spec, spec_kind, (size_t)qq[k].a); ya1[l] = apply_spec(ya1[l], P.spec, spec_kind, (size_t)qq[k].a + 1);
                            yb[l] = apply_spec(yb[l], P.spec, spec_kind, (size_t)qq[k].b); yb1[l] = apply_spec(yb1[l], P.spec, spec_kind, (size_t)qq[k].b + 1);
                        }
                    }
                }
                const Lane<T> A = lane_make(ya[0], ya[LPT - 1]), A1 = lane_make(ya1[0], ya1[LPT - 1]);
                const Lane<T> Bn = lane_make(yb[0], yb[LPT - 1]), B1 = lane_make(yb1[0], yb1[LPT - 1]);
                Lane<T>* pa = s + (G::slot(qq[k].a) * NL + lane);
                Lane<T>* pb = s + (G::slot(qq[k].b) * NL + lane);
                // bin a with mirror b+1
                Lane<T> c = lconj(B1);
                Lane<T> E = A + c;
                Lane<T> iO = lmi<true>(lmulc(A - c, wa[k]));
                pa[0] = E + iO;
                pb[NL] = lconj(E - iO);
                if (!self) {   // bin a+1 with mirror b
                    c = lconj(Bn);
                    E = A1 + c;
                    iO = lmi<true>(lmulc(A1 - c, wa1[k]));
                    pa[NL] = E + iO;
                    pb[0] = lconj(E - iO);
                }
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
            size_t off[LPT];
            bool rok[LPT];
#pragma unroll
            for (int l = 0; l < LPT; ++l) { const int row = lane * LPT + l; accd[l] = 0.0; rok[l] = row < nl; off[l] = (size_t)(g0 + row) * n + 2 * j; }
            auto st_elem = [&](int l, int r, cplx<T> val, bool dot) {
                const int i = 2 * (j + r * S0);
                const bool ok0 = rok[l] && i < n, ok1 = rok[l] && i + 1 < n;
                if (ok0) {
                    const size_t o = off[l] + 2 * r * S0;
                    st2(P.out + o, vec, val.x, val.y, ok0, ok1);
                    if (dot) {
                        T o0, o1; ld2((const T*)P.v0 + o, vec, o0, o1, ok0, ok1);
                        accd[l] += (double)(val.x * o0) + (ok1 ? (double)(val.y * o1) : 0.0);
                    }
                }
            };
            lbfly<R0, true, T>(v);
#define HIPGP_DRAIN(RC, DOT)                                                                                \
            _Pragma("unroll") for (int r = 0; r < (RC); ++r) { _Pragma("unroll") for (int l = 0; l < LPT; ++l) st_elem(l, r, lane_get(v[r], l), DOT); }
            if (out_lo) { if (want_dot) { HIPGP_DRAIN(R0 / 2, true) } else { HIPGP_DRAIN(R0 / 2, false) } }
            else { if (want_dot) { HIPGP_DRAIN(R0, true) } else { HIPGP_DRAIN(R0, false) } }
#undef HIPGP_DRAIN
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
