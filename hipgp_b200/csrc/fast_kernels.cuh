// fast_kernels.cuh -- compile-time specialised versions of the three matvec passes.
//
// Same math and the same digit-reversed layouts as conv_kernels.cuh (the generic, runtime-radix kernels stay as the
// fallback for lengths without an instantiation), but:
//   * the radix list is a template parameter pack, so all index arithmetic is shifts / constant divisions;
//   * radix-16 butterflies (4x4) cut the number of shared-memory sweeps;
//   * a thread owns a BUTTERFLY and walks the lines of the tile with it: positions and twiddles (read once from a
//     per-stage table laid out [r][j], i.e. coalesced) are loop invariants, so per line only the 2R shared-memory
//     accesses and the butterfly arithmetic remain;
//   * fp32 complex arithmetic uses Blackwell's packed FADD2/FMUL2/FFMA2 (fft_engine.cuh);
//   * row passes: the first forward stage reads its operands straight from global memory (zero padding = predicated
//     loads, PCG vector updates fused into those loads) and the last inverse stage writes straight to global memory
//     (crop = predicated stores, dot products fused into those stores);
//   * column pass: last forward stage, spectrum multiply and first inverse stage run back to back in registers; the
//     spectrum is stored transposed ([line][position]) so that those reads are contiguous per thread;
//   * shared memory is line-major, s[line][pad(pos)], with a padding function that keeps the strided stage accesses
//     spread over the banks.
#pragma once
#include "conv_kernels.cuh"

namespace hipgp {

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
__device__ __forceinline__ void cp_async_wait_all() {
#ifndef HIPGP_EMU
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#endif
}

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

// padding of a position inside a line: keeps stride-2^k butterflies of neighbouring threads on distinct banks
template <class T> __host__ __device__ constexpr int rpad(int pos) {
    return sizeof(T) == 4 ? pos + (pos >> 4) : pos + (pos >> 3) + (pos >> 6);
}
template <class T> __host__ __device__ constexpr int line_stride(int L) {
    // == 2 (mod 16) complex slots: in the column pass the TB lines of one position (and the next position) fall on
    // disjoint banks during the lines-fast copy in / copy out
    return ((rpad<T>(L) + 15) / 16) * 16 + 2;
}
__host__ __device__ constexpr bool is_pow2(int x) { return (x & (x - 1)) == 0; }
// padded offset of element r of a butterfly starting at p0 with stride S inside a sub-transform of length Nt.
// For power-of-two Nt the padding is additive (no carries between p0's and r*S's low bits), so the offsets relative
// to rpad(p0) are compile-time constants and every access is [register + immediate].
template <class T, int Nt, int S> __device__ __forceinline__ int boff(int rp0, int p0, int r) {
    return is_pow2(Nt) ? rp0 + rpad<T>(r * S) : rpad<T>(p0 + r * S);
}

// per-stage twiddles w^{j r}, r = 1..R-1 (conjugated for the inverse), from the [r][j] table
template <int R, int S, bool INV, class T>
__device__ __forceinline__ void load_twiddles(cplx<T>* w, const cplx<T>* __restrict__ tab, int j) {
#pragma unroll
    for (int r = 1; r < R; ++r) { w[r] = ldg_c(tab + (r - 1) * S + j); if (INV) w[r] = conj(w[r]); }
}

// ---- one in-place shared-memory stage over `nlines` lines, compile-time geometry -------------------------
// thread -> (butterfly bf, line group g); the thread applies its butterfly to lines g, g+G, ...
template <class T, int Ln, int Nt, int R, bool INV>
__device__ __forceinline__ void smem_stage(cplx<T>* s, int RS, int nlines, const cplx<T>* __restrict__ twtab, int tid, int nthreads) {
    constexpr int S = Nt / R, NB = Ln / R;
    auto process = [&](int bf, int g, int G) {
        const int blk = bf / S, j = bf - blk * S;
        const int p0 = blk * Nt + j;
        const int rp0 = rpad<T>(p0);
        int o[R];
#pragma unroll
        for (int r = 0; r < R; ++r) o[r] = boff<T, Nt, S>(0, p0, r) - (is_pow2(Nt) ? 0 : rp0);
        cplx<T> w[R];
        const bool tw_on = (S > 1) && (j != 0);
        if (S > 1) load_twiddles<R, S, INV>(w, twtab, j);
        for (int line = g; line < nlines; line += G) {
            cplx<T>* base = s + (line * RS + rp0);
            cplx<T> v[R];
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = base[o[r]];
            if (INV) {
                if (tw_on) {
#pragma unroll
                    for (int r = 1; r < R; ++r) v[r] = v[r] * w[r];
                }
                bfly<R, true>(v);
            } else {
                bfly<R, false>(v);
                if (tw_on) {
#pragma unroll
                    for (int r = 1; r < R; ++r) v[r] = v[r] * w[r];
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) base[o[r]] = v[r];
        }
    };
    if (nthreads >= NB) {
        const int G = nthreads / NB, g = tid / NB;
        if (g < G) process(tid % NB, g, G);
    } else {
        for (int bf = tid; bf < NB; bf += nthreads) process(bf, 0, 1);
    }
}

// middle DIF stages (all but the first and the last of the list), each followed by a barrier.  STG = stage index.
template <class T, int Ln, int Nt, int STG, int... Rs> struct MidFwd;
template <class T, int Ln, int Nt, int STG, int R0> struct MidFwd<T, Ln, Nt, STG, R0> {
    static __device__ __forceinline__ void run(cplx<T>*, int, int, const LineFft<T>&, int, int) {}
};
template <class T, int Ln, int Nt, int STG, int R0, int R1, int... Rs> struct MidFwd<T, Ln, Nt, STG, R0, R1, Rs...> {
    static __device__ __forceinline__ void run(cplx<T>* s, int RS, int nl, const LineFft<T>& f, int tid, int nth) {
        smem_stage<T, Ln, Nt, R0, false>(s, RS, nl, f.twst + f.twoff[STG], tid, nth);
        __syncthreads();
        MidFwd<T, Ln, Nt / R0, STG + 1, R1, Rs...>::run(s, RS, nl, f, tid, nth);
    }
};
template <class T, int Ln, int Nt, int STG, int... Rs> struct MidInv;
template <class T, int Ln, int Nt, int STG, int R0> struct MidInv<T, Ln, Nt, STG, R0> {
    static __device__ __forceinline__ void run(cplx<T>*, int, int, const LineFft<T>&, int, int) {}
};
template <class T, int Ln, int Nt, int STG, int R0, int R1, int... Rs> struct MidInv<T, Ln, Nt, STG, R0, R1, Rs...> {
    static __device__ __forceinline__ void run(cplx<T>* s, int RS, int nl, const LineFft<T>& f, int tid, int nth) {
        MidInv<T, Ln, Nt / R0, STG + 1, R1, Rs...>::run(s, RS, nl, f, tid, nth);
        smem_stage<T, Ln, Nt, R0, true>(s, RS, nl, f.twst + f.twoff[STG], tid, nth);
        __syncthreads();
    }
};

// =====================================================================================================
// Column pass.  Radix list <R0, Rmid..., RLAST>; TB lines per CTA.
// =====================================================================================================
template <class T, int TB, int R0, int... Rs>
__global__ void __launch_bounds__(512) cols_fast_kernel(ColsParams<T> P) {
    using List = RL<R0, Rs...>;
    constexpr int Ln = RLInfo<List>::N;
    constexpr int NST = RLInfo<List>::count;
    constexpr int RLAST = RLLast<List>::value;
    constexpr int RS = line_stride<T>(Ln);
    HIPGP_DYN_SMEM(smem_raw);
    cplx<T>* s = reinterpret_cast<cplx<T>*>(smem_raw);
    const int tid = threadIdx.x, nthreads = blockDim.x;
    if (P.done_flag && *P.done_flag) return;
    const long c0 = (long)blockIdx.x * TB;
    const int nc = (int)(P.inner - c0 < TB ? P.inner - c0 : TB);
    const cplx<T>* in = P.in + (size_t)blockIdx.y * P.in_ostride + (size_t)blockIdx.z * P.in_bstride + c0;
    cplx<T>* out = P.out + (size_t)blockIdx.y * P.out_ostride + (size_t)blockIdx.z * P.out_bstride + c0;
    const long pitch = P.pitch;
    const int mode = P.mode;

    // ---- copy in (lines fast => TB contiguous elements per row), zero padding ----
    {
        const int rows_in = mode == CM_INV ? Ln : P.n_in;
        for (int w = tid; w < Ln * TB; w += nthreads) {
            const int c = w % TB, i = w / TB;
            cplx<T>* d = s + (c * RS + rpad<T>(i));
            if (i < rows_in && c < nc) cp_async<(int)sizeof(cplx<T>)>(d, in + (cols_rowoff<T>(i, pitch, P.in_split_len, P.in_split_stride) + c));
            else *d = mk<T>(0, 0);
        }
        cp_async_wait_all();
        __syncthreads();
    }

    // ---- forward stages except the last ----
    if (mode != CM_INV) {
        if constexpr (NST > 1) MidFwd<T, Ln, Ln, 0, R0, Rs...>::run(s, RS, TB, P.f, tid, nthreads);
    }

    // ---- last forward stage + spectrum + first inverse stage: same RLAST contiguous positions ----
    {
        constexpr int NB = Ln / RLAST;
        auto process = [&](int bf, int g, int G) {
            const int p0 = bf * RLAST;
            const int rp0 = rpad<T>(p0);
            int o[RLAST];
#pragma unroll
            for (int r = 0; r < RLAST; ++r) o[r] = boff<T, RLAST, 1>(0, p0, r) - (is_pow2(RLAST) ? 0 : rp0);
            for (int line = g; line < TB; line += G) {
                cplx<T>* base = s + (line * RS + rp0);
                cplx<T> v[RLAST];
#pragma unroll
                for (int r = 0; r < RLAST; ++r) v[r] = base[o[r]];
                if (mode != CM_INV) bfly<RLAST, false>(v);
                if (mode == CM_FUSED && line < nc) {
                    // transposed spectrum: [line][position], RLAST contiguous values per thread
                    const size_t sidx = (size_t)(c0 + line) * Ln + p0;
                    if (P.spec_kind == SPEC_REAL) {
                        const T* sp = reinterpret_cast<const T*>(P.spec) + sidx;
#pragma unroll
                        for (int r = 0; r < RLAST; ++r) v[r] = v[r] * __ldg(sp + r);
                    } else {
                        const cplx<T>* sp = reinterpret_cast<const cplx<T>*>(P.spec) + sidx;
#pragma unroll
                        for (int r = 0; r < RLAST; ++r) {
                            const cplx<T> sv = ldg_c(sp + r);
                            v[r] = P.spec_kind == SPEC_CPLX ? v[r] * sv : mulc(v[r], sv);
                        }
                    }
                }
                if (mode != CM_FWD) bfly<RLAST, true>(v);
#pragma unroll
                for (int r = 0; r < RLAST; ++r) base[o[r]] = v[r];
            }
        };
        if (nthreads >= NB) {
            const int G = nthreads / NB, g = tid / NB;
            if (g < G) process(tid % NB, g, G);
        } else {
            for (int bf = tid; bf < NB; bf += nthreads) process(bf, 0, 1);
        }
        __syncthreads();
    }

    // ---- inverse stages except the first ----
    if (mode != CM_FWD) {
        if constexpr (NST > 1) MidInv<T, Ln, Ln, 0, R0, Rs...>::run(s, RS, TB, P.f, tid, nthreads);
    }

    // ---- copy out (crop) ----
    {
        const int rows_out = mode == CM_FWD ? Ln : P.n_out;
        for (int w = tid; w < rows_out * TB; w += nthreads) {
            const int c = w % TB, i = w / TB;
            if (c < nc) out[cols_rowoff<T>(i, pitch, P.out_split_len, P.out_split_stride) + c] = s[(size_t)c * RS + rpad<T>(i)];
        }
    }
}

// =====================================================================================================
// Row passes.  H = product of the radix list; a CTA owns RB rows.
// =====================================================================================================
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

template <class T, int R0, int... Rs>
__global__ void __launch_bounds__(256) rows_fwd_fast_kernel(RowsParams<T> P) {
    using List = RL<R0, Rs...>;
    constexpr int H = RLInfo<List>::N;
    constexpr int NST = RLInfo<List>::count;
    constexpr int RLAST = RLLast<List>::value;
    constexpr int S0 = H / R0;
    constexpr int RS = line_stride<T>(H);
    HIPGP_DYN_SMEM(smem_raw);
    const int RB = P.RB;
    cplx<T>* s = reinterpret_cast<cplx<T>*>(smem_raw);
    double* scratch = reinterpret_cast<double*>(smem_raw + sizeof(cplx<T>) * (size_t)RS * RB);
    const int tid = threadIdx.x, nthreads = blockDim.x;
    if (P.mode != RF_PLAIN && P.st.flags[0]) return;
    const long g0 = (long)blockIdx.x * RB;
    const long g1 = g0 + RB < P.total_rows ? g0 + RB : P.total_rows;
    const int nl = (int)(g1 - g0);
    const int n = P.n_real;
    const int mode = P.mode;
    const bool vec = ((n & 1) == 0) && P.vec_ok;             // rows start on even offsets => 2-element accesses are aligned
    const bool first_it = (mode == RF_PUPDATE) && (P.st.flags[2] != 0);
    const bool want_dot = (mode == RF_XRUPDATE || mode == RF_SELFDOT);
    // per-row scalars, computed once (the 64-bit divisions stay out of the element loops)
    __shared__ T s_coef[32];
    __shared__ long s_wbase[32];
    if (tid < RB && tid < nl) {
        const long gr = g0 + tid;
        const long b = gr / P.nrows;
        s_wbase[tid] = (b * P.W_rows + (gr - b * P.nrows)) * P.W_pitch;
        T coef = 0;
        if (mode == RF_PUPDATE) coef = first_it ? (T)0 : (T)(P.st.zr[b] / P.st.zr_prev[b]);
        else if (mode == RF_XRUPDATE) coef = (T)(P.st.zr[b] / P.st.pAp[b]);
        s_coef[tid] = coef;
    }
    __syncthreads();

    // ---- first DIF stage fused with the load (and the PCG vector update); thread = butterfly j, loops over rows ----
    {
        constexpr int NI = S0;                         // partial slots per row
        auto process = [&](int j, int g, int G) {
            cplx<T> w[R0];
            if (NST > 1) load_twiddles<R0, S0, false>(w, P.f.twst + P.f.twoff[0], j);
            const int rp0 = rpad<T>(j);
            int o[R0];
#pragma unroll
            for (int r = 0; r < R0; ++r) o[r] = boff<T, H, S0>(0, j, r) - (is_pow2(H) ? 0 : rp0);
            for (int row = g; row < RB; row += G) {
                cplx<T> v[R0];
                double accd = 0.0;
                if (row < nl) {
                    const size_t off = (size_t)(g0 + row) * n;
                    const T coef = s_coef[row];
#pragma unroll
                    for (int r = 0; r < R0; ++r) {
                        const int i = 2 * (j + r * S0);
                        const bool ok0 = i < n, ok1 = i + 1 < n;
                        T a = 0, b2 = 0;
                        if (ok0) {
                            if (mode == RF_PLAIN) {
                                ld2(P.in + off + i, vec, a, b2, ok0, ok1);
                            } else if (mode == RF_PUPDATE) {
                                T z0, z1; ld2(P.in + off + i, vec, z0, z1, ok0, ok1);
                                if (first_it) { a = z0; b2 = z1; }
                                else { T p0, p1; ld2((const T*)P.v0 + off + i, vec, p0, p1, ok0, ok1); a = z0 + coef * p0; b2 = z1 + coef * p1; }
                                st2(P.v0 + off + i, vec, a, b2, ok0, ok1);
                            } else if (mode == RF_SELFDOT) {
                                ld2(P.in + off + i, vec, a, b2, ok0, ok1);
                                accd += (double)(a * a) + (double)(b2 * b2);
                            } else {
                                T p0, p1, x0, x1, r0, r1, q0, q1;
                                ld2(P.v2 + off + i, vec, p0, p1, ok0, ok1);
                                ld2((const T*)P.v1 + off + i, vec, x0, x1, ok0, ok1);
                                ld2((const T*)P.v0 + off + i, vec, r0, r1, ok0, ok1);
                                ld2(P.in + off + i, vec, q0, q1, ok0, ok1);
                                st2(P.v1 + off + i, vec, x0 + coef * p0, x1 + coef * p1, ok0, ok1);
                                a = r0 - coef * q0; b2 = ok1 ? r1 - coef * q1 : (T)0;
                                st2(P.v0 + off + i, vec, a, b2, ok0, ok1);
                                accd += (double)(a * a) + (double)(b2 * b2);
                            }
                        }
                        v[r] = mk<T>(a, b2);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < R0; ++r) v[r] = mk<T>(0, 0);
                }
                if (want_dot) scratch[row * NI + j] = accd;
                if (P.do_fft) {
                    bfly<R0, false>(v);
                    if (NST > 1 && j != 0) {
#pragma unroll
                        for (int r = 1; r < R0; ++r) v[r] = v[r] * w[r];
                    }
                    cplx<T>* base = s + (row * RS + rp0);
#pragma unroll
                    for (int r = 0; r < R0; ++r) base[o[r]] = v[r];
                }
            }
        };
        if (nthreads >= S0) {
            const int G = nthreads / S0, g = tid / S0;
            if (g < G) process(tid % S0, g, G);
        } else {
            for (int j = tid; j < S0; j += nthreads) process(j, 0, 1);
        }
    }
    __syncthreads();
    if (want_dot) {
        rows_reduce_partials(scratch, S0, nl, g0, P.st.partial, tid, nthreads);
        pcg_finalize(P.st, mode == RF_XRUPDATE ? DOT_RR : DOT_ZR, g0, g1, P.nrows, tid, nthreads);
    }
    if (!P.do_fft) return;

    if constexpr (NST > 1) {
        MidFwd<T, H, H / R0, 1, Rs...>::run(s, RS, RB, P.f, tid, nthreads);
        smem_stage<T, H, RLAST, RLAST, false>(s, RS, RB, P.f.twst, tid, nthreads);
        __syncthreads();
    }

    // ---- split (pairs k, H-k) straight to global: a warp walks a row; partner / twiddle tables are coalesced ----
    {
        const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
        for (int row = warp; row < nl; row += nwarps) {
            cplx<T>* dst = P.W + s_wbase[row];
            const cplx<T>* base = s + row * RS;
            const cplx<T> z0 = base[0];
            for (int q = lane; q <= H; q += 32) {
                cplx<T> X;
                if (q == H) {
                    X = mk<T>((T)2 * (z0.x - z0.y), 0);
                } else if (q == 0) {
                    X = mk<T>((T)2 * (z0.x + z0.y), 0);
                } else {
                    const int q2 = P.part[q];
                    const cplx<T> a = base[rpad<T>(q)], c = conj(base[rpad<T>(q2)]);
                    const cplx<T> E = a + c, d = a - c;
                    const cplx<T> O = mk<T>(d.y, -d.x);
                    X = E + P.twLp[q] * O;
                }
                dst[q] = X;
            }
        }
    }
}

template <class T, int R0, int... Rs>
__global__ void __launch_bounds__(256) rows_inv_fast_kernel(RowsParams<T> P) {
    using List = RL<R0, Rs...>;
    constexpr int H = RLInfo<List>::N;
    constexpr int NST = RLInfo<List>::count;
    constexpr int RLAST = RLLast<List>::value;
    constexpr int S0 = H / R0;
    constexpr int RS = line_stride<T>(H);
    HIPGP_DYN_SMEM(smem_raw);
    const int RB = P.RB;
    cplx<T>* s = reinterpret_cast<cplx<T>*>(smem_raw);
    double* scratch = reinterpret_cast<double*>(smem_raw + sizeof(cplx<T>) * (size_t)RS * RB);
    const int tid = threadIdx.x, nthreads = blockDim.x;
    if (P.mode != RI_PLAIN && P.st.flags[0]) return;
    const long g0 = (long)blockIdx.x * RB;
    const long g1 = g0 + RB < P.total_rows ? g0 + RB : P.total_rows;
    const int nl = (int)(g1 - g0);
    const int n = P.n_real;

    __shared__ long s_wbase[32];
    if (tid < RB && tid < nl) {
        const long gr = g0 + tid;
        const long b = gr / P.nrows;
        s_wbase[tid] = (b * P.W_rows + (gr - b * P.nrows)) * P.W_pitch;
    }
    __syncthreads();
    // ---- async copy of the H+1 bins of every row into shared memory, then the merge pairwise in place ----
    {
        const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
        for (int row = warp; row < RB; row += nwarps) {
            cplx<T>* base = s + row * RS;
            if (row >= nl) {
                for (int q = lane; q <= H; q += 32) base[rpad<T>(q)] = mk<T>(0, 0);
                continue;
            }
            const cplx<T>* src = P.W + s_wbase[row];
            for (int q = lane; q <= H; q += 32) cp_async<(int)sizeof(cplx<T>)>(base + rpad<T>(q), src + q);
        }
        cp_async_wait_all();
        __syncthreads();
        for (int row = warp; row < nl; row += nwarps) {
            cplx<T>* base = s + row * RS;
            for (int q = lane; q < H; q += 32) {
                if (q == 0) {
                    cplx<T> a = base[0], c = base[rpad<T>(H)];
                    if (P.spec_kind != SPEC_NONE) { a = apply_spec(a, P.spec, P.spec_kind, (size_t)0); c = apply_spec(c, P.spec, P.spec_kind, (size_t)H); }
                    base[0] = mk<T>(a.x + c.x, a.x - c.x);
                    continue;
                }
                const int q2 = P.part[q];
                if (q > q2) continue;                       // the pair is handled by its smaller member
                cplx<T> a = base[rpad<T>(q)], c = base[rpad<T>(q2)];
                if (P.spec_kind != SPEC_NONE) { a = apply_spec(a, P.spec, P.spec_kind, (size_t)q); c = apply_spec(c, P.spec, P.spec_kind, (size_t)q2); }
                // bin q (frequency k):  E + i O with O = conj(w^k)(Y[k] - conj Y[k']);  bin q2 is conj(E - i O)
                c = conj(c);
                const cplx<T> E = a + c;
                const cplx<T> O = mulc(a - c, P.twLp[q]);
                const cplx<T> iO = mk<T>(-O.y, O.x);
                base[rpad<T>(q)] = E + iO;
                if (q != q2) base[rpad<T>(q2)] = conj(E - iO);
            }
        }
    }
    __syncthreads();

    if constexpr (NST > 1) {
        smem_stage<T, H, RLAST, RLAST, true>(s, RS, RB, P.f.twst, tid, nthreads);
        __syncthreads();
        MidInv<T, H, H / R0, 1, Rs...>::run(s, RS, RB, P.f, tid, nthreads);
    }

    // ---- last inverse stage fused with the store (crop) and the dot product ----
    const bool vec = ((n & 1) == 0) && P.vec_ok;
    const bool want_dot = P.mode == RI_DOT;
    {
        auto process = [&](int j, int g, int G) {
            cplx<T> w[R0];
            if (NST > 1) load_twiddles<R0, S0, true>(w, P.f.twst + P.f.twoff[0], j);
            const int rp0 = rpad<T>(j);
            int o[R0];
#pragma unroll
            for (int r = 0; r < R0; ++r) o[r] = boff<T, H, S0>(0, j, r) - (is_pow2(H) ? 0 : rp0);
            for (int row = g; row < RB; row += G) {
                double accd = 0.0;
                if (row < nl) {
                    const cplx<T>* base = s + (row * RS + rp0);
                    cplx<T> v[R0];
#pragma unroll
                    for (int r = 0; r < R0; ++r) v[r] = base[o[r]];
                    if (NST > 1 && j != 0) {
#pragma unroll
                        for (int r = 1; r < R0; ++r) v[r] = v[r] * w[r];
                    }
                    bfly<R0, true>(v);
                    const size_t off = (size_t)(g0 + row) * n;
#pragma unroll
                    for (int r = 0; r < R0; ++r) {
                        const int i = 2 * (j + r * S0);
                        const bool ok0 = i < n, ok1 = i + 1 < n;
                        if (ok0) {
                            st2(P.out + off + i, vec, v[r].x, v[r].y, ok0, ok1);
                            if (want_dot) {
                                T o0, o1; ld2((const T*)P.v0 + off + i, vec, o0, o1, ok0, ok1);
                                accd += (double)(v[r].x * o0) + (ok1 ? (double)(v[r].y * o1) : 0.0);
                            }
                        }
                    }
                }
                if (want_dot) scratch[row * S0 + j] = accd;
            }
        };
        if (nthreads >= S0) {
            const int G = nthreads / S0, g = tid / S0;
            if (g < G) process(tid % S0, g, G);
        } else {
            for (int j = tid; j < S0; j += nthreads) process(j, 0, 1);
        }
    }
    if (want_dot) {
        __syncthreads();
        rows_reduce_partials(scratch, S0, nl, g0, P.st.partial, tid, nthreads);
        pcg_finalize(P.st, P.dot_kind, g0, g1, P.nrows, tid, nthreads);
    }
}

}  // namespace hipgp
