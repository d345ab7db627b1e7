// fast_launch.cuh -- launch code of the specialised kernels (included only by fast_inst.cu).
#pragma once
#include "plan_types.cuh"
#include "fast_kernels.cuh"

namespace hipgp {
template <class T> constexpr int rows_pad(int pos) { return sizeof(T) == 4 ? pos + (pos >> 4) : pos + (pos >> 3) + (pos >> 6); }

template <class T, int R0, int... Rs>
static void launch_rows_fast_t(hipgp_plan* pl, bool inverse, RowsParams<T>& P, cudaStream_t st) {
    constexpr int H = RLInfo<RL<R0, Rs...>>::N;
    constexpr int S0 = H / R0;
    const int RS = line_stride<T>(H);
    auto smem_for = [&](int rb) { return sizeof(cplx<T>) * (size_t)RS * rb + sizeof(double) * (size_t)S0 * rb; };
    int rb = 16;
    while (rb > 1 && (smem_for(rb) > 72 * 1024 || P.total_rows < (long)rb * 148 * 2)) rb >>= 1;
    if (smem_for(rb) > 220 * 1024) throw Error("row axis too long for the shared-memory FFT");
    P.RB = rb; P.RBP = rb;
    const size_t smem = smem_for(rb);
    const long items = (long)S0 * rb;
    int nth = items >= 256 ? 256 : (items >= 128 ? 128 : (items >= 64 ? 64 : 32));
    dim3 grid((unsigned)((P.total_rows + rb - 1) / rb));
    PROF_BEGIN(pl, inverse ? 2 : 0, st);
    if (inverse) {
        auto k = rows_inv_fast_kernel<T, R0, Rs...>;
        if (smem > 48 * 1024) HIPGP_SET_MAX_SMEM(k, smem);
        HIPGP_LAUNCH(k, grid, dim3(nth), smem, st, P);
    } else {
        auto k = rows_fwd_fast_kernel<T, R0, Rs...>;
        if (smem > 48 * 1024) HIPGP_SET_MAX_SMEM(k, smem);
        HIPGP_LAUNCH(k, grid, dim3(nth), smem, st, P);
    }
    PROF_END(pl, st);
    CK_LAUNCH();
    pl->launches++;
}

#ifndef HIPGP_TB_SHIFT
#define HIPGP_TB_SHIFT 0
#endif
template <class T> constexpr int cols_tb_base(int L);
template <class T> constexpr int cols_tb(int L) { return cols_tb_base<T>(L) >> HIPGP_TB_SHIFT > 0 ? cols_tb_base<T>(L) >> HIPGP_TB_SHIFT : 1; }
template <class T> constexpr int cols_tb_base(int L) {
    // lines per CTA: keep the tile <= 128 KB and, where it fits, >= 64 B of contiguous lines per position
    return sizeof(T) == 4 ? (L <= 256 ? 32 : (L <= 512 ? 16 : (L <= 2048 ? 8 : (L <= 4096 ? 4 : 2))))
                          : (L <= 256 ? 16 : (L <= 512 ? 8 : (L <= 2048 ? 4 : (L <= 4096 ? 2 : 1))));
}

template <class T, int LEN, int R0, int... Rs>
static void launch_cols_fast_t(hipgp_plan* pl, ColsParams<T>& P, long n_outer, long B, cudaStream_t st) {
    constexpr int TB = cols_tb<T>(LEN);
    static_assert(RLInfo<RL<R0, Rs...>>::N == LEN, "radix list does not multiply to the length");
    P.TB = TB; P.TBP = TB;
    const size_t smem = sizeof(cplx<T>) * (size_t)line_stride<T>(LEN) * TB;
    const long items = (long)LEN * TB / 16;
    int nth = items >= 512 ? 512 : (items >= 256 ? 256 : (items >= 128 ? 128 : (items >= 64 ? 64 : 32)));
    static const char* env_nth = getenv("HIPGP_COLS_NTH");
    if (env_nth) nth = atoi(env_nth);
    dim3 grid((unsigned)((P.inner + TB - 1) / TB), (unsigned)n_outer, (unsigned)B);
    auto k = cols_fast_kernel<T, TB, R0, Rs...>;
    if (smem > 48 * 1024) HIPGP_SET_MAX_SMEM(k, smem);
    PROF_BEGIN(pl, 1, st);
    HIPGP_LAUNCH(k, grid, dim3(nth), smem, st, P);
    PROF_END(pl, st);
    CK_LAUNCH();
    pl->launches++;
}


template <class T, int LEN> struct FastList;
#define X(LEN, ...)                                                                                         \
    template <class T> struct FastList<T, LEN> {                                                            \
        static void rows(hipgp_plan* pl, bool inv, RowsParams<T>& P, cudaStream_t st) { launch_rows_fast_t<T, __VA_ARGS__>(pl, inv, P, st); } \
        static void cols(hipgp_plan* pl, ColsParams<T>& P, long no, long B, cudaStream_t st) { launch_cols_fast_t<T, LEN, __VA_ARGS__>(pl, P, no, B, st); } \
    };
HIPGP_FAST_LIST(X)
#undef X
template <class T, int LEN> void launch_rows_fast_len(hipgp_plan* pl, bool inverse, RowsParams<T>& P, cudaStream_t st) { FastList<T, LEN>::rows(pl, inverse, P, st); }
template <class T, int LEN> void launch_cols_fast_len(hipgp_plan* pl, ColsParams<T>& P, long n_outer, long B, cudaStream_t st) { FastList<T, LEN>::cols(pl, P, n_outer, B, st); }
}  // namespace hipgp
