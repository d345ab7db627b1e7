// corr_api.inl -- the Toeplitz-column quadratic form behind InvMatmul.backward (learn_kernel=True).
//
// Reference: ziggy/misc/_inv_matmul.py:39-55 calls gpt_toeplitz.py:169-209 `sym_toeplitz_derivative_quadratic_form` on
// the FLATTENED M-vectors: for pairs (u_j, v_j)
//     out[i] = sum_j ( c_j[i] + c_j[-i] )  (i >= 1),   out[0] = sum_j u_j . v_j,     c_j[i] = sum_k u_j[k + i] v_j[k]
// i.e. a symmetrised LINEAR cross-correlation of length M (the reference runs two 1-D FFT Toeplitz products of size
// 2M - 1 per pair).  Here: a flattened lag i = (i_0, .., i_{D-1}) (row major) is the sum of at most 2^(D-1) lags of the
// D-dimensional linear correlation (one per carry pattern of k + i), and that correlation is a circular one on the
// plan's own narrow embedding (L_d >= 2 m_d - 1).  So: forward transforms of u_j and v_j with the row / column passes
// of the matvec, S = sum_j 2 Re(conj(V_j) U_j) (real: the symmetrised correlation is even), ONE inverse transform of S
// and a gather over carry patterns.  2S forward + 1 inverse transforms instead of the reference's 4S of length 2M - 1.
namespace hipgp {

// U, V: nvec spectra of `ngroups` groups of 4 reals each (stride reals apart); S += sum 2 Re(conj V U), Im S = 0.
// lane_layout: a group is (re0, re1, im0, im1) (fp32 specialised kernels) instead of (re0, im0, re1, im1).
template <class T>
__global__ void corr_accumulate_kernel(const T* __restrict__ U, const T* __restrict__ V, T* __restrict__ S, long ngroups,
                                       long stride, int nvec, int lane_layout, int first) {
    for (long gi = (long)blockIdx.x * blockDim.x + threadIdx.x; gi < ngroups; gi += (long)gridDim.x * blockDim.x) {
        double a0 = 0.0, a1 = 0.0;
        for (int b = 0; b < nvec; ++b) {
            const T* u = U + (size_t)b * stride + 4 * gi;
            const T* v = V + (size_t)b * stride + 4 * gi;
            const double u0 = u[0], u1 = u[1], u2 = u[2], u3 = u[3], v0 = v[0], v1 = v[1], v2 = v[2], v3 = v[3];
            if (lane_layout) { a0 += u0 * v0 + u2 * v2; a1 += u1 * v1 + u3 * v3; }
            else { a0 += u0 * v0 + u1 * v1; a1 += u2 * v2 + u3 * v3; }
        }
        T* s = S + 4 * gi;
        const int i1 = lane_layout ? 1 : 2;
        const double p0 = first ? 0.0 : (double)s[0], p1 = first ? 0.0 : (double)s[i1];
        s[0] = (T)0; s[1] = (T)0; s[2] = (T)0; s[3] = (T)0;
        s[0] = (T)(p0 + 2.0 * a0); s[i1] = (T)(p1 + 2.0 * a1);
    }
}

struct CorrDims { int D; int m[3]; int L[3]; };

// lag: the even D-dimensional correlation at lags (j_0 mod L_0, .., j_{D-2} mod L_{D-2}, j_{D-1} in [0, m_{D-1}));
// out[i] = scale * (i == 0 ? 1/2 : 1) * sum over carry patterns of lag[j(i, e)]
template <class T>
__global__ void corr_combine_kernel(const T* __restrict__ lag, T* __restrict__ out, CorrDims q, long M, double scale) {
    const int D = q.D;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long)gridDim.x * blockDim.x) {
        int id[3] = {0, 0, 0};
        long r = i;
        for (int d = D - 1; d >= 0; --d) { id[d] = (int)(r % q.m[d]); r /= q.m[d]; }
        double acc = 0.0;
        for (int e = 0; e < (1 << (D - 1)); ++e) {
            // bit (d - 1) of e = carry out of axis d into axis d - 1
            int j[3] = {0, 0, 0};
            bool ok = true;
            for (int d = 0; d < D; ++d) {
                const int cin = (d + 1 < D) ? ((e >> d) & 1) : 0;          // carry coming in from axis d + 1
                const int cout = (d > 0) ? ((e >> (d - 1)) & 1) : 0;       // carry going out to axis d - 1
                j[d] = id[d] + cin - q.m[d] * cout;
                if (j[d] >= q.m[d] || j[d] <= -q.m[d]) ok = false;
            }
            if (!ok) continue;
            if (j[D - 1] < 0) for (int d = 0; d < D; ++d) j[d] = -j[d];
            long idx = 0;
            for (int d = 0; d + 1 < D; ++d) idx = idx * q.L[d] + (j[d] < 0 ? j[d] + q.L[d] : j[d]);
            idx = idx * q.m[D - 1] + j[D - 1];
            acc += (double)lag[idx];
        }
        out[i] = (T)(acc * scale * (i == 0 ? 0.5 : 1.0));
    }
}

// forward transform of nb vectors (no spectrum, no inverse): dst[b] = the plan's half-spectrum layout, all L rows
template <class T>
static void corr_forward(hipgp_plan* pl, Geom<T>& g, const T* in, long nb, cplx<T>* dst, long spec_elems, cudaStream_t s) {
    const int D = g.D; const long P = g.P; const bool fast = geom_allows_fast(g);
    long rows_in = 1;
    for (int d = 0; d + 1 < D; ++d) rows_in *= pl->m[d];
    cplx<T>* W1 = D == 1 ? dst : pl->W1.as<cplx<T>>();
    RowsParams<T> R{};
    rows_geom(R, g);
    R.in = in; R.W = W1; R.W_rows = (int)rows_in; R.mode = RF_PLAIN; R.do_fft = 1; R.total_rows = nb * rows_in;
    R.nrows = (int)rows_in; R.n_real = pl->m[D - 1]; R.st = null_state(); R.spec = nullptr; R.spec_kind = SPEC_NONE;
    launch_rows<T>(pl, false, R, s, fast);
    if (D == 2) {
        ColsParams<T> C{};
        C.in = W1; C.out = dst; C.n_in = pl->m[0]; C.n_out = pl->m[0]; C.inner = g.H + 1; C.pitch = P;
        C.in_bstride = (long)pl->m[0] * P; C.out_bstride = spec_elems;
        C.f = g.fcol[0].dev; C.mode = CM_FWD; C.spec = nullptr; C.spec_kind = SPEC_NONE;
        launch_cols<T>(pl, C, 1, nb, s, fast);
    } else if (D == 3) {
        const long L1 = g.L[1];
        ColsParams<T> C{};
        C.spec = nullptr; C.spec_kind = SPEC_NONE; C.mode = CM_FWD;
        C.in = W1; C.out = dst; C.n_in = pl->m[1]; C.n_out = pl->m[1]; C.inner = g.H + 1; C.pitch = P;
        C.in_ostride = (long)pl->m[1] * P; C.in_bstride = rows_in * P; C.out_ostride = L1 * P; C.out_bstride = spec_elems;
        C.f = g.fcol[1].dev;
        launch_cols<T>(pl, C, pl->m[0], nb, s, fast);
        C.in = dst; C.out = dst; C.n_in = pl->m[0]; C.n_out = pl->m[0]; C.inner = L1 * P; C.pitch = L1 * P;
        C.in_ostride = C.out_ostride = 0; C.in_bstride = C.out_bstride = spec_elems;
        C.f = g.fcol[0].dev;
        launch_cols<T>(pl, C, 1, nb, s, fast);
    }
}

// inverse transform of ONE spectrum, in place, to real lags [L_0 (x L_1)][m_last]
template <class T>
static void corr_inverse(hipgp_plan* pl, Geom<T>& g, cplx<T>* S, T* lag, cudaStream_t s) {
    const int D = g.D; const long P = g.P; const bool fast = geom_allows_fast(g);
    long lrows = 1;
    for (int d = 0; d + 1 < D; ++d) lrows *= g.L[d];
    if (D == 2) {
        ColsParams<T> C{};
        C.in = S; C.out = S; C.n_in = g.L[0]; C.n_out = g.L[0]; C.inner = g.H + 1; C.pitch = P;
        C.f = g.fcol[0].dev; C.mode = CM_INV; C.spec = nullptr; C.spec_kind = SPEC_NONE;
        launch_cols<T>(pl, C, 1, 1, s, fast);
    } else if (D == 3) {
        const long L1 = g.L[1];
        ColsParams<T> C{};
        C.spec = nullptr; C.spec_kind = SPEC_NONE; C.mode = CM_INV;
        C.in = S; C.out = S; C.n_in = g.L[0]; C.n_out = g.L[0]; C.inner = L1 * P; C.pitch = L1 * P;
        C.f = g.fcol[0].dev;
        launch_cols<T>(pl, C, 1, 1, s, fast);
        C.n_in = (int)L1; C.n_out = (int)L1; C.inner = g.H + 1; C.pitch = P; C.in_ostride = C.out_ostride = L1 * P;
        C.f = g.fcol[1].dev;
        launch_cols<T>(pl, C, g.L[0], 1, s, fast);
    }
    RowsParams<T> R{};
    rows_geom(R, g);
    R.out = lag; R.W = S; R.W_rows = (int)lrows; R.mode = RI_PLAIN; R.do_fft = 1; R.total_rows = lrows; R.nrows = (int)lrows;
    R.n_real = pl->m[D - 1]; R.st = null_state(); R.spec = nullptr; R.spec_kind = SPEC_NONE;
    launch_rows<T>(pl, true, R, s, fast);
}

template <class T>
static void toeplitz_quadform(hipgp_plan* pl, const void* left, const void* right, long S, double user_scale, void* out,
                              cudaStream_t s) {
    if (!pl->have_spec) throw Error("plan has no spectrum: call hipgp_plan_set_first_row first");
    if (S < 0) throw Error("negative number of vector pairs");
    if (!out || (S > 0 && (!left || !right))) throw Error("null vector pointer");
    const long M = pl->M;
    if (S == 0) { CK(cudaMemsetAsync(out, 0, sizeof(T) * (size_t)M, s)); return; }
    Geom<T>& g = geom(pl, false, Tag<T>());
    const int D = g.D; const long P = g.P;
    long lrows = 1, rows_in = 1;
    for (int d = 0; d + 1 < D; ++d) { lrows *= g.L[d]; rows_in *= pl->m[d]; }
    const long spec_elems = lrows * P;
    const size_t spec_bytes = sizeof(cplx<T>) * (size_t)spec_elems;
    const long chunk = std::max<long>(1, std::min<long>(std::min<long>(16, S), (long)(((size_t)1 << 31) / spec_bytes)));
    pl->corrU.ensure(spec_bytes * chunk, &pl->dev_bytes); pl->corrV.ensure(spec_bytes * chunk, &pl->dev_bytes);
    pl->corrS.ensure(spec_bytes, &pl->dev_bytes);
    pl->corrLag.ensure(sizeof(T) * (size_t)lrows * pl->m[D - 1], &pl->dev_bytes);
    if (D > 1) pl->W1.ensure(sizeof(cplx<T>) * (size_t)(chunk * rows_in * P), &pl->dev_bytes);
    const int lane_layout = (sizeof(T) == 4 && geom_allows_fast(g)) ? 1 : 0;
    const long ngroups = spec_elems / 2;                 // 4 reals = 2 complex bins per group (P is a multiple of 8)
    const unsigned nblk = (unsigned)std::min<long>((ngroups + 255) / 256, 148L * 16);
    for (long c0 = 0; c0 < S; c0 += chunk) {
        const long nb = std::min(chunk, S - c0);
        corr_forward<T>(pl, g, (const T*)left + (size_t)c0 * M, nb, pl->corrU.as<cplx<T>>(), spec_elems, s);
        corr_forward<T>(pl, g, (const T*)right + (size_t)c0 * M, nb, pl->corrV.as<cplx<T>>(), spec_elems, s);
        auto k = corr_accumulate_kernel<T>;
        HIPGP_LAUNCH(k, dim3(nblk), dim3(256), 0, s, pl->corrU.as<T>(), pl->corrV.as<T>(), pl->corrS.as<T>(), ngroups,
                     2 * spec_elems, (int)nb, lane_layout, c0 == 0 ? 1 : 0);
        CK_LAUNCH(); pl->launches++;
    }
    corr_inverse<T>(pl, g, pl->corrS.as<cplx<T>>(), pl->corrLag.as<T>(), s);
    CorrDims q{};
    q.D = D;
    double norm = 0.25;                                   // forward passes are unscaled, the inverse carries 4 prod L (as the stored spectra assume)
    for (int d = 0; d < D; ++d) { q.m[d] = pl->m[d]; q.L[d] = g.L[d]; norm /= (double)g.L[d]; }
    auto k = corr_combine_kernel<T>;
    const unsigned nb2 = (unsigned)std::min<long>((M + 255) / 256, 148L * 16);
    HIPGP_LAUNCH(k, dim3(nb2), dim3(256), 0, s, pl->corrLag.as<T>(), (T*)out, q, M, norm * user_scale);
    CK_LAUNCH(); pl->launches++;
}

}  // namespace hipgp

extern "C" {
int hipgp_toeplitz_quadform(hipgp_plan* pl, const void* left, const void* right, int64_t S, double scale, void* out, void* stream) {
    API_BEGIN
    set_device(pl);
    DISPATCH(pl, toeplitz_quadform<float>(pl, left, right, (long)S, scale, out, (cudaStream_t)stream),
             toeplitz_quadform<double>(pl, left, right, (long)S, scale, out, (cudaStream_t)stream));
    API_END
}
}
