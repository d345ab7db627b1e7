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

// S (+)= sum_b conj(V_b) U_b, the spectrum of the (non-symmetrised) linear cross-correlation l[lag] = sum_t v[t] u[t + lag]
template <class T>
__global__ void corr_accumulate_cplx_kernel(const T* __restrict__ U, const T* __restrict__ V, T* __restrict__ S, long ngroups,
                                            long stride, int nvec, int lane_layout, int first) {
    for (long gi = (long)blockIdx.x * blockDim.x + threadIdx.x; gi < ngroups; gi += (long)gridDim.x * blockDim.x) {
        double r0 = 0.0, i0 = 0.0, r1 = 0.0, i1 = 0.0;
        const int ia0 = 0, ib0 = lane_layout ? 2 : 1, ia1 = lane_layout ? 1 : 2, ib1 = 3;      // (re, im) slots of the two bins of a group
        for (int b = 0; b < nvec; ++b) {
            const T* u = U + (size_t)b * stride + 4 * gi;
            const T* v = V + (size_t)b * stride + 4 * gi;
            const double ur0 = u[ia0], ui0 = u[ib0], ur1 = u[ia1], ui1 = u[ib1], vr0 = v[ia0], vi0 = v[ib0], vr1 = v[ia1], vi1 = v[ib1];
            r0 += vr0 * ur0 + vi0 * ui0; i0 += vr0 * ui0 - vi0 * ur0;
            r1 += vr1 * ur1 + vi1 * ui1; i1 += vr1 * ui1 - vi1 * ur1;
        }
        T* s = S + 4 * gi;
        if (first) { s[ia0] = (T)r0; s[ib0] = (T)i0; s[ia1] = (T)r1; s[ib1] = (T)i1; }
        else { s[ia0] = (T)((double)s[ia0] + r0); s[ib0] = (T)((double)s[ib0] + i0); s[ia1] = (T)((double)s[ia1] + r1); s[ib1] = (T)((double)s[ib1] + i1); }
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
static void corr_forward(hipgp_plan* pl, Geom<T>& g, const T* in, long nb, cplx<T>* dst, long spec_elems, cudaStream_t s,
                         const int* n_in = nullptr) {
    const int D = g.D; const long P = g.P; const bool fast = geom_allows_fast(g);
    int nin[3];
    for (int d = 0; d < 3; ++d) nin[d] = n_in ? n_in[d] : pl->m[d];      // extents of the real input (default: the grid)
    long rows_in = 1;
    for (int d = 0; d + 1 < D; ++d) rows_in *= nin[d];
    cplx<T>* W1 = D == 1 ? dst : pl->W1.as<cplx<T>>();
    RowsParams<T> R{};
    rows_geom(R, g);
    R.in = in; R.W = W1; R.W_rows = (int)rows_in; R.mode = RF_PLAIN; R.do_fft = 1; R.total_rows = nb * rows_in;
    R.nrows = (int)rows_in; R.n_real = nin[D - 1]; R.st = null_state(); R.spec = nullptr; R.spec_kind = SPEC_NONE;
    launch_rows<T>(pl, false, R, s, fast);
    if (D == 2) {
        ColsParams<T> C{};
        C.in = W1; C.out = dst; C.n_in = nin[0]; C.n_out = nin[0]; C.inner = g.H + 1; C.pitch = P;
        C.in_bstride = (long)nin[0] * P; C.out_bstride = spec_elems;
        C.f = g.fcol[0].dev; C.mode = CM_FWD; C.spec = nullptr; C.spec_kind = SPEC_NONE;
        launch_cols<T>(pl, C, 1, nb, s, fast);
    } else if (D == 3) {
        const long L1 = g.L[1];
        ColsParams<T> C{};
        C.spec = nullptr; C.spec_kind = SPEC_NONE; C.mode = CM_FWD;
        C.in = W1; C.out = dst; C.n_in = nin[1]; C.n_out = nin[1]; C.inner = g.H + 1; C.pitch = P;
        C.in_ostride = (long)nin[1] * P; C.in_bstride = rows_in * P; C.out_ostride = L1 * P; C.out_bstride = spec_elems;
        C.f = g.fcol[1].dev;
        launch_cols<T>(pl, C, nin[0], nb, s, fast);
        C.in = dst; C.out = dst; C.n_in = nin[0]; C.n_out = nin[0]; C.inner = L1 * P; C.pitch = L1 * P;
        C.in_ostride = C.out_ostride = 0; C.in_bstride = C.out_bstride = spec_elems;
        C.f = g.fcol[0].dev;
        launch_cols<T>(pl, C, 1, nb, s, fast);
    }
}

// inverse transform of ONE spectrum, in place, to real lags [L_0 (x L_1)][m_last]
template <class T>
static void corr_inverse(hipgp_plan* pl, Geom<T>& g, cplx<T>* S, T* lag, cudaStream_t s, int n_last = 0) {
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
    R.n_real = n_last > 0 ? n_last : pl->m[D - 1]; R.st = null_state(); R.spec = nullptr; R.spec_kind = SPEC_NONE;
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


// ---------------------------------------------------------------------------------------------------------------------
// Gradient of sum_b g_b . (R^T v_b) with respect to the Toeplitz column (learn_kernel = True: the reference differentiates
// toeplitz_tensor.py:85-97 through D_sqrt = sqrt(max(Re FFT_N C, 1e-6)) with autograd).  With s = D^(1/2) on the (m_1..m_D)
// grid of distinct eigenvalues, w the DCT-I weights (1 at the ends of an axis, 2 inside) and Ntot = prod N_d:
//     qs    = the circular cross-correlation (period N_d per axis) of pad(v_b) and g_b, summed over b, averaged over the
//             reflections tau_d -> N_d - tau_d of every axis, restricted to tau in the grid
//     A     = DCT-I(qs)                           ( = sum over the reflections of Re conj(F pad v) F g )
//     X     = [D > clamp] A / (2 Ntot s)          ( dL/dD for the distinct eigenvalues, divided by their multiplicity w )
//     dL/dc = w . DCT-I(X)
// (verified against autograd through the reference formula to 1e-15, scripts/dev/proto_rt_grad.py).  qs needs no length-N
// transform: with l[lag] = sum_t v[t] g[t + lag] the LINEAR cross-correlation (lags -(m-1) .. N-1 per axis), the period-N
// correlation at tau is l[tau] + l[tau - N], and l[tau - N] is identically zero on the grid (tau - N < -(m-1)); l is circular
// on the plan's wide embedding (L'_d >= N_d + m_d - 1), so: forward passes of v and g, S = sum conj(V) G, ONE inverse (all
// L' lags of the last axis), a gather of at most 3^D lags {tau_d, -tau_d, N_d - tau_d}.  (The symmetrised accumulation of the
// quadratic form above would alias here: an even sequence of lags up to N - 1 needs L' >= 2N - 1.)
struct RtGradDims { int D; int m[3]; int N[3]; int L[3]; };

template <class T>
__global__ void rt_grad_gather_kernel(const T* __restrict__ lag, double* __restrict__ qs, RtGradDims q, long M, double scale) {
    const int D = q.D;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long)gridDim.x * blockDim.x) {
        int id[3] = {0, 0, 0};
        long r = i;
        for (int d = D - 1; d >= 0; --d) { id[d] = (int)(r % q.m[d]); r /= q.m[d]; }
        int cand[3][3], nc[3] = {1, 1, 1};
        double om = 1.0;
        for (int d = 0; d < D; ++d) {
            const int t = id[d], N = q.N[d];
            if (t == 0) { cand[d][0] = 0; nc[d] = 1; }
            else if (2 * t == N) { cand[d][0] = t; cand[d][1] = -t; nc[d] = 2; }
            else { cand[d][0] = t; cand[d][1] = -t; cand[d][2] = N - t; nc[d] = 3; om *= 0.5; }
        }
        double acc = 0.0;
        const int total = nc[0] * nc[1] * nc[2];
        for (int e = 0; e < total; ++e) {
            int ee = e;
            long idx = 0;
            for (int d = 0; d < D; ++d) {
                const int j = cand[d][ee % nc[d]]; ee /= nc[d];
                idx = idx * q.L[d] + (j < 0 ? j + q.L[d] : j);
            }
            acc += (double)lag[idx];
        }
        qs[i] = acc * om * scale;
    }
}
// X = [D > clamp] A / (2 Ntot sqrt(D))
__global__ void rt_grad_scale_kernel(double* __restrict__ A, const double* __restrict__ Dm, const double* __restrict__ Dsqrt, long M,
                                     double clampv, double inv2ntot) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) A[i] = Dm[i] > clampv ? A[i] * inv2ntot / Dsqrt[i] : 0.0;
}
// out = user_scale * w . H
template <class T>
__global__ void rt_grad_weight_kernel(const double* __restrict__ H, T* __restrict__ out, RtGradDims q, long M, double scale) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    long r = i; double w = 1.0;
    for (int d = q.D - 1; d >= 0; --d) { const int t = (int)(r % q.m[d]); r /= q.m[d]; w *= (t == 0 || t == q.m[d] - 1) ? 1.0 : 2.0; }
    out[i] = (T)(H[i] * w * scale);
}

template <class T>
static void rt_column_grad(hipgp_plan* pl, const void* vec, const void* gout, long B, double user_scale, void* out, cudaStream_t s) {
    if (!pl->have_spec) throw Error("plan has no spectrum: call hipgp_plan_set_first_row first");
    if (B < 0) throw Error("negative number of vectors");
    if (!out || (B > 0 && (!vec || !gout))) throw Error("null vector pointer");
    const long M = pl->M, E = pl->E;
    if (B == 0) { CK(cudaMemsetAsync(out, 0, sizeof(T) * (size_t)M, s)); return; }
    ensure_wide<T>(pl, s);
    Geom<T>& g = geom(pl, true, Tag<T>());
    const int D = g.D; const long P = g.P;
    long lrows = 1, rows_v = 1, rows_g = 1;
    for (int d = 0; d + 1 < D; ++d) { lrows *= g.L[d]; rows_v *= pl->m[d]; rows_g *= pl->N[d]; }
    const int nlast = g.L[D - 1];                        // all lags of the last axis (negative ones sit at L' + lag)
    const long spec_elems = lrows * P;
    const size_t spec_bytes = sizeof(cplx<T>) * (size_t)spec_elems;
    const long chunk = std::max<long>(1, std::min<long>(std::min<long>(16, B), (long)(((size_t)1 << 31) / spec_bytes)));
    pl->corrU.ensure(spec_bytes * chunk, &pl->dev_bytes); pl->corrV.ensure(spec_bytes * chunk, &pl->dev_bytes);
    pl->corrS.ensure(spec_bytes, &pl->dev_bytes);
    pl->corrLag.ensure(sizeof(T) * (size_t)lrows * nlast, &pl->dev_bytes);
    pl->gradA.ensure(sizeof(double) * (size_t)M, &pl->dev_bytes);
    if (D > 1) pl->W1.ensure(sizeof(cplx<T>) * (size_t)(chunk * std::max(rows_v, rows_g) * P), &pl->dev_bytes);
    const int lane_layout = (sizeof(T) == 4 && geom_allows_fast(g)) ? 1 : 0;
    const long ngroups = spec_elems / 2;
    const unsigned nblk = (unsigned)std::min<long>((ngroups + 255) / 256, 148L * 16);
    for (long c0 = 0; c0 < B; c0 += chunk) {
        const long nb = std::min(chunk, B - c0);
        corr_forward<T>(pl, g, (const T*)vec + (size_t)c0 * M, nb, pl->corrU.as<cplx<T>>(), spec_elems, s, pl->m);
        corr_forward<T>(pl, g, (const T*)gout + (size_t)c0 * E, nb, pl->corrV.as<cplx<T>>(), spec_elems, s, pl->N);
        auto k = corr_accumulate_cplx_kernel<T>;          // S += conj(FFT pad v) FFT pad g  (U = spectrum of g, V = spectrum of v)
        HIPGP_LAUNCH(k, dim3(nblk), dim3(256), 0, s, pl->corrV.as<T>(), pl->corrU.as<T>(), pl->corrS.as<T>(), ngroups,
                     2 * spec_elems, (int)nb, lane_layout, c0 == 0 ? 1 : 0);
        CK_LAUNCH(); pl->launches++;
    }
    corr_inverse<T>(pl, g, pl->corrS.as<cplx<T>>(), pl->corrLag.as<T>(), s, nlast);
    RtGradDims q{};
    q.D = D;
    double norm = 0.25, ntot = 1.0;
    for (int d = 0; d < 3; ++d) { q.m[d] = 1; q.N[d] = 1; q.L[d] = 1; }
    for (int d = 0; d < D; ++d) { q.m[d] = pl->m[d]; q.N[d] = pl->N[d]; q.L[d] = g.L[d]; norm /= (double)g.L[d]; ntot *= (double)pl->N[d]; }
    const unsigned nbm = (unsigned)std::min<long>((M + 255) / 256, 148L * 16);
    const unsigned nbe = (unsigned)((M + 255) / 256);
    {
        auto k = rt_grad_gather_kernel<T>;
        HIPGP_LAUNCH(k, dim3(nbm), dim3(256), 0, s, pl->corrLag.as<T>(), pl->tmpA.as<double>(), q, M, norm);
        CK_LAUNCH(); pl->launches++;
    }
    dct_all_axes(pl, pl->tmpA.as<double>(), pl->gradA.as<double>(), pl->tmpB.as<double>(), false, s);
    {
        auto k = rt_grad_scale_kernel;
        HIPGP_LAUNCH(k, dim3(nbe), dim3(256), 0, s, pl->gradA.as<double>(), pl->Dm.as<double>(), pl->Dsqrt.as<double>(), M, pl->clampv, 0.5 / ntot);
        CK_LAUNCH(); pl->launches++;
    }
    dct_all_axes(pl, pl->gradA.as<double>(), pl->tmpA.as<double>(), pl->tmpB.as<double>(), false, s);
    {
        auto k = rt_grad_weight_kernel<T>;
        HIPGP_LAUNCH(k, dim3(nbe), dim3(256), 0, s, pl->tmpA.as<double>(), (T*)out, q, M, user_scale);
        CK_LAUNCH(); pl->launches++;
    }
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
int hipgp_rt_column_grad(hipgp_plan* pl, const void* vec, const void* grad_out, int64_t B, double scale, void* out, void* stream) {
    API_BEGIN
    set_device(pl);
    DISPATCH(pl, rt_column_grad<float>(pl, vec, grad_out, (long)B, scale, out, (cudaStream_t)stream),
             rt_column_grad<double>(pl, vec, grad_out, (long)B, scale, out, (cudaStream_t)stream));
    API_END
}
}
