// kxu_api.inl -- C ABI for the on-the-fly cross-covariance kernels.
namespace hipgp {
template <class T>
static void kxu_launch(const KxuParams& P, const void* x, const void* grids, const void* ypts, const void* alphas, void* out,
                       cudaStream_t s) {
    if (P.B <= 0 || P.M <= 0) return;
    auto k = kxu_kernel<T>;
    // grid.y is limited to 65535: walk the observation axis in slabs
    for (long b0 = 0; b0 < P.B; b0 += 65535) {
        const long nb = std::min<long>(65535, P.B - b0);
        dim3 grid((unsigned)((P.M + 1023) / 1024), (unsigned)nb);
        HIPGP_LAUNCH(k, grid, dim3(256), 0, s, P, (const T*)x + b0 * P.ndim, (const T*)grids, (const T*)ypts, (const T*)alphas,
                     (T*)out + b0 * P.M);
        CK_LAUNCH();
    }
}
static void kxu_check(int dtype, int kernel_id, int mode, int ndim, int n_ell, const void* mc_alphas, int npts) {
    if (ndim < 1 || ndim > 3) throw Error("hipgp_kxu: ndim must be 1..3");
    if (kernel_id < 0 || kernel_id > 4) throw Error("hipgp_kxu: unknown kernel id");
    if (n_ell != 1 && n_ell != ndim) throw Error("hipgp_kxu: ell must have 1 or ndim entries");
    if (mode < 0 || mode > 4) throw Error("hipgp_kxu: unknown mode");
    if (mode == HIPGP_KXU_SEMI_ANALYTIC && kernel_id != HIPGP_K_SQEXP) throw Error("hipgp_kxu: analytic line integral exists for SqExp only");
    if ((mode == HIPGP_KXU_DERIV || mode == 4) && ndim != 1) throw Error("hipgp_kxu: derivative kernels are 1-D");
    if (mode == HIPGP_KXU_SEMI_MC && (npts < 1 || !mc_alphas)) throw Error("hipgp_kxu: SEMI_MC needs npts >= 1 and mc_alphas");
    if ((kernel_id >= 1 && kernel_id <= 3) && n_ell != 1) throw Error("hipgp_kxu: Matern takes a scalar ell");
    if (dtype != HIPGP_F32 && dtype != HIPGP_F64) throw Error("bad dtype");
}
}  // namespace hipgp

extern "C" {
int hipgp_kxu(int dtype, int kernel_id, int mode, double sig2, const double* ell, int n_ell, double gneiting_alpha,
              const void* x, int64_t B, int ndim, const int64_t* m, const void* grids, const void* mc_alphas, int npts,
              void* out, void* stream) {
    API_BEGIN
    kxu_check(dtype, kernel_id, mode, ndim, n_ell, mc_alphas, npts);
    KxuParams P{};
    P.kernel_id = kernel_id; P.mode = mode; P.ndim = ndim; P.npts = npts; P.sig2 = sig2; P.alpha = gneiting_alpha;
    P.B = (long)B; P.M = 1;
    int off = 0;
    for (int d = 0; d < 3; ++d) { P.m[d] = 1; P.goff[d] = 0; P.ell[d] = 1.0; }
    for (int d = 0; d < ndim; ++d) {
        P.m[d] = (int)m[d]; P.goff[d] = off; off += (int)m[d]; P.M *= m[d];
        P.ell[d] = n_ell == 1 ? ell[0] : ell[d];
    }
    P.ell0 = ell[0];
    if (dtype == HIPGP_F32) kxu_launch<float>(P, x, grids, nullptr, mc_alphas, out, (cudaStream_t)stream);
    else kxu_launch<double>(P, x, grids, nullptr, mc_alphas, out, (cudaStream_t)stream);
    API_END
}

int hipgp_kernel_pairwise(int dtype, int kernel_id, int mode, double sig2, const double* ell, int n_ell, double gneiting_alpha,
                          const void* x, int64_t n, const void* y, int64_t m, int ndim, const void* mc_alphas, int npts,
                          void* out, void* stream) {
    API_BEGIN
    kxu_check(dtype, kernel_id, mode, ndim, n_ell, mc_alphas, npts);
    KxuParams P{};
    P.kernel_id = kernel_id; P.mode = mode; P.ndim = ndim; P.npts = npts; P.sig2 = sig2; P.alpha = gneiting_alpha;
    P.B = (long)n; P.M = (long)m;
    for (int d = 0; d < 3; ++d) { P.m[d] = 1; P.goff[d] = 0; P.ell[d] = 1.0; }
    for (int d = 0; d < ndim; ++d) P.ell[d] = n_ell == 1 ? ell[0] : ell[d];
    P.ell0 = ell[0];
    if (dtype == HIPGP_F32) kxu_launch<float>(P, x, nullptr, y, mc_alphas, out, (cudaStream_t)stream);
    else kxu_launch<double>(P, x, nullptr, y, mc_alphas, out, (cudaStream_t)stream);
    API_END
}

/* d/d(sig2, ell) of sum_{b,j} G[b][j] k(x_b, u_j): partial (B * ceil(M/1024) * 4) doubles = per-block sums of
 * G * {dk/dsig2, dk/dell_0, dk/dell_1, dk/dell_2}; the caller adds them up (scalar ell: the three ell entries add). */
int hipgp_kxu_param_grad(int dtype, int kernel_id, int mode, double sig2, const double* ell, int n_ell, const void* x, int64_t B,
                         int ndim, const int64_t* m, const void* grids, const void* ypts, int64_t n_y, const void* mc_alphas, int npts,
                         const void* G, double* partial, void* stream) {
    API_BEGIN
    kxu_check(dtype, kernel_id, mode, ndim, n_ell, mc_alphas, npts);
    if (kernel_id == HIPGP_K_GNEITING) throw Error("hipgp_kxu_param_grad: no closed-form hyper-parameter derivative for the Gneiting kernel");
    if (mode != HIPGP_KXU_POINT && mode != HIPGP_KXU_SEMI_MC) throw Error("hipgp_kxu_param_grad: modes POINT and SEMI_MC only");
    if (!grids && !ypts) throw Error("hipgp_kxu_param_grad: give the 1-D grids or an explicit point set");
    KxuParams P{};
    P.kernel_id = kernel_id; P.mode = mode; P.ndim = ndim; P.npts = npts; P.sig2 = sig2; P.alpha = 1.0;
    P.B = (long)B; P.M = 1;
    int off = 0;
    for (int d = 0; d < 3; ++d) { P.m[d] = 1; P.goff[d] = 0; P.ell[d] = 1.0; }
    for (int d = 0; d < ndim; ++d) {
        if (grids) { P.m[d] = (int)m[d]; P.goff[d] = off; off += (int)m[d]; P.M *= m[d]; }
        P.ell[d] = n_ell == 1 ? ell[0] : ell[d];
    }
    if (!grids) P.M = (long)n_y;
    P.ell0 = ell[0];
    if (P.B <= 0 || P.M <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned nbx = (unsigned)((P.M + 1023) / 1024);
    for (long b0 = 0; b0 < P.B; b0 += 65535) {
        const long nb = std::min<long>(65535, P.B - b0);
        dim3 grid(nbx, (unsigned)nb);
        if (dtype == HIPGP_F32) {
            auto k = kxu_grad_kernel<float>;
            HIPGP_LAUNCH(k, grid, dim3(256), 0, s, P, (const float*)x + b0 * ndim, (const float*)grids, (const float*)ypts, (const float*)mc_alphas,
                         (const float*)G + b0 * P.M, partial + (size_t)b0 * nbx * 4);
        } else {
            auto k = kxu_grad_kernel<double>;
            HIPGP_LAUNCH(k, grid, dim3(256), 0, s, P, (const double*)x + b0 * ndim, (const double*)grids, (const double*)ypts, (const double*)mc_alphas,
                         (const double*)G + b0 * P.M, partial + (size_t)b0 * nbx * 4);
        }
        CK_LAUNCH();
    }
    API_END
}

int hipgp_doubly_diag(int dtype, const void* x, int64_t B, int ndim, double sig2, const double* ell, int n_ell,
                      const void* dgrid, const void* slopes, const void* knn, int ntab, void* out, void* stream) {
    API_BEGIN
    if (ndim < 1 || ndim > 3) throw Error("hipgp_doubly_diag: ndim must be 1..3");
    if (B <= 0) return 0;
    double e[3] = {1, 1, 1};
    for (int d = 0; d < ndim; ++d) e[d] = n_ell == 1 ? ell[0] : ell[d];
    const unsigned nb = (unsigned)((B + 127) / 128);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == HIPGP_F32) {
        auto k = doubly_diag_kernel<float>;
        HIPGP_LAUNCH(k, dim3(nb), dim3(128), 0, s, (const float*)x, (long)B, ndim, sig2, ell[0], (const float*)dgrid,
                     (const float*)slopes, (const float*)knn, ntab, (float*)out, e[0], e[1], e[2]);
    } else {
        auto k = doubly_diag_kernel<double>;
        HIPGP_LAUNCH(k, dim3(nb), dim3(128), 0, s, (const double*)x, (long)B, ndim, sig2, ell[0], (const double*)dgrid,
                     (const double*)slopes, (const double*)knn, ntab, (double*)out, e[0], e[1], e[2]);
    }
    CK_LAUNCH();
    API_END
}
}
