// vec_api.inl -- C ABI for the stand-alone fused CG vector kernels and their scratch space.
#include <mutex>
namespace hipgp {
static std::mutex g_scratch_mu;
static DevBuf g_scratch;
static size_t g_scratch_total = 0;
static const int kChunks = 64;

template <class T>
static void vec_op(int op, void* x, void* r, const void* a, const void* b2, const double* num, const double* den, double* out,
                   long B, long M, cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    int nchunk = (int)std::min<long>(kChunks, std::max<long>(1, (M + 4095) / 4096));
    g_scratch.ensure(sizeof(double) * (size_t)B * kChunks, &g_scratch_total);
    auto k = vec_kernel<T>;
    HIPGP_LAUNCH(k, dim3(nchunk, (unsigned)B), dim3(256), 0, s, op, (T*)x, (T*)r, (const T*)a, (const T*)b2, num, den,
                 g_scratch.as<double>(), M, nchunk);
    CK_LAUNCH();
    if (op != 2) {
        auto k2 = vec_reduce_kernel;
        HIPGP_LAUNCH(k2, dim3((unsigned)((B + 127) / 128)), dim3(128), 0, s, g_scratch.as<double>(), out, nchunk, B);
        CK_LAUNCH();
    }
}
template <class T>
static void mf_rowstats(const void* kn, const void* qm, const void* qS, long B, long E, void* out, cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    const int nchunk = (int)std::min<long>(kChunks, std::max<long>(1, (E + 8191) / 8192));
    g_scratch.ensure(sizeof(double) * (size_t)B * 3 * kChunks, &g_scratch_total);
    auto k = mf_rowstats_kernel<T>;
    HIPGP_LAUNCH(k, dim3(nchunk, (unsigned)B), dim3(256), 0, s, (const T*)kn, (const T*)qm, (const T*)qS, g_scratch.as<double>(), E, nchunk);
    CK_LAUNCH();
    auto k2 = mf_rowstats_reduce_kernel<T>;
    HIPGP_LAUNCH(k2, dim3((unsigned)((3 * B + 127) / 128)), dim3(128), 0, s, g_scratch.as<double>(), (T*)out, nchunk, B);
    CK_LAUNCH();
}
template <class T>
static void mf_colstats(const void* kn, const void* w1, const void* w2, long B, long E, void* dm, void* lam, cudaStream_t s) {
    auto k = mf_colstats_kernel<T>;
    HIPGP_LAUNCH(k, dim3((unsigned)((E + 255) / 256)), dim3(256), 0, s, (const T*)kn, (const T*)w1, (const T*)w2, (T*)dm, (T*)lam, B, E);
    CK_LAUNCH();
}
}  // namespace hipgp

extern "C" {
int hipgp_meanfield_rowstats(int dtype, const void* kn, const void* qm, const void* qS, int64_t B, int64_t E, void* out, void* stream) {
    API_BEGIN
    if (B <= 0 || E <= 0) return 0;
    if (dtype == HIPGP_F32) mf_rowstats<float>(kn, qm, qS, (long)B, (long)E, out, (cudaStream_t)stream);
    else mf_rowstats<double>(kn, qm, qS, (long)B, (long)E, out, (cudaStream_t)stream);
    API_END
}
int hipgp_meanfield_colstats(int dtype, const void* kn, const void* w1, const void* w2, int64_t B, int64_t E, void* dm, void* lam,
                             void* stream) {
    API_BEGIN
    if (B <= 0 || E <= 0) return 0;
    if (dtype == HIPGP_F32) mf_colstats<float>(kn, w1, w2, (long)B, (long)E, dm, lam, (cudaStream_t)stream);
    else mf_colstats<double>(kn, w1, w2, (long)B, (long)E, dm, lam, (cudaStream_t)stream);
    API_END
}
int hipgp_vec_dot(int dtype, const void* a, const void* b, double* out, int64_t B, int64_t M, void* stream) {
    API_BEGIN
    if (dtype == HIPGP_F32) vec_op<float>(0, nullptr, nullptr, a, b, nullptr, nullptr, out, (long)B, (long)M, (cudaStream_t)stream);
    else vec_op<double>(0, nullptr, nullptr, a, b, nullptr, nullptr, out, (long)B, (long)M, (cudaStream_t)stream);
    API_END
}
int hipgp_vec_xr_update(int dtype, void* x, void* r, const void* p, const void* Ap, const double* rs, const double* pAp,
                        double* rr_out, int64_t B, int64_t M, void* stream) {
    API_BEGIN
    if (dtype == HIPGP_F32) vec_op<float>(1, x, r, p, Ap, rs, pAp, rr_out, (long)B, (long)M, (cudaStream_t)stream);
    else vec_op<double>(1, x, r, p, Ap, rs, pAp, rr_out, (long)B, (long)M, (cudaStream_t)stream);
    API_END
}
int hipgp_vec_p_update(int dtype, void* p, const void* z, const double* zr_new, const double* zr_old, int64_t B, int64_t M,
                       void* stream) {
    API_BEGIN
    if (dtype == HIPGP_F32) vec_op<float>(2, p, nullptr, z, nullptr, zr_new, zr_old, nullptr, (long)B, (long)M, (cudaStream_t)stream);
    else vec_op<double>(2, p, nullptr, z, nullptr, zr_new, zr_old, nullptr, (long)B, (long)M, (cudaStream_t)stream);
    API_END
}
}
