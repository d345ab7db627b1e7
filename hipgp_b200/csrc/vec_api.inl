// vec_api.inl -- C ABI for the stand-alone fused CG vector kernels and their scratch space.
#include <map>
#include <mutex>
#include <utility>
namespace hipgp {
// Partial-sum scratch, one buffer per (device, stream): the kernel and its reduction are ordered on the caller's stream,
// so two streams (or two devices of one process) must never share a buffer.
static std::mutex g_scratch_mu;
static std::map<std::pair<int, cudaStream_t>, DevBuf> g_scratch;
static size_t g_scratch_total = 0;
static const int kChunks = 64;
static const long kFillBlocks = 148 * 8;
static const long kMaxGridY = 65535;

static double* scratch_for(cudaStream_t s, size_t bytes) {
    int dev = 0;
    CK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    DevBuf& b = g_scratch[std::make_pair(dev, s)];
    if (bytes > b.bytes) {
        // an older, smaller buffer may still be read by work queued on this stream
        if (b.p) CK(cudaStreamSynchronize(s));
        b.ensure(bytes, &g_scratch_total);
    }
    return b.as<double>();
}

template <class T>
static void vec_op(int op, void* x, void* r, const void* a, const void* b2, const double* num, const double* den, double* out,
                   long B, long M, cudaStream_t s) {
    if (B <= 0 || M <= 0) return;
    // at least kChunks chunks per right-hand side; few long right-hand sides get enough chunks to fill the machine
    const long want = std::max<long>(kChunks, (kFillBlocks + B - 1) / B);
    int nchunk = (int)std::min<long>(want, std::max<long>(1, (M + 4095) / 4096));
    double* scr = scratch_for(s, sizeof(double) * (size_t)B * (size_t)nchunk);
    auto k = vec_kernel<T>;
    for (long b0 = 0; b0 < B; b0 += kMaxGridY) {        // grid.y is limited to 65535 right-hand sides per launch
        const long nb = std::min(kMaxGridY, B - b0);
        const size_t o = (size_t)b0 * (size_t)M;
        HIPGP_LAUNCH(k, dim3(nchunk, (unsigned)nb), dim3(256), 0, s, op, x ? (T*)x + o : nullptr, r ? (T*)r + o : nullptr,
                     a ? (const T*)a + o : nullptr, b2 ? (const T*)b2 + o : nullptr, num ? num + b0 : nullptr, den ? den + b0 : nullptr,
                     scr + (size_t)b0 * nchunk, M, nchunk);
        CK_LAUNCH();
    }
    if (op != 2) {
        auto k2 = vec_reduce_kernel;
        HIPGP_LAUNCH(k2, dim3((unsigned)((B + 127) / 128)), dim3(128), 0, s, scr, out, nchunk, B);
        CK_LAUNCH();
    }
}
template <class T>
static void mf_rowstats(const void* kn, const void* qm, const void* qS, long B, long E, void* out, cudaStream_t s) {
    const int nchunk = (int)std::min<long>(kChunks, std::max<long>(1, (E + 8191) / 8192));
    double* scr = scratch_for(s, sizeof(double) * (size_t)B * 3 * kChunks);
    auto k = mf_rowstats_kernel<T>;
    for (long b0 = 0; b0 < B; b0 += kMaxGridY) {
        const long nb = std::min(kMaxGridY, B - b0);
        HIPGP_LAUNCH(k, dim3(nchunk, (unsigned)nb), dim3(256), 0, s, (const T*)kn + (size_t)b0 * (size_t)E, (const T*)qm, (const T*)qS,
                     scr + (size_t)b0 * 3 * nchunk, E, nchunk);
        CK_LAUNCH();
    }
    auto k2 = mf_rowstats_reduce_kernel<T>;
    HIPGP_LAUNCH(k2, dim3((unsigned)((3 * B + 127) / 128)), dim3(128), 0, s, scr, (T*)out, nchunk, B);
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
    if (E <= 0) return 0;
    if (B <= 0) {   // an empty shard contributes zeros (the caller all-reduces these buffers)
        const size_t w = dtype == HIPGP_F32 ? 4 : 8;
        CK(cudaMemsetAsync(dm, 0, w * (size_t)E, (cudaStream_t)stream));
        CK(cudaMemsetAsync(lam, 0, w * (size_t)E, (cudaStream_t)stream));
        return 0;
    }
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
