// block_api.inl -- block-diagonal variational family (ziggy/hipgp.py:527-690): the two contractions over k_n (B, M') that
// the reference writes as  to_blocks -> batched matmul -> from_blocks  (util.py:79-126 index maps), fused so that the
// permuted copies and the (B, num_blocks, bs, bs) outer-product tensor of hipgp.py:252-255 are never materialised.
//   blk_idx : (num_blocks, bs) int64, a permutation of [0, M') (define_block_chunks, util.py:79-117)
//   lam[k][i][j] = scale * sum_n w[n] kn[n, idx[k,i]] kn[n, idx[k,j]] + diag * (i == j)        (get_lam, hipgp.py:666-685)
//   out[b, idx[k,i]] = sum_j S[k][i][j] v[b, idx[k,j]]                                          (block_diag_multiply, :640-652)
namespace hipgp {

// one CTA per block; rows of k_n stream through shared memory in chunks of `rchunk`, partial sums live in `out`
template <class T>
__global__ void __launch_bounds__(256) block_lam_kernel(const T* __restrict__ kn, const T* __restrict__ w,
                                                        const long long* __restrict__ idx, long B, long E, int bs, int rchunk,
                                                        T scale, T diag, T* __restrict__ out) {
    HIPGP_DYN_SMEM(smem_raw);
    long long* sidx = reinterpret_cast<long long*>(smem_raw);
    T* A = reinterpret_cast<T*>(sidx + bs);              // [rchunk][bs]
    T* ws = A + (size_t)rchunk * bs;                      // [rchunk]
    const long k = blockIdx.x;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < bs; i += nt) sidx[i] = idx[k * bs + i];
    __syncthreads();
    T* o = out + (size_t)k * bs * bs;
    for (long r0 = 0; r0 < B; r0 += rchunk) {
        const int nr = (int)((B - r0) < rchunk ? (B - r0) : rchunk);
        for (int t = tid; t < nr * bs; t += nt) {
            const int n = t / bs, i = t - n * bs;
            A[(size_t)n * bs + i] = kn[(size_t)(r0 + n) * E + sidx[i]];
        }
        for (int n = tid; n < nr; n += nt) ws[n] = w[r0 + n];
        __syncthreads();
        const bool last = r0 + nr >= B;
        for (int t = tid; t < bs * bs; t += nt) {
            const int i = t / bs, j = t - i * bs;
            T acc = r0 == 0 ? (T)0 : o[t];
            for (int n = 0; n < nr; ++n) acc += (ws[n] * A[(size_t)n * bs + i]) * A[(size_t)n * bs + j];
            if (last) acc = scale * acc + (i == j ? diag : (T)0);
            o[t] = acc;
        }
        __syncthreads();
    }
}

// grid (num_blocks, ceil(B / 8)); a warp owns output row i of the block, its lanes stride over j (coalesced S reads)
template <class T>
__global__ void __launch_bounds__(256) block_diag_multiply_kernel(const T* __restrict__ S, const T* __restrict__ v,
                                                                  const long long* __restrict__ idx, long B, long E, int bs,
                                                                  T* __restrict__ out) {
    constexpr int BCH = 8;
    HIPGP_DYN_SMEM(smem_raw);
    long long* sidx = reinterpret_cast<long long*>(smem_raw);
    T* vs = reinterpret_cast<T*>(sidx + bs);             // [BCH][bs]
    const long k = blockIdx.x;
    const long b0 = (long)blockIdx.y * BCH;
    const int nb = (int)((B - b0) < BCH ? (B - b0) : BCH);
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    for (int i = tid; i < bs; i += nt) sidx[i] = idx[k * bs + i];
    __syncthreads();
    for (int t = tid; t < BCH * bs; t += nt) {
        const int b = t / bs, j = t - b * bs;
        vs[t] = b < nb ? v[(size_t)(b0 + b) * E + sidx[j]] : (T)0;
    }
    __syncthreads();
    const T* Sk = S + (size_t)k * bs * bs;
    for (int i = warp; i < bs; i += nwarp) {
        T acc[BCH];
#pragma unroll
        for (int b = 0; b < BCH; ++b) acc[b] = (T)0;
        for (int j = lane; j < bs; j += 32) {
            const T s = Sk[(size_t)i * bs + j];
#pragma unroll
            for (int b = 0; b < BCH; ++b) acc[b] += s * vs[b * bs + j];
        }
#pragma unroll
        for (int b = 0; b < BCH; ++b) {
            T a = acc[b];
            for (int m = 16; m >= 1; m >>= 1) a += __shfl_xor_sync(0xffffffffu, a, m);
            acc[b] = a;
        }
        if (lane == 0)
            for (int b = 0; b < nb; ++b) out[(size_t)(b0 + b) * E + sidx[i]] = acc[b];
    }
}

static void block_check(const void* idx, long B, long E, long nblk, long bs) {
    if (!idx) throw Error("null block index");
    if (B < 0) throw Error("negative batch");
    if (nblk < 1 || bs < 1 || nblk * bs != E) throw Error("block index must be (num_blocks, block_size) with num_blocks * block_size = M'");
    if (bs > 4096) throw Error("block size above 4096 is not supported");
}

template <class T>
static void block_lam(const void* kn, const void* w, const void* idx, long B, long E, long nblk, long bs, double scale,
                      double diag, void* out, cudaStream_t s) {
    block_check(idx, B, E, nblk, bs);
    if (!out || (B > 0 && (!kn || !w))) throw Error("null pointer");
    long rchunk = (long)(64 * 1024 / (sizeof(T) * (size_t)bs));
    rchunk = std::max<long>(1, std::min<long>(rchunk, std::max<long>(B, 1)));
    const size_t smem = sizeof(long long) * (size_t)bs + sizeof(T) * (size_t)(rchunk * bs + rchunk);
    auto k = block_lam_kernel<T>;
    if (smem > 40 * 1024) HIPGP_SET_MAX_SMEM(k, smem);
    if (B == 0) {                                          // empty minibatch: lam = diag * I
        std::vector<T> h((size_t)nblk * bs * bs, (T)0);
        for (long q = 0; q < nblk; ++q) for (long i = 0; i < bs; ++i) h[(size_t)q * bs * bs + i * bs + i] = (T)diag;
        CK(cudaMemcpyAsync(out, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
        return;
    }
    HIPGP_LAUNCH(k, dim3((unsigned)nblk), dim3(256), smem, s, (const T*)kn, (const T*)w, (const long long*)idx, B, E, (int)bs,
                 (int)rchunk, (T)scale, (T)diag, (T*)out);
    CK_LAUNCH();
}

template <class T>
static void block_diag_multiply(const void* S, const void* v, const void* idx, long B, long E, long nblk, long bs, void* out,
                                cudaStream_t s) {
    block_check(idx, B, E, nblk, bs);
    if (B == 0) return;
    if (!S || !v || !out) throw Error("null pointer");
    const size_t smem = sizeof(long long) * (size_t)bs + sizeof(T) * (size_t)(8 * bs);
    auto k = block_diag_multiply_kernel<T>;
    if (smem > 40 * 1024) HIPGP_SET_MAX_SMEM(k, smem);
    for (long b0 = 0; b0 < B; b0 += 8 * 65535L) {          // grid.y limit
        const long nb = std::min<long>(B - b0, 8 * 65535L);
        HIPGP_LAUNCH(k, dim3((unsigned)nblk, (unsigned)((nb + 7) / 8)), dim3(256), smem, s, (const T*)S,
                     (const T*)v + (size_t)b0 * E, (const long long*)idx, nb, E, (int)bs, (T*)out + (size_t)b0 * E);
        CK_LAUNCH();
    }
}

}  // namespace hipgp

extern "C" {
int hipgp_block_lam(int dtype, const void* kn, const void* w, const int64_t* blk_idx, int64_t B, int64_t E, int64_t nblk,
                    int64_t bs, double scale, double diag, void* out, void* stream) {
    API_BEGIN
    if (dtype == HIPGP_F32) block_lam<float>(kn, w, blk_idx, (long)B, (long)E, (long)nblk, (long)bs, scale, diag, out, (cudaStream_t)stream);
    else if (dtype == HIPGP_F64) block_lam<double>(kn, w, blk_idx, (long)B, (long)E, (long)nblk, (long)bs, scale, diag, out, (cudaStream_t)stream);
    else throw Error("dtype must be HIPGP_F32 or HIPGP_F64");
    API_END
}
int hipgp_block_diag_multiply(int dtype, const void* S, const void* v, const int64_t* blk_idx, int64_t B, int64_t E,
                              int64_t nblk, int64_t bs, void* out, void* stream) {
    API_BEGIN
    if (dtype == HIPGP_F32) block_diag_multiply<float>(S, v, blk_idx, (long)B, (long)E, (long)nblk, (long)bs, out, (cudaStream_t)stream);
    else if (dtype == HIPGP_F64) block_diag_multiply<double>(S, v, blk_idx, (long)B, (long)E, (long)nblk, (long)bs, out, (cudaStream_t)stream);
    else throw Error("dtype must be HIPGP_F32 or HIPGP_F64");
    API_END
}
}
