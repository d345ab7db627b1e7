// block_api.inl -- block-diagonal variational family (ziggy/hipgp.py:527-690): the two contractions over k_n (B, M') that
// the reference writes as  to_blocks -> batched matmul -> from_blocks  (util.py:79-126 index maps), fused so that the
// permuted copies and the (B, num_blocks, bs, bs) outer-product tensor of hipgp.py:252-255 are never materialised.
//   blk_idx : (num_blocks, bs) int64, a permutation of [0, M') (define_block_chunks, util.py:79-117)
//   lam[k][i][j] = scale * sum_n w[n] kn[n, idx[k,i]] kn[n, idx[k,j]] + diag * (i == j)        (get_lam, hipgp.py:666-685)
//   out[b, idx[k,i]] = sum_j S[k][i][j] v[b, idx[k,j]]                                          (block_diag_multiply, :640-652)
namespace hipgp {

// ---- shared 64 x 64 register-tiled contraction:  C[m][n] += sum_k X[k][m] * Y[k][n]  with both operands k-major in
// shared memory (16 x 16 threads, 4 x 4 outputs each: two 16-byte shared loads per 16 FMAs)
constexpr int BK_TILE = 64, BK_KT = 16, BK_LD = 68;
template <class T>
__device__ __forceinline__ void block_tile_fma(const T* __restrict__ Xs, const T* __restrict__ Ys, int kt, int m0, int n0,
                                               T (&acc)[4][4]) {
    for (int kk = 0; kk < kt; ++kk) {
        T xa[4], yb[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) { xa[a] = Xs[kk * BK_LD + m0 + a]; yb[a] = Ys[kk * BK_LD + n0 + a]; }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] += xa[a] * yb[b];
    }
}

// grid (num_blocks, tile pairs ti <= tj): lam tile = sum_n (w_n kn[n, idx[i]]) kn[n, idx[j]], rows gathered through the
// index map 16 at a time; accumulators stay in registers over the whole minibatch; the mirror tile is written too
template <class T>
__global__ void __launch_bounds__(256) block_lam_kernel(const T* __restrict__ kn, const T* __restrict__ w,
                                                        const long long* __restrict__ idx, long B, long E, int bs, int ntile,
                                                        T scale, T diag, T* __restrict__ out) {
    __align__(16) __shared__ T Xs[BK_KT * BK_LD];
    __align__(16) __shared__ T Ys[BK_KT * BK_LD];
    __shared__ long long si[BK_TILE], sj[BK_TILE];
    const long k = blockIdx.x;
    int ti = 0, tj = 0;
    {   // pair index -> (ti <= tj)
        int p = blockIdx.y;
        while (p >= ntile - ti) { p -= ntile - ti; ++ti; }
        tj = ti + p;
    }
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, m0 = ty * 4, n0 = tx * 4;
    const int i0 = ti * BK_TILE, j0 = tj * BK_TILE;
    if (tid < BK_TILE) si[tid] = i0 + tid < bs ? idx[k * bs + i0 + tid] : -1;
    else if (tid < 2 * BK_TILE) sj[tid - BK_TILE] = j0 + tid - BK_TILE < bs ? idx[k * bs + j0 + tid - BK_TILE] : -1;
    __syncthreads();
    T acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = (T)0;
    for (long r0 = 0; r0 < B; r0 += BK_KT) {
        const int kt = (int)((B - r0) < BK_KT ? (B - r0) : BK_KT);
        for (int t = tid; t < BK_KT * BK_TILE; t += 256) {
            const int n = t / BK_TILE, m = t - n * BK_TILE;
            T x = (T)0, y = (T)0;
            if (n < kt) {
                const T* row = kn + (size_t)(r0 + n) * E;
                if (si[m] >= 0) x = w[r0 + n] * row[si[m]];
                if (sj[m] >= 0) y = row[sj[m]];
            }
            Xs[n * BK_LD + m] = x; Ys[n * BK_LD + m] = y;
        }
        __syncthreads();
        block_tile_fma<T>(Xs, Ys, kt, m0, n0, acc);
        __syncthreads();
    }
    T* o = out + (size_t)k * bs * bs;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = i0 + m0 + a, j = j0 + n0 + b;
            if (i < bs && j < bs) {
                const T v = scale * acc[a][b] + (i == j ? diag : (T)0);
                o[(size_t)i * bs + j] = v;
                if (ti != tj) o[(size_t)j * bs + i] = v;
            }
        }
}

// grid (num_blocks, row tiles of the block, tiles of 64 right-hand sides): out[b, idx[i]] = sum_j S[k][i][j] v[b, idx[j]]
template <class T>
__global__ void __launch_bounds__(256) block_diag_multiply_kernel(const T* __restrict__ S, const T* __restrict__ v,
                                                                  const long long* __restrict__ idx, long B, long E, int bs,
                                                                  T* __restrict__ out) {
    __align__(16) __shared__ T Xs[BK_KT * BK_LD];       // X[jj][m] = S[k][i0 + m][j0 + jj]
    __align__(16) __shared__ T Ys[BK_KT * BK_LD];       // Y[jj][n] = v[b0 + n][idx[j0 + jj]]
    __shared__ long long sj[BK_KT];
    const long k = blockIdx.x;
    const int i0 = blockIdx.y * BK_TILE;
    const long b0 = (long)blockIdx.z * BK_TILE;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, m0 = ty * 4, n0 = tx * 4;
    const T* Sk = S + (size_t)k * bs * bs;
    T acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = (T)0;
    for (int j0 = 0; j0 < bs; j0 += BK_KT) {
        const int kt = bs - j0 < BK_KT ? bs - j0 : BK_KT;
        if (tid < BK_KT) sj[tid] = tid < kt ? idx[k * bs + j0 + tid] : -1;
        __syncthreads();
        for (int t = tid; t < BK_KT * BK_TILE; t += 256) {
            const int jj = t & (BK_KT - 1), m = t >> 4;          // jj fastest: coalesced along a row of S / the gathered v
            T x = (T)0, y = (T)0;
            if (jj < kt) {
                if (i0 + m < bs) x = Sk[(size_t)(i0 + m) * bs + j0 + jj];
                if (b0 + m < B) y = v[(size_t)(b0 + m) * E + sj[jj]];
            }
            Xs[jj * BK_LD + m] = x; Ys[jj * BK_LD + m] = y;
        }
        __syncthreads();
        block_tile_fma<T>(Xs, Ys, kt, m0, n0, acc);
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = i0 + m0 + a;
        if (i >= bs) continue;
        const long long col = idx[k * bs + i];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const long bb = b0 + n0 + b;
            if (bb < B) out[(size_t)bb * E + col] = acc[a][b];
        }
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
    auto k = block_lam_kernel<T>;
    const int ntile = (int)((bs + BK_TILE - 1) / BK_TILE);
    if (B == 0) {                                          // empty minibatch: lam = diag * I
        std::vector<T> h((size_t)nblk * bs * bs, (T)0);
        for (long q = 0; q < nblk; ++q) for (long i = 0; i < bs; ++i) h[(size_t)q * bs * bs + i * bs + i] = (T)diag;
        CK(cudaMemcpyAsync(out, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
        return;
    }
    const long npair = (long)ntile * (ntile + 1) / 2;
    for (long k0 = 0; k0 < nblk; k0 += 1L << 30) {
        const long nk = std::min<long>(nblk - k0, 1L << 30);
        HIPGP_LAUNCH(k, dim3((unsigned)nk, (unsigned)npair), dim3(256), 0, s, (const T*)kn, (const T*)w,
                     (const long long*)idx + (size_t)k0 * bs, B, E, (int)bs, ntile, (T)scale, (T)diag,
                     (T*)out + (size_t)k0 * bs * bs);
        CK_LAUNCH();
    }
}

template <class T>
static void block_diag_multiply(const void* S, const void* v, const void* idx, long B, long E, long nblk, long bs, void* out,
                                cudaStream_t s) {
    block_check(idx, B, E, nblk, bs);
    if (B == 0) return;
    if (!S || !v || !out) throw Error("null pointer");
    auto k = block_diag_multiply_kernel<T>;
    const unsigned nti = (unsigned)((bs + BK_TILE - 1) / BK_TILE);
    for (long b0 = 0; b0 < B; b0 += BK_TILE * 65535L) {    // grid.z limit
        const long nb = std::min<long>(B - b0, BK_TILE * 65535L);
        HIPGP_LAUNCH(k, dim3((unsigned)nblk, nti, (unsigned)((nb + BK_TILE - 1) / BK_TILE)), dim3(256), 0, s, (const T*)S,
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
