// slab_api.inl -- grids sharded over several GPUs: the local stages of the slab-decomposed 3-D matvec, the
// peer-memory exchange, and their C ABI (declared in include/hipgp_b200.h; host orchestration: hipgp_b200/slab.py).
namespace hipgp {
// ---- slab-decomposed 3-D matvec (K / C^-1), axis 0 split over `nranks` ranks -------------------------------
// stage 1 (local)  : rows r2c + axis-1 forward on this rank's slab of n0/P planes, written straight into the all-to-all
//                    send buffer [dest q][i0_loc][L1/P positions of axis 1][P3]
// (all-to-all)     : every rank now holds ALL i0 for its chunk of (axis-1 position, bin) lines
// stage 2 (local)  : axis-0 forward, spectrum multiply, axis-0 inverse on that chunk, in place; the result is already in
//                    [dest q][i0_loc][chunk] order
// (all-to-all back)
// stage 3 (local)  : axis-1 inverse reading the receive buffer in place, rows c2r, crop
struct SlabGeo { long n0_loc, Lq, chunk, P3, exch_elems; };
template <class T>
static SlabGeo slab_geo(hipgp_plan* pl, Geom<T>& g) {
    if (g.L[1] % pl->slab_nranks) throw Error("slab decomposition v1: embedding length of axis 1 must be divisible by the number of ranks");
    SlabGeo q;
    q.n0_loc = pl->m[0] / pl->slab_nranks;
    q.Lq = g.L[1] / pl->slab_nranks;
    q.P3 = g.P;
    q.chunk = q.Lq * q.P3;
    q.exch_elems = (long)pl->slab_nranks * q.n0_loc * q.chunk;
    return q;
}

template <class T>
static void slab_stage1(hipgp_plan* pl, const void* in_slab, void* send_buf, cudaStream_t s) {
    Geom<T>& g = geom(pl, false, Tag<T>());
    const SlabGeo q = slab_geo<T>(pl, g);
    const long rows = q.n0_loc * pl->m[1];
    pl->W1.ensure(sizeof(cplx<T>) * (size_t)rows * g.P, &pl->dev_bytes);
    cplx<T>* W1 = pl->W1.as<cplx<T>>();
    RowsParams<T> R{};
    rows_geom(R, g);
    R.in = (const T*)in_slab; R.W = W1; R.W_rows = (int)rows;
    R.mode = RF_PLAIN; R.do_fft = 1; R.total_rows = rows; R.nrows = (int)rows; R.n_real = pl->m[2]; R.st = null_state();
    launch_rows<T>(pl, false, R, s, geom_allows_fast(g));
    ColsParams<T> C{};
    C.in = W1; C.out = (cplx<T>*)send_buf; C.n_in = pl->m[1]; C.n_out = pl->m[1]; C.inner = g.H + 1; C.pitch = g.P;
    C.in_ostride = (long)pl->m[1] * g.P; C.in_bstride = 0; C.out_ostride = q.chunk; C.out_bstride = 0;
    C.out_split_len = (int)q.Lq; C.out_split_stride = q.n0_loc * q.chunk;
    C.f = g.fcol[1].dev; C.mode = CM_FWD;
    launch_cols<T>(pl, C, q.n0_loc, 1, s, geom_allows_fast(g));
}

template <class T>
static void slab_stage2(hipgp_plan* pl, int mode, void* buf, cudaStream_t s) {
    Geom<T>& g = geom(pl, false, Tag<T>());
    const SlabGeo q = slab_geo<T>(pl, g);
    // this rank's chunk of lines inside the full [L0][L1 * P3] spectrum
    const T* spec = (mode == HIPGP_MV_K ? pl->specK.as<T>() : pl->specCinv.as<T>()) + (size_t)pl->slab_rank * q.chunk;
    ColsParams<T> C{};
    C.in = (cplx<T>*)buf; C.out = (cplx<T>*)buf; C.n_in = pl->m[0]; C.n_out = pl->m[0]; C.inner = q.chunk; C.pitch = q.chunk;
    C.f = g.fcol[0].dev; C.mode = CM_FUSED; C.spec = spec; C.spec_kind = SPEC_REAL; C.spec_pitch = (long)g.L[1] * q.P3;
    launch_cols<T>(pl, C, 1, 1, s, geom_allows_fast(g));
}

template <class T>
static void slab_stage3(hipgp_plan* pl, const void* recv_buf, void* out_slab, cudaStream_t s) {
    Geom<T>& g = geom(pl, false, Tag<T>());
    const SlabGeo q = slab_geo<T>(pl, g);
    const long rows = q.n0_loc * pl->m[1];
    cplx<T>* W1 = pl->W1.as<cplx<T>>();
    ColsParams<T> C{};
    C.in = (const cplx<T>*)recv_buf; C.out = W1; C.n_in = pl->m[1]; C.n_out = pl->m[1]; C.inner = g.H + 1; C.pitch = g.P;
    C.in_ostride = q.chunk; C.in_bstride = 0; C.out_ostride = (long)pl->m[1] * g.P; C.out_bstride = 0;
    C.in_split_len = (int)q.Lq; C.in_split_stride = q.n0_loc * q.chunk;
    C.f = g.fcol[1].dev; C.mode = CM_INV;
    launch_cols<T>(pl, C, q.n0_loc, 1, s, geom_allows_fast(g));
    RowsParams<T> R{};
    rows_geom(R, g);
    R.out = (T*)out_slab; R.W = W1; R.W_rows = (int)rows;
    R.mode = RI_PLAIN; R.do_fft = 1; R.total_rows = rows; R.nrows = (int)rows; R.n_real = pl->m[2]; R.st = null_state();
    R.spec = nullptr; R.spec_kind = SPEC_NONE;
    launch_rows<T>(pl, true, R, s, geom_allows_fast(g));
}

// ---- slab decomposition, version 2: exchange the UN-PADDED row-pass output, split along the last-axis BINS ------------------
// stage A (local)  : rows r2c on this rank's n0/P planes -> W1[n0_loc m1][P3]; pack -> send[dest q][n0_loc m1][Pq]
//                    (Pq = bins per rank; the zero-padded axis-1 transform is NOT part of what travels: half the bytes of v1)
// (all-to-all)     : recv[src p][n0_loc m1][Pq] = [i0 (all m0)][i1][Pq]: every rank holds ALL planes for its Pq bins
// stage B (local)  : axis-1 forward (pad m1 -> L1), axis-0 forward x spectrum slice x inverse, axis-1 inverse (crop) -- the
//                    ordinary three column passes of the undecomposed pipeline on a [m0][.][Pq] array -- back into recv
// (all-to-all back)
// stage C (local)  : unpack -> W1, rows c2r, crop
struct Slab2Geo { long n0_loc, rows, Pq, Pqc, exch; int nch; };
// bins per rank: a multiple of 2 per chunk (16-byte lanes in fp32)
static long slab2_pq(long H1, int nranks, int nch) { const long u = 2L * nch; return ((H1 + nranks - 1) / nranks + u - 1) / u * u; }
template <class T>
static Slab2Geo slab2_geo(hipgp_plan* pl, Geom<T>& g) {
    Slab2Geo q;
    q.n0_loc = pl->m[0] / pl->slab_nranks;
    q.rows = q.n0_loc * pl->m[1];
    q.nch = pl->slab_chunks;
    q.Pq = slab2_pq(g.H + 1, pl->slab_nranks, q.nch);
    q.Pqc = q.Pq / q.nch;
    q.exch = (long)pl->slab_nranks * q.rows * q.Pq;
    return q;
}
// send[((c nranks + q) rows + r) Pqc + b] = W[r P + q Pq + c Pqc + b]  (zero past the row pitch) / the inverse.
// Chunk-major: chunk c of every destination is one contiguous all-to-all of its own.
template <class T>
__global__ void slab_pack_kernel(const cplx<T>* __restrict__ W, cplx<T>* __restrict__ buf, long rows, long P, long Pq, long Pqc, int nranks,
                                 int unpack) {
    const long total = (long)nranks * rows * Pq;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long b = i % Pqc, r = (i / Pqc) % rows, q = (i / (Pqc * rows)) % nranks, ch = i / (Pqc * rows * nranks);
        const long c = q * Pq + ch * Pqc + b;
        if (unpack) { if (c < P) const_cast<cplx<T>*>(W)[r * P + c] = buf[i]; }
        else buf[i] = c < P ? W[r * P + c] : mk<T>(0, 0);
    }
}
// out[(c lines + l) Pqc + b] = spec[l P + c0 + c Pqc + b]
template <class T>
__global__ void slab_spec_slice_kernel(const T* __restrict__ spec, T* __restrict__ out, long lines, long P, long Pq, long Pqc, long c0) {
    const long total = lines * Pq;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long b = i % Pqc, l = (i / Pqc) % lines, ch = i / (Pqc * lines);
        const long c = c0 + ch * Pqc + b;
        out[i] = c < P ? spec[l * P + c] : (T)0;
    }
}

template <class T>
static void slab2_stageA(hipgp_plan* pl, const void* in_slab, void* send_buf, cudaStream_t s) {
    Geom<T>& g = geom(pl, false, Tag<T>());
    const Slab2Geo q = slab2_geo<T>(pl, g);
    pl->W1.ensure(sizeof(cplx<T>) * (size_t)q.rows * g.P, &pl->dev_bytes);
    cplx<T>* W1 = pl->W1.as<cplx<T>>();
    RowsParams<T> R{};
    rows_geom(R, g);
    R.in = (const T*)in_slab; R.W = W1; R.W_rows = (int)q.rows;
    R.mode = RF_PLAIN; R.do_fft = 1; R.total_rows = q.rows; R.nrows = (int)q.rows; R.n_real = pl->m[2]; R.st = null_state();
    launch_rows<T>(pl, false, R, s, geom_allows_fast(g));
    auto k = slab_pack_kernel<T>;
    const unsigned nb = (unsigned)std::min<long>((q.exch + 255) / 256, 148L * 16);
    HIPGP_LAUNCH(k, dim3(nb), dim3(256), 0, s, (const cplx<T>*)W1, (cplx<T>*)send_buf, q.rows, (long)g.P, q.Pq, q.Pqc, pl->slab_nranks, 0);
    CK_LAUNCH(); pl->launches++;
}

// One chunk of bins (chunk < 0: every chunk in turn); `buf` is the whole exchange buffer [chunk][i0][i1][Pqc].
template <class T>
static void slab2_stageB(hipgp_plan* pl, int mode, void* buf, int chunk, cudaStream_t s) {
    Geom<T>& g = geom(pl, false, Tag<T>());
    const Slab2Geo q = slab2_geo<T>(pl, g);
    if (chunk >= q.nch) throw Error("slab chunk out of range");
    const long L1 = g.L[1], L0 = g.L[0], Pq = q.Pq, Pc = q.Pqc;
    const int m0 = pl->m[0], m1 = pl->m[1];
    const bool fast = geom_allows_fast(g);
    // this rank's bins of the spectrum, repacked once per spectrum to [chunk][L0][L1][Pqc]
    DevBuf& slice = mode == HIPGP_MV_K ? pl->slabSpecK : pl->slabSpecCinv;
    bool& have = mode == HIPGP_MV_K ? pl->have_slabK : pl->have_slabCinv;
    if (!have) {
        slice.ensure(sizeof(T) * (size_t)(L0 * L1 * Pq), &pl->dev_bytes);
        auto k = slab_spec_slice_kernel<T>;
        const long total = L0 * L1 * Pq;
        const unsigned nb = (unsigned)std::min<long>((total + 255) / 256, 148L * 16);
        HIPGP_LAUNCH(k, dim3(nb), dim3(256), 0, s, (mode == HIPGP_MV_K ? pl->specK.as<T>() : pl->specCinv.as<T>()), slice.as<T>(), L0 * L1, (long)g.P, Pq, Pc,
                     (long)pl->slab_rank * Pq);
        CK_LAUNCH(); pl->launches++;
        have = true;
    }
    pl->W2.ensure(sizeof(cplx<T>) * (size_t)((long)m0 * L1 * Pc), &pl->dev_bytes);
    cplx<T>* W2 = pl->W2.as<cplx<T>>();
    for (int c = chunk < 0 ? 0 : chunk; c < (chunk < 0 ? q.nch : chunk + 1); ++c) {
        cplx<T>* part = (cplx<T>*)buf + (size_t)c * m0 * m1 * Pc;
        ColsParams<T> C{};
        C.spec = nullptr; C.spec_kind = SPEC_NONE;
        // axis 1 forward: part[i0][m1][Pc] -> W2[i0][L1][Pc]
        C.in = part; C.out = W2; C.n_in = m1; C.n_out = m1; C.inner = Pc; C.pitch = Pc;
        C.in_ostride = (long)m1 * Pc; C.out_ostride = L1 * Pc; C.f = g.fcol[1].dev; C.mode = CM_FWD;
        launch_cols<T>(pl, C, m0, 1, s, fast);
        // axis 0 fused, in place on W2
        C = ColsParams<T>{};
        C.in = W2; C.out = W2; C.n_in = m0; C.n_out = m0; C.inner = L1 * Pc; C.pitch = L1 * Pc;
        C.f = g.fcol[0].dev; C.mode = CM_FUSED; C.spec = slice.as<T>() + (size_t)c * L0 * L1 * Pc; C.spec_kind = SPEC_REAL;
        launch_cols<T>(pl, C, 1, 1, s, fast);
        // axis 1 inverse: W2 -> part[i0][m1][Pc]
        C = ColsParams<T>{};
        C.in = W2; C.out = part; C.n_in = m1; C.n_out = m1; C.inner = Pc; C.pitch = Pc;
        C.in_ostride = L1 * Pc; C.out_ostride = (long)m1 * Pc; C.f = g.fcol[1].dev; C.mode = CM_INV; C.spec = nullptr; C.spec_kind = SPEC_NONE;
        launch_cols<T>(pl, C, m0, 1, s, fast);
    }
}

template <class T>
static void slab2_stageC(hipgp_plan* pl, const void* recv_buf, void* out_slab, cudaStream_t s) {
    Geom<T>& g = geom(pl, false, Tag<T>());
    const Slab2Geo q = slab2_geo<T>(pl, g);
    pl->W1.ensure(sizeof(cplx<T>) * (size_t)q.rows * g.P, &pl->dev_bytes);
    cplx<T>* W1 = pl->W1.as<cplx<T>>();
    auto k = slab_pack_kernel<T>;
    const unsigned nb = (unsigned)std::min<long>((q.exch + 255) / 256, 148L * 16);
    PROF_BEGIN(pl, 3, s);
    HIPGP_LAUNCH(k, dim3(nb), dim3(256), 0, s, (const cplx<T>*)W1, (cplx<T>*)const_cast<void*>(recv_buf), q.rows, (long)g.P, q.Pq, q.Pqc, pl->slab_nranks, 1);
    PROF_END(pl, s);
    CK_LAUNCH(); pl->launches++;
    RowsParams<T> R{};
    rows_geom(R, g);
    R.out = (T*)out_slab; R.W = W1; R.W_rows = (int)q.rows;
    R.mode = RI_PLAIN; R.do_fft = 1; R.total_rows = q.rows; R.nrows = (int)q.rows; R.n_real = pl->m[2]; R.st = null_state();
    R.spec = nullptr; R.spec_kind = SPEC_NONE;
    launch_rows<T>(pl, true, R, s, geom_allows_fast(g));
}

// ---- peer-memory exchange: the packing kernel's stores ARE the transfer (NVLink stores into the peers' receive buffers) ----
struct PeerPtrs { void* p[16]; };
struct alignas(16) Unit16 { unsigned long long a, b; };
// 16-byte units; W[r P + q Pq + c Pqc + b] -> peer q's buffer at [((c nranks + me) rows + r) Pqc + b].
// Four independent loads in flight per thread before the (remote) stores.
template <class T>
__global__ void slab_push_pack_kernel(const cplx<T>* __restrict__ W, PeerPtrs dst, long rows, long P, long Pq, long Pqc, int nranks, int me) {
    constexpr int PER = 16 / (int)sizeof(cplx<T>);
    constexpr int U = 4;
    const long upc = Pqc / PER;                         // units per row per chunk
    const long total = (long)nranks * rows * (Pq / PER);
    const long step = (long)gridDim.x * blockDim.x;
    for (long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += U * step) {
        Unit16 v[U]; Unit16* out[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const long i = i0 + k * step;
            v[k] = Unit16{0ull, 0ull}; out[k] = nullptr;
            if (i < total) {
                // destination varies fastest after the unit-in-row index and is rotated by the sender's rank: at any moment every
                // rank writes to every peer, and no peer is everybody's target at once
                const long u = i % upc, q = ((i / upc) % nranks + me) % nranks, r = (i / (upc * nranks)) % rows, ch = i / (upc * rows * nranks);
                const long c = q * Pq + ch * Pqc + u * PER;
                if (c + PER <= P) v[k] = *reinterpret_cast<const Unit16*>(W + r * P + c);      // P and c are even: no unit straddles the pitch
                out[k] = reinterpret_cast<Unit16*>(reinterpret_cast<cplx<T>*>(dst.p[q]) + (((long)ch * nranks + me) * rows + r) * Pqc + u * PER);
            }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) if (out[k]) *out[k] = v[k];
    }
}
// the way back: local buffer [c][q][rows][Pqc] (block q belongs to rank q) -> straight into rank q's ROW WORKSPACE
// W[r P + me Pq + c Pqc + b], so the receiving side's inverse row pass starts without an unpack pass
template <class T>
__global__ void slab_push_back_kernel(const cplx<T>* __restrict__ buf, PeerPtrs dst, long rows, long P, long Pq, long Pqc, int nranks, int nch, int me,
                                      int ch0, int ch1) {
    constexpr int PER = 16 / (int)sizeof(cplx<T>);
    constexpr int U = 4;
    const long upc = Pqc / PER;
    const long blk = rows * upc;                          // units per (chunk, rank) block
    const long total = (long)(ch1 - ch0) * nranks * blk;
    const long step = (long)gridDim.x * blockDim.x;
    for (long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += U * step) {
        Unit16 v[U]; Unit16* out[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const long i = i0 + k * step;
            out[k] = nullptr;
            if (i < total) {
                const long u = i % upc, q = ((i / upc) % nranks + me) % nranks, r = (i / (upc * nranks)) % rows, ch = ch0 + i / (blk * nranks);
                const long c = (long)me * Pq + ch * Pqc + u * PER;
                if (c + PER <= P) {
                    v[k] = reinterpret_cast<const Unit16*>(buf)[((long)ch * nranks + q) * blk + r * upc + u];
                    out[k] = reinterpret_cast<Unit16*>(reinterpret_cast<cplx<T>*>(dst.p[q]) + r * P + c);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) if (out[k]) *out[k] = v[k];
    }
}
// grid of the transfer kernels: 8 CTAs per SM, fewer when the exchange is small (each thread moves four 16-byte units)
static unsigned push_grid(long exch_complex, int per) {
    const long units = exch_complex / per;
    return (unsigned)std::max<long>(1, std::min<long>(148L * 8, (units + 4 * 256 - 1) / (4 * 256)));
}
static PeerPtrs peer_table(hipgp_plan* pl, bool back) {
    if (!pl->peers_ready) throw Error("slab peer buffers are not connected: call hipgp_slab2_peer_open / _peer_set first");
    PeerPtrs t{};
    for (int q = 0; q < pl->slab_nranks; ++q) t.p[q] = back ? pl->peerR2[q] : pl->peerR1[q];
    return t;
}
template <class T>
static void slab2_pushA(hipgp_plan* pl, const void* in_slab, cudaStream_t s) {
    Geom<T>& g = geom(pl, false, Tag<T>());
    const Slab2Geo q = slab2_geo<T>(pl, g);
    pl->W1.ensure(sizeof(cplx<T>) * (size_t)q.rows * g.P, &pl->dev_bytes);
    cplx<T>* W1 = pl->W1.as<cplx<T>>();
    RowsParams<T> R{};
    rows_geom(R, g);
    R.in = (const T*)in_slab; R.W = W1; R.W_rows = (int)q.rows;
    R.mode = RF_PLAIN; R.do_fft = 1; R.total_rows = q.rows; R.nrows = (int)q.rows; R.n_real = pl->m[2]; R.st = null_state();
    launch_rows<T>(pl, false, R, s, geom_allows_fast(g));
    auto k = slab_push_pack_kernel<T>;
    PROF_BEGIN(pl, 3, s);
    HIPGP_LAUNCH(k, dim3(push_grid(q.exch, 16 / (int)sizeof(cplx<T>))), dim3(256), 0, s, (const cplx<T>*)W1, peer_table(pl, false), q.rows, (long)g.P, q.Pq, q.Pqc, pl->slab_nranks, pl->slab_rank);
    PROF_END(pl, s);
    CK_LAUNCH(); pl->launches++;
}
// the transfer kernels alone (measurement: bytes that leave the GPU / their duration = achieved NVLink rate)
template <class T>
static void slab2_push_only(hipgp_plan* pl, int back, cudaStream_t s) {
    Geom<T>& g = geom(pl, false, Tag<T>());
    const Slab2Geo q = slab2_geo<T>(pl, g);
    if (!back) {
        pl->W1.ensure(sizeof(cplx<T>) * (size_t)q.rows * g.P, &pl->dev_bytes);
        auto k = slab_push_pack_kernel<T>;
        HIPGP_LAUNCH(k, dim3(push_grid(q.exch, 16 / (int)sizeof(cplx<T>))), dim3(256), 0, s, (const cplx<T>*)pl->W1.p, peer_table(pl, false), q.rows, (long)g.P, q.Pq, q.Pqc, pl->slab_nranks, pl->slab_rank);
    } else {
        auto k = slab_push_back_kernel<T>;
        HIPGP_LAUNCH(k, dim3(push_grid(q.exch, 16 / (int)sizeof(cplx<T>))), dim3(256), 0, s, (const cplx<T>*)pl->slabR1.p, peer_table(pl, true), q.rows, (long)g.P, q.Pq, q.Pqc, pl->slab_nranks, q.nch, pl->slab_rank, 0, q.nch);
    }
    CK_LAUNCH(); pl->launches++;
}
template <class T>
static void slab2_pushB(hipgp_plan* pl, int mode, int chunk, cudaStream_t s) {
    Geom<T>& g = geom(pl, false, Tag<T>());
    const Slab2Geo q = slab2_geo<T>(pl, g);
    slab2_stageB<T>(pl, mode, pl->slabR1.p, chunk, s);
    auto k = slab_push_back_kernel<T>;
    const int c0 = chunk < 0 ? 0 : chunk, c1 = chunk < 0 ? q.nch : chunk + 1;
    PROF_BEGIN(pl, 3, s);
    HIPGP_LAUNCH(k, dim3(push_grid(q.exch, 16 / (int)sizeof(cplx<T>))), dim3(256), 0, s, (const cplx<T>*)pl->slabR1.p, peer_table(pl, true), q.rows, (long)g.P, q.Pq, q.Pqc, pl->slab_nranks, q.nch, pl->slab_rank, c0, c1);
    PROF_END(pl, s);
    CK_LAUNCH(); pl->launches++;
}

template <class T>
static void slab2_finish(hipgp_plan* pl, void* out_slab, cudaStream_t s) {
    Geom<T>& g = geom(pl, false, Tag<T>());
    const Slab2Geo q = slab2_geo<T>(pl, g);
    RowsParams<T> R{};
    rows_geom(R, g);
    R.out = (T*)out_slab; R.W = pl->slabR2.as<cplx<T>>(); R.W_rows = (int)q.rows;
    R.mode = RI_PLAIN; R.do_fft = 1; R.total_rows = q.rows; R.nrows = (int)q.rows; R.n_real = pl->m[2]; R.st = null_state();
    R.spec = nullptr; R.spec_kind = SPEC_NONE;
    launch_rows<T>(pl, true, R, s, geom_allows_fast(g));
}

}  // namespace hipgp

extern "C" {
int hipgp_plan_set_slab(hipgp_plan* pl, int rank, int nranks) {
    API_BEGIN
    need_plan(pl);
    if (pl->D != 3) throw Error("slab decomposition needs a 3-D grid");
    if (nranks < 1 || rank < 0 || rank >= nranks) throw Error("bad rank / nranks");
    if (pl->m[0] % nranks) throw Error("grid extent of axis 0 must be divisible by the number of ranks");
    pl->slab_rank = rank; pl->slab_nranks = nranks;
    API_END
}
int hipgp_slab_sizes(const hipgp_plan* pl, int64_t* slab_reals, int64_t* exchange_complex) {
    API_BEGIN
    need_plan(pl);
    const long n0 = pl->m[0] / pl->slab_nranks;
    if (slab_reals) *slab_reals = n0 * pl->m[1] * pl->m[2];
    const long P3 = ((long)pl->Ln[2] / 2 + 1 + 7) / 8 * 8;
    if (exchange_complex) *exchange_complex = (long)pl->slab_nranks * n0 * (pl->Ln[1] / pl->slab_nranks) * P3;
    API_END
}
int hipgp_slab_stage1(hipgp_plan* pl, const void* in_slab, void* send_buf, void* stream) {
    API_BEGIN
    set_device(pl);
    if (!pl->have_spec) throw Error("plan has no spectrum");
    DISPATCH(pl, slab_stage1<float>(pl, in_slab, send_buf, (cudaStream_t)stream), slab_stage1<double>(pl, in_slab, send_buf, (cudaStream_t)stream));
    API_END
}
int hipgp_slab_stage2(hipgp_plan* pl, int mode, void* buf, void* stream) {
    API_BEGIN
    set_device(pl);
    if (mode != HIPGP_MV_K && mode != HIPGP_MV_CINV) throw Error("slab mode supports K and C^-1");
    DISPATCH(pl, slab_stage2<float>(pl, mode, buf, (cudaStream_t)stream), slab_stage2<double>(pl, mode, buf, (cudaStream_t)stream));
    API_END
}
int hipgp_slab_stage3(hipgp_plan* pl, const void* recv_buf, void* out_slab, void* stream) {
    API_BEGIN
    set_device(pl);
    DISPATCH(pl, slab_stage3<float>(pl, recv_buf, out_slab, (cudaStream_t)stream), slab_stage3<double>(pl, recv_buf, out_slab, (cudaStream_t)stream));
    API_END
}
int hipgp_slab2_sizes(const hipgp_plan* pl, int64_t* slab_reals, int64_t* exchange_complex) {
    API_BEGIN
    need_plan(pl);
    if (pl->D != 3) throw Error("slab decomposition needs a 3-D grid");
    const long n0 = pl->m[0] / pl->slab_nranks;
    if (slab_reals) *slab_reals = n0 * pl->m[1] * pl->m[2];
    const long H1 = (long)pl->Ln[2] / 2 + 1;
    const long Pq = slab2_pq(H1, pl->slab_nranks, pl->slab_chunks);
    if (exchange_complex) *exchange_complex = (long)pl->slab_nranks * n0 * pl->m[1] * Pq;
    API_END
}
int hipgp_slab2_stage_a(hipgp_plan* pl, const void* in_slab, void* send_buf, void* stream) {
    API_BEGIN
    set_device(pl);
    if (!pl->have_spec) throw Error("plan has no spectrum");
    if (pl->D != 3) throw Error("slab decomposition needs a 3-D grid");
    DISPATCH(pl, slab2_stageA<float>(pl, in_slab, send_buf, (cudaStream_t)stream), slab2_stageA<double>(pl, in_slab, send_buf, (cudaStream_t)stream));
    API_END
}
int hipgp_slab2_stage_b(hipgp_plan* pl, int mode, void* buf, void* stream) { return hipgp_slab2_stage_b_chunk(pl, mode, buf, -1, stream); }
int hipgp_slab2_stage_b_chunk(hipgp_plan* pl, int mode, void* buf, int chunk, void* stream) {
    API_BEGIN
    set_device(pl);
    if (!pl->have_spec) throw Error("plan has no spectrum");
    if (mode != HIPGP_MV_K && mode != HIPGP_MV_CINV) throw Error("slab mode supports K and C^-1");
    DISPATCH(pl, slab2_stageB<float>(pl, mode, buf, chunk, (cudaStream_t)stream), slab2_stageB<double>(pl, mode, buf, chunk, (cudaStream_t)stream));
    API_END
}
int hipgp_plan_set_slab_chunks(hipgp_plan* pl, int nchunks) {
    API_BEGIN
    need_plan(pl);
    if (nchunks < 1 || nchunks > 16) throw Error("slab chunks must be 1..16");
    pl->slab_chunks = nchunks; pl->have_slabK = pl->have_slabCinv = false;
    API_END
}
int hipgp_slab2_stage_c(hipgp_plan* pl, const void* recv_buf, void* out_slab, void* stream) {
    API_BEGIN
    set_device(pl);
    DISPATCH(pl, slab2_stageC<float>(pl, recv_buf, out_slab, (cudaStream_t)stream), slab2_stageC<double>(pl, recv_buf, out_slab, (cudaStream_t)stream));
    API_END
}
int hipgp_slab2_peer_alloc(hipgp_plan* pl, void** r1_out, void** r2_out, void* handle1_64, void* handle2_64) {
    API_BEGIN
    need_plan(pl); set_device(pl);
    if (pl->D != 3) throw Error("slab decomposition needs a 3-D grid");
    if (pl->slab_nranks > 16) throw Error("peer exchange supports up to 16 ranks");
    int64_t slab = 0, exch = 0;
    if (hipgp_slab2_sizes(pl, &slab, &exch)) return -1;
    const size_t w = pl->dtype == HIPGP_F32 ? 8 : 16;
    slab_peer_close(pl);
    const long P3 = ((long)pl->Ln[2] / 2 + 1 + 7) / 8 * 8;
    const size_t w1 = w * (size_t)(pl->m[0] / pl->slab_nranks) * (size_t)pl->m[1] * (size_t)P3;       // the row workspace itself
    pl->slabR1.ensure(w * (size_t)exch, &pl->dev_bytes); pl->slabR2.ensure(w1, &pl->dev_bytes);
    CK(cudaMemset(pl->slabR2.p, 0, w1));
    if (r1_out) *r1_out = pl->slabR1.p;
    if (r2_out) *r2_out = pl->slabR2.p;
#ifndef HIPGP_EMU
    if (handle1_64) { cudaIpcMemHandle_t h; CK(cudaIpcGetMemHandle(&h, pl->slabR1.p)); static_assert(sizeof(h) == 64, "ipc handle"); memcpy(handle1_64, &h, 64); }
    if (handle2_64) { cudaIpcMemHandle_t h; CK(cudaIpcGetMemHandle(&h, pl->slabR2.p)); memcpy(handle2_64, &h, 64); }
#else
    if (handle1_64) memset(handle1_64, 0, 64);
    if (handle2_64) memset(handle2_64, 0, 64);
#endif
    API_END
}
/* handles: nranks x 64 bytes each, in rank order (own entry ignored) */
int hipgp_slab2_peer_open(hipgp_plan* pl, const void* handles1, const void* handles2) {
    API_BEGIN
    need_plan(pl); set_device(pl);
    if (!pl->slabR1.p || !pl->slabR2.p) throw Error("call hipgp_slab2_peer_alloc first");
#ifdef HIPGP_EMU
    throw Error("inter-process handles are not available in the emulation build");
#else
    slab_peer_close(pl);
    for (int q = 0; q < pl->slab_nranks; ++q) {
        if (q == pl->slab_rank) { pl->peerR1[q] = pl->slabR1.p; pl->peerR2[q] = pl->slabR2.p; continue; }
        for (int k = 0; k < 2; ++k) {
            cudaIpcMemHandle_t h; memcpy(&h, (const char*)(k ? handles2 : handles1) + 64 * (size_t)q, 64);
            void* ptr = nullptr;
            CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
            pl->ipc_opened[pl->n_ipc_opened++] = ptr;
            (k ? pl->peerR2 : pl->peerR1)[q] = ptr;
        }
    }
    pl->peers_ready = true;
#endif
    API_END
}
/* same-process ranks (tests, one process driving several devices): the buffers' addresses themselves */
int hipgp_slab2_peer_set(hipgp_plan* pl, void* const* r1_all, void* const* r2_all) {
    API_BEGIN
    need_plan(pl);
    if (!pl->slabR1.p || !pl->slabR2.p) throw Error("call hipgp_slab2_peer_alloc first");
    slab_peer_close(pl);
    for (int q = 0; q < pl->slab_nranks; ++q) { pl->peerR1[q] = r1_all[q]; pl->peerR2[q] = r2_all[q]; }
    pl->peers_ready = true;
    API_END
}
int hipgp_slab2_push_a(hipgp_plan* pl, const void* in_slab, void* stream) {
    API_BEGIN
    set_device(pl);
    if (!pl->have_spec) throw Error("plan has no spectrum");
    DISPATCH(pl, slab2_pushA<float>(pl, in_slab, (cudaStream_t)stream), slab2_pushA<double>(pl, in_slab, (cudaStream_t)stream));
    API_END
}
int hipgp_slab2_push_b(hipgp_plan* pl, int mode, int chunk, void* stream) {
    API_BEGIN
    set_device(pl);
    if (!pl->have_spec) throw Error("plan has no spectrum");
    if (mode != HIPGP_MV_K && mode != HIPGP_MV_CINV) throw Error("slab mode supports K and C^-1");
    DISPATCH(pl, slab2_pushB<float>(pl, mode, chunk, (cudaStream_t)stream), slab2_pushB<double>(pl, mode, chunk, (cudaStream_t)stream));
    API_END
}
int hipgp_slab2_push_only(hipgp_plan* pl, int back, void* stream) {
    API_BEGIN
    set_device(pl);
    DISPATCH(pl, slab2_push_only<float>(pl, back, (cudaStream_t)stream), slab2_push_only<double>(pl, back, (cudaStream_t)stream));
    API_END
}
int hipgp_slab2_finish(hipgp_plan* pl, void* out_slab, void* stream) {
    API_BEGIN
    set_device(pl);
    if (!pl->peers_ready) throw Error("slab peer buffers are not connected");
    DISPATCH(pl, slab2_finish<float>(pl, out_slab, (cudaStream_t)stream), slab2_finish<double>(pl, out_slab, (cudaStream_t)stream));
    API_END
}
}
