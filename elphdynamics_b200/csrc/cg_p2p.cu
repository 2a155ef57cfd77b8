// Multi-GPU conjugate gradient on A = M^T M for ONE tau-sharded lattice, fused with its collectives over peer memory.
//
// rank g owns a contiguous slab of time slices.  A CG iteration needs (1) the neighbouring slices of p
// across the slab boundary and (2) two scalar all-reduces (p.Ap and |r|^2).  With NCCL calls between kernel launches
// that is ~77 us per product at config E; here the whole solve is ONE cooperative persistent kernel per GPU (the loop
// of cg_persistent.cu, src/IterativeSolvers.jl:239-314) and both collectives happen inside it over NVLink peer memory:
//
//   * halo: the CTAs of the first / last slice of a slab PUSH their new r and p tiles straight into the neighbour GPU's
//     halo rows (posted remote stores); every CTA then reads tau-1 / tau+1 from local memory only;
//   * all-reduce + barrier: the last CTA to arrive on a GPU folds the GPU's partials in index order and writes
//     {value, sequence number} into a mailbox slot on EVERY GPU (two 64-bit words, each carrying half of the double and
//     the sequence number, as one 16-byte store); each CTA polls its own GPU's mailbox until all `world` slots carry
//     the sequence number and sums them in rank order -- the same bits on every GPU, so all GPUs take the same branch.
//
// Memory: each process allocates one arena (cudaMalloc), exports it with cudaIpcGetMemHandle and opens the others'
// (elph_shard_p2p_*).  Arena = R, P0, P1 as [halo_lo][Lmax slices][halo_hi] plus the mailboxes; Lmax = ceil(Lglob /
// world), so the layout is the same on every rank.  Sequence numbers increase monotonically over the life of the handle
// (all ranks execute the same number of barriers), mailbox slots alternate by parity, nothing is ever reset.
// Holstein on periodic square lattices (the register tiles of mtm_square.cu); the reference has no counterpart.
#include "square_tiles.cuh"

#include <algorithm>
#include <cstring>

namespace {

using namespace sqt;

constexpr int kMaxWorld = 16;
constexpr unsigned int kSpinLimit = 1u << 27;   // ~ seconds: a dead peer ends the solve with an error instead of a hang

struct P2pParams {
    const double* __restrict__ D;   // expnV with halos: slice index -1 .. L valid
    const double* __restrict__ b;   // [L][N] right-hand side (initial guess is zero)
    double* __restrict__ x;         // [L][N] out
    double* R;                      // own slice 0 of the arena's R (rows -1 and L are the halo rows)
    double* P0;
    double* P1;
    double* left_R;                 // left neighbour's halo_hi row of R / P0 / P1 (peer memory)
    double* left_P0;
    double* left_P1;
    double* right_R;                // right neighbour's halo_lo row
    double* right_P0;
    double* right_P1;
    unsigned long long* mbox[kMaxWorld];   // mailbox base of every rank (own included): [2 parities][world][2 words]
    unsigned int* left_hi_flag;     // in the left neighbour's arena: "your halo_hi row is complete up to barrier seq"
    unsigned int* right_lo_flag;    // in the right neighbour's arena: same for its halo_lo row
    const unsigned int* my_lo_flag; // own arena, written by the left neighbour's last slice
    const unsigned int* my_hi_flag; // own arena, written by the right neighbour's first slice
    double* partial;                // [L] per-CTA partials of this GPU
    unsigned int* bar;              // arrival counter of this GPU (monotonic over the launch, zeroed by the host)
    CgScalars* S;                   // in: tol, kappa_max, maxiter; out: iter, eps, normb, done (2 = peer timeout)
    unsigned int seq_base;          // sequence number of the last barrier of the previous solve
    int L, Ly, rank, world, tau0, Lglob;
    double c0, s0, c1, s1, c2, s2, c3, s3;
};

// acq_rel fences (lighter than the sequentially-consistent membar behind __threadfence_system: measured 27.6 -> see
// DESIGN.md us/iteration at world = 1)
__device__ __forceinline__ void fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

__device__ __forceinline__ void st_mbox(unsigned long long* p, unsigned long long w0, unsigned long long w1) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ void ld_mbox(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

__device__ __forceinline__ void st_flag(unsigned int* p, unsigned int v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_flag(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Barrier over all CTAs of all GPUs fused with the sum of one double per CTA.  `nbar` counts this GPU's barriers of the
// launch from 1 (local arrival target = nbar * gridDim.x), seq is the global sequence number.  Returns false on a peer
// timeout.  The all-reduce itself carries no cross-GPU ordering obligation (a mailbox word validates itself through its
// sequence number): remote data is only ever read through the halo rows, and those are guarded point-to-point -- the
// CTA of the first / last slice, AFTER arriving (so its system-scope fence overlaps the wait instead of delaying
// everybody), fences its pushes and stores seq into the neighbour's halo flag.  Measured: with the system fences inside
// the barrier's critical path an iteration cost 17.2 us (27.6 us with sequentially-consistent membar.sys).
__device__ __forceinline__ bool global_sum(double block_value, const P2pParams& P, unsigned int nbar, unsigned int seq, bool first,
                                           bool last, double* red, int* flag, double& out) {
    const int nb = gridDim.x;
    if (threadIdx.x == 0) {
        P.partial[blockIdx.x] = block_value;
        fence_gpu();
        const unsigned int prev = atomicAdd(P.bar, 1u);
        flag[0] = (prev == nbar * (unsigned int)nb - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (flag[0]) {
        // last CTA of this GPU: fixed-order fold of the GPU's partials, then publish to every GPU's mailbox
        fence_gpu();
        double s = 0.0;
        for (int k = threadIdx.x; k < nb; k += blockDim.x) s += __ldcg(P.partial + k);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
            const unsigned long long bits = (unsigned long long)__double_as_longlong(t);
            const unsigned long long w0 = (bits & 0xffffffffull) | ((unsigned long long)seq << 32);
            const unsigned long long w1 = (bits >> 32) | ((unsigned long long)seq << 32);
            fence_gpu();   // local rows written before the arrivals are ordered before this GPU's own mailbox word
            const size_t slot = ((size_t)(seq & 1u) * P.world + P.rank) * 2;
            for (int q = 0; q < P.world; ++q) st_mbox(P.mbox[(P.rank + q) % P.world] + slot, w0, w1);
        }
        __syncthreads();
    }
    if (threadIdx.x == 32 && (first || last)) {
        // a thread of ANOTHER warp, after the CTA's publishing duties: the system-scope fence overlaps thread 0's wait on
        // the mailbox below instead of delaying the barrier (the boundary CTAs do extra work and tend to arrive last).
        // The caller's __syncthreads ordered the whole CTA's pushes before this fence.
        fence_sys();   // the tiles pushed so far have reached the neighbour's memory ...
        if (first) st_flag(P.left_hi_flag, seq);    // ... before it can see its halo flag move
        if (last) st_flag(P.right_lo_flag, seq);
    }
    // every CTA: wait for all GPUs' contributions in the own mailbox, sum in rank order
    if (threadIdx.x < 32) {
        double t = 0.0;
        int ok = 1;
        if (threadIdx.x == 0) {
            const unsigned long long* mine = P.mbox[P.rank] + (size_t)(seq & 1u) * P.world * 2;
            for (int g = 0; g < P.world && ok; ++g) {
                unsigned long long a, b;
                unsigned int spins = 0;
                do {
                    ld_mbox(mine + 2 * g, a, b);
                    if (++spins > kSpinLimit) { ok = 0; break; }
                } while ((unsigned int)(a >> 32) != seq || (unsigned int)(b >> 32) != seq);
                t += __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
            }
            fence_gpu();   // acquire the local rows of the other CTAs of this GPU
            red[0] = t;
            flag[1] = ok;
        }
    }
    __syncthreads();
    out = red[0];
    const bool good = (flag[1] != 0);
    __syncthreads();   // red / flag are reused by the caller
    return good;
}

template <int NSEG, int PY>
__device__ __forceinline__ double tile_sum(double v, double* red, int lane, int warp, int nwarps) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int k = 0; k < nwarps; ++k) t += red[k];
    __syncthreads();
    return t;
}

template <int NSEG, int PY, int MAXT>
__global__ void __launch_bounds__(MAXT) cg_p2p_kernel(P2pParams P) {
    constexpr int LX = 32 * NSEG;
    extern __shared__ __align__(16) double strips[];   // 2 x [nwarps][4][LX]
    __shared__ double red[32];
    __shared__ int flag[2];   // [0] this CTA arrived last on its GPU, [1] barrier completed without a peer timeout
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int L = P.L, N = LX * P.Ly;
    const int tau = blockIdx.x;                      // local slice; rows tau-1 = -1 and tau+1 = L are the halo rows
    const bool first = (tau == 0), last = (tau == L - 1);
    const size_t tile_off = (size_t)warp * PY * LX;
    auto eidx = [&](int r, int q) -> size_t { return tile_off + r * LX + 32 * q + lane; };
    const long long row = (long long)tau * N, rowm = row - N, rowp = row + N;

    Tile<NSEG, PY> x, r, pprev, pc, Dc, Dn, t1, t2;
    double accb = 0.0;
#pragma unroll
    for (int rr = 0; rr < PY; ++rr)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const size_t e = eidx(rr, q);
            const double bv = P.b[row + e];          // x0 = 0: r0 = b
            x.a[rr][q] = 0.0;
            r.a[rr][q] = bv;
            pprev.a[rr][q] = 0.0;
            Dc.a[rr][q] = P.D[row + e];
            Dn.a[rr][q] = P.D[rowp + e];
            P.R[row + e] = bv;
            if (first) P.left_R[e] = bv;             // push r0 into the neighbours' halo rows
            if (last) P.right_R[e] = bv;
            accb = fma(bv, bv, accb);
        }
    const double tol = P.S->tol, kappa_max = P.S->kappa_max;
    const long long maxiter = P.S->maxiter;
    unsigned int nbar = 0, seq = P.seq_base;
    const bool boundary = first || last;
    double rdotr;
    bool alive = global_sum(tile_sum<NSEG, PY>(accb, red, lane, warp, nwarps), P, ++nbar, ++seq, first, last, red, flag, rdotr);
    const double normb = sqrt(rdotr), eps0 = 1.0;
    double beta = 0.0, kmin = 0.0, eps = eps0;
    long long j = 0;
    int xbuf = 0;
    double* Pold = P.P1;   // zeros on entry (own rows and halo rows)
    double* Pnew = P.P0;
    double* left_Pnew = P.left_P0;
    double* right_Pnew = P.right_P0;
    double* left_Pother = P.left_P1;
    double* right_Pother = P.right_P1;
    const int tg = P.tau0 + tau;                       // global slice index
    const bool wrap_c = (tg == 0);
    const bool wrap_n = (tg + 1 == P.Lglob);

    while (alive && j < maxiter) {
        ++j;
        if (boundary) {
            // the halo rows read below were pushed by the neighbour GPU before its barrier number `seq` (or earlier):
            // wait for its point-to-point flag (monotonic; the neighbour may already be one barrier ahead)
            if (threadIdx.x == 0) {
                unsigned int spins = 0;
                if (first)
                    while ((int)(ld_flag(P.my_lo_flag) - seq) < 0 && ++spins < kSpinLimit) {}
                if (last)
                    while ((int)(ld_flag(P.my_hi_flag) - seq) < 0 && ++spins < kSpinLimit) {}
                fence_gpu();   // the tiles are in this GPU's memory, ordered before the flag by the writer's fence.sys
            }
            __syncthreads();
        }
#pragma unroll
        for (int rr = 0; rr < PY; ++rr)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const size_t e = eidx(rr, q);
                const double pm = fma(beta, __ldcg(Pold + rowm + e), __ldcg(P.R + rowm + e));
                const double pcv = fma(beta, pprev.a[rr][q], r.a[rr][q]);
                pc.a[rr][q] = pcv;
                Pnew[row + e] = pcv;
                if (first) left_Pnew[e] = pcv;
                if (last) right_Pnew[e] = pcv;
                t1.a[rr][q] = Dc.a[rr][q] * pm;
                t2.a[rr][q] = Dn.a[rr][q] * pcv;
            }
        g0_x_even(t1, P.c0, P.s0);
        g0_x_even(t2, P.c0, P.s0);
        g1_x_odd(t1, P.c1, P.s1, lane);
        g1_x_odd(t2, P.c1, P.s1, lane);
        g2_y_even(t1, P.c2, P.s2);
        g2_y_even(t2, P.c2, P.s2);
        {
            double a1[NSEG], a2[NSEG], b1[NSEG], b2[NSEG];
            exchange_edges2(t1, t2, strips + (size_t)xbuf * nwarps * 4 * LX, warp, nwarps, lane, a1, a2, b1, b2);
            xbuf ^= 1;
            g3_y_odd(t1, P.c3, P.s3, a1, b1);
            g3_y_odd(t2, P.c3, P.s3, a2, b2);
        }
        double acc = 0.0;
#pragma unroll
        for (int rr = 0; rr < PY; ++rr)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const size_t e = eidx(rr, q);
                const double pn = fma(beta, __ldcg(Pold + rowp + e), __ldcg(P.R + rowp + e));
                const double wc = wrap_c ? (pc.a[rr][q] + t1.a[rr][q]) : (pc.a[rr][q] - t1.a[rr][q]);
                const double wn = wrap_n ? (pn + t2.a[rr][q]) : (pn - t2.a[rr][q]);
                t1.a[rr][q] = wc;
                t2.a[rr][q] = wn;
                acc = fma(wc, wc, acc);
            }
        {
            double ab[NSEG], be[NSEG];
            exchange_edges1(t2, strips + (size_t)xbuf * nwarps * 4 * LX, warp, nwarps, lane, ab, be);
            xbuf ^= 1;
            g3_y_odd(t2, P.c3, P.s3, ab, be);
        }
        g2_y_even(t2, P.c2, P.s2);
        g1_x_odd(t2, P.c1, P.s1, lane);
        g0_x_even(t2, P.c0, P.s0);
        double pAp;
        // the p pushes of this iteration become visible with this barrier; they are read in the NEXT iteration (as Pold)
        alive = global_sum(tile_sum<NSEG, PY>(acc, red, lane, warp, nwarps), P, ++nbar, ++seq, first, last, red, flag, pAp);
        if (!alive) break;
        const double alpha = rdotr / pAp;
        double accr = 0.0;
#pragma unroll
        for (int rr = 0; rr < PY; ++rr)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const size_t e = eidx(rr, q);
                const double du = Dn.a[rr][q] * t2.a[rr][q];
                const double z = wrap_n ? (t1.a[rr][q] + du) : (t1.a[rr][q] - du);
                x.a[rr][q] = fma(alpha, pc.a[rr][q], x.a[rr][q]);
                const double rv = fma(-alpha, z, r.a[rr][q]);
                r.a[rr][q] = rv;
                P.R[row + e] = rv;
                if (first) P.left_R[e] = rv;
                if (last) P.right_R[e] = rv;
                accr = fma(rv, rv, accr);
                pprev.a[rr][q] = pc.a[rr][q];
            }
        double rrn;
        alive = global_sum(tile_sum<NSEG, PY>(accr, red, lane, warp, nwarps), P, ++nbar, ++seq, first, last, red, flag, rrn);
        if (!alive) break;
        eps = sqrt(rrn) / normb;
        const double lg = log(2.0 * eps0 / eps);
        const double qq = 2.0 * (double)j / lg;
        const double kap = qq * qq;
        if (kap > kmin) kmin = kap;
        if (eps < tol || kmin > kappa_max) break;
        beta = rrn / rdotr;
        rdotr = rrn;
        double* tmp = Pold; Pold = Pnew; Pnew = tmp;
        tmp = left_Pnew; left_Pnew = left_Pother; left_Pother = tmp;
        tmp = right_Pnew; right_Pnew = right_Pother; right_Pother = tmp;
    }
#pragma unroll
    for (int rr = 0; rr < PY; ++rr)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) P.x[row + eidx(rr, q)] = x.a[rr][q];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        P.S->iter = j;
        P.S->eps = eps;
        P.S->normb = normb;
        P.S->kappa_min = kmin;
        P.S->done = alive ? 1 : 2;
    }
}

template <int NSEG, int PY, int MAXT>
bool launch_p2p(elph_handle* h, P2pParams& P, int nwarps) {
    constexpr int LX = 32 * NSEG;
    const size_t smem = 2ull * nwarps * 4 * LX * sizeof(double);
    auto kern = cg_p2p_kernel<NSEG, PY, MAXT>;
    elph_enable_smem(h, kern);
    int per_sm = 0;
    ELPH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nwarps * 32, smem));
    if ((long long)per_sm * h->sm_count < h->L) return false;   // all slices of the slab must be co-resident
    ELPH_CUDA(cudaMemsetAsync(h->d_bar, 0, sizeof(unsigned int), h->stream));
    void* args[] = {&P};
    ELPH_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(h->L), dim3(nwarps * 32), args, smem, h->stream));
    h->launches++;
    return true;
}

size_t arena_vec_doubles(const elph_handle* h) { return (size_t)(h->p2p.Lmax + 2) * h->N; }
size_t arena_bytes(const elph_handle* h) {
    return 3 * arena_vec_doubles(h) * sizeof(double) + 2ull * kMaxWorld * 2 * sizeof(unsigned long long) + 256;
}
// halo flags behind the mailboxes, each on its own 128-byte line: which = 0 (halo_lo ready), 1 (halo_hi ready)
unsigned int* arena_flag(const elph_handle* h, void* base, int which) {
    unsigned char* p = reinterpret_cast<unsigned char*>(base) + 3 * arena_vec_doubles(h) * sizeof(double) +
                       2ull * kMaxWorld * 2 * sizeof(unsigned long long);
    return reinterpret_cast<unsigned int*>(p + 128 * which);
}
double* arena_vec(const elph_handle* h, void* base, int which) {   // own slice 0 of R (0), P0 (1), P1 (2)
    return reinterpret_cast<double*>(base) + which * arena_vec_doubles(h) + h->N;
}
unsigned long long* arena_mbox(const elph_handle* h, void* base) {
    return reinterpret_cast<unsigned long long*>(reinterpret_cast<double*>(base) + 3 * arena_vec_doubles(h));
}

}  // namespace

// Allocate the arena (once) and export its IPC handle (64 bytes).
void elph_shard_p2p_export_impl(elph_handle* h, int rank, int world, unsigned char* handle_out) {
    ELPH_REQUIRE(h->sharded, ELPH_ERR_STATE, "elph_set_shard has not been called");
    ELPH_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, ELPH_ERR_INVALID, "bad rank / world");
    ELPH_REQUIRE(h->sq.enabled, ELPH_ERR_UNSUPPORTED, "the peer-memory CG needs the square-lattice register kernels");
    auto& A = h->p2p;
    if (!A.arena) {
        A.rank = rank;
        A.world = world;
        A.Lmax = (h->shard_Lglob + world - 1) / world;
        ELPH_REQUIRE(h->L <= A.Lmax, ELPH_ERR_INVALID, "slab longer than ceil(Lglob / world)");
        ELPH_CUDA(cudaMalloc(&A.arena, arena_bytes(h)));
        ELPH_CUDA(cudaMemset(A.arena, 0, arena_bytes(h)));
        ELPH_CUDA(cudaDeviceSynchronize());
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t ipc;
    ELPH_CUDA(cudaIpcGetMemHandle(&ipc, A.arena));
    memcpy(handle_out, &ipc, sizeof(ipc));
}

// Open the arenas of all ranks (handles: world x 64 bytes, rank order).
void elph_shard_p2p_open_impl(elph_handle* h, const unsigned char* handles, const int64_t* slab_lengths) {
    auto& A = h->p2p;
    ELPH_REQUIRE(A.arena, ELPH_ERR_STATE, "elph_shard_p2p_export must be called first");
    ELPH_REQUIRE(handles && slab_lengths, ELPH_ERR_INVALID, "null argument");
    A.peer.assign(A.world, nullptr);
    A.peer_L.assign(slab_lengths, slab_lengths + A.world);
    ELPH_REQUIRE(A.peer_L[A.rank] == h->L, ELPH_ERR_INVALID, "slab_lengths[rank] differs from this handle's Ltau");
    for (int q = 0; q < A.world; ++q) ELPH_REQUIRE(A.peer_L[q] >= 1 && A.peer_L[q] <= A.Lmax, ELPH_ERR_INVALID, "bad slab length");
    for (int q = 0; q < A.world; ++q) {
        if (q == A.rank) {
            A.peer[q] = A.arena;
            continue;
        }
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, handles + (size_t)q * sizeof(ipc), sizeof(ipc));
        ELPH_CUDA(cudaIpcOpenMemHandle(&A.peer[q], ipc, cudaIpcMemLazyEnablePeerAccess));
    }
    A.opened = true;
}

void elph_shard_p2p_close_impl(elph_handle* h) {
    auto& A = h->p2p;
    for (int q = 0; q < (int)A.peer.size(); ++q)
        if (A.peer[q] && q != A.rank) cudaIpcCloseMemHandle(A.peer[q]);
    A.peer.clear();
    if (A.arena) cudaFree(A.arena);
    A.arena = nullptr;
    A.opened = false;
}

// Solve A x = b with x0 = 0 on the slab owned by this rank; every rank of the ring must make the same call.
// b_own / x_own: [L][N] own slices (engine layout).  Returns false if the kernel does not apply (caller falls back).
bool elph_shard_cg_p2p_impl(elph_handle* h, const double* b_own, double* x_own, double tol, int64_t maxiter, int64_t* iters,
                            double* eps) {
    auto& A = h->p2p;
    ELPH_REQUIRE(A.opened, ELPH_ERR_STATE, "elph_shard_p2p_open has not been called");
    if (tol == 0.0) tol = h->cg_tol;
    if (maxiter == 0) maxiter = h->cg_maxiter;
    const int Lx = h->sq.Lx, Ly = h->sq.Ly;
    const int PY = (Lx == 32) ? 8 : 4;
    if (Ly % PY) return false;
    const int nwarps = Ly / PY;
    if (nwarps < 2 || nwarps > 32 || h->partial_cap < h->L) return false;
    cudaStream_t st = h->stream;
    const int left = (A.rank + A.world - 1) % A.world, right = (A.rank + 1) % A.world;
    P2pParams P;
    P.D = h->d_D; P.b = b_own; P.x = x_own;
    P.R = arena_vec(h, A.arena, 0); P.P0 = arena_vec(h, A.arena, 1); P.P1 = arena_vec(h, A.arena, 2);
    // left neighbour's halo_hi row sits right after ITS last own slice; right neighbour's halo_lo row is its row -1
    const int L_left = (int)A.peer_L[left];
    P.left_R = arena_vec(h, A.peer[left], 0) + (size_t)L_left * h->N;
    P.left_P0 = arena_vec(h, A.peer[left], 1) + (size_t)L_left * h->N;
    P.left_P1 = arena_vec(h, A.peer[left], 2) + (size_t)L_left * h->N;
    P.right_R = arena_vec(h, A.peer[right], 0) - h->N;
    P.right_P0 = arena_vec(h, A.peer[right], 1) - h->N;
    P.right_P1 = arena_vec(h, A.peer[right], 2) - h->N;
    for (int q = 0; q < kMaxWorld; ++q) P.mbox[q] = (q < A.world) ? arena_mbox(h, A.peer[q]) : nullptr;
    P.left_hi_flag = arena_flag(h, A.peer[left], 1);
    P.right_lo_flag = arena_flag(h, A.peer[right], 0);
    P.my_lo_flag = arena_flag(h, A.arena, 0);
    P.my_hi_flag = arena_flag(h, A.arena, 1);
    P.partial = h->d_partial; P.bar = h->d_bar; P.S = h->d_cg;
    P.seq_base = A.seq;
    P.L = h->L; P.Ly = Ly; P.rank = A.rank; P.world = A.world; P.tau0 = h->shard_tau0; P.Lglob = h->shard_Lglob;
    P.c0 = h->sq.c[0]; P.s0 = h->sq.s[0]; P.c1 = h->sq.c[1]; P.s1 = h->sq.s[1];
    P.c2 = h->sq.c[2]; P.s2 = h->sq.s[2]; P.c3 = h->sq.c[3]; P.s3 = h->sq.s[3];
    // p_old of the first iteration (own rows and halo rows).  Peers push into P halos only after the first barrier of
    // their kernel, which needs this rank's kernel to be running, i.e. this memset to be complete.
    ELPH_CUDA(cudaMemsetAsync(P.P1 - h->N, 0, arena_vec_doubles(h) * sizeof(double), st));
    CgScalars init = {};
    init.tol = tol; init.kappa_max = h->cg_kappa_max; init.maxiter = maxiter;
    *h->h_cg = init;
    ELPH_CUDA(cudaMemcpyAsync(h->d_cg, h->h_cg, sizeof(CgScalars), cudaMemcpyHostToDevice, st));
    bool ok = false;
    if (Lx == 32 && PY == 8 && nwarps * 32 <= 256) ok = launch_p2p<1, 8, 256>(h, P, nwarps);
    else if (Lx == 64 && PY == 4 && nwarps * 32 <= 512) ok = launch_p2p<2, 4, 512>(h, P, nwarps);
    if (!ok) return false;
    ELPH_CUDA(cudaMemcpyAsync(h->h_cg, h->d_cg, sizeof(CgScalars), cudaMemcpyDeviceToHost, st));
    ELPH_CUDA(cudaStreamSynchronize(st));
    ELPH_REQUIRE(h->h_cg->done == 1, ELPH_ERR_STATE, "peer-memory CG: a peer GPU did not reach the barrier (timeout)");
    A.seq += 1u + 2u * (unsigned int)h->h_cg->iter;   // barriers executed: the initial one + two per iteration
    if (iters) *iters = h->h_cg->iter;
    if (eps) *eps = h->h_cg->eps;
    return true;
}
