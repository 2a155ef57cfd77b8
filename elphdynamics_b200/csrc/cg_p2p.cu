// Multi-GPU conjugate gradient on A = M^T M for ONE tau-sharded lattice, fused with its collectives over peer memory.
//
// rank g owns a contiguous slab of time slices.  A CG iteration needs (1) the neighbouring slices of p
// across the slab boundary and (2) two scalar all-reduces (p.Ap and |r|^2).  With NCCL calls between kernel launches
// that is ~77 us per product at config E; here the whole solve is ONE cooperative persistent kernel per GPU (the loop
// of cg_persistent.cu, src/IterativeSolvers.jl:239-314) and both collectives happen inside it over NVLink peer memory:
//
//   * halo: the CTAs of the first / last slice of a slab PUSH their new r and p tiles straight into the neighbour GPU's
//     halo rows as SELF-VALIDATING words (the "LL" idea of NCCL): every double travels as one 16-byte store of two
//     64-bit words, each carrying half of the double and a 32-bit tag that names the iteration that produced it.  The
//     reader spins on the element itself until both tags match -- no flag, no system-scope fence, and the posted stores
//     overlap the barrier that follows them.  (First version: plain rows + fence.acq_rel.sys + a flag per side; the
//     fence sat on the critical path of the boundary CTAs: 13.0 us/iteration at config B against 7 us for the
//     single-GPU persistent kernel.)
//   * all-reduce + barrier: on each GPU the CTAs arrive on a counter and every CTA reads back the GPU's partials in
//     index order (the barrier of cg_persistent.cu); CTA 0 then writes {sum, sequence number} into a mailbox slot on
//     every OTHER GPU (same two-word encoding) and each CTA polls its own GPU's mailbox until the world-1 remote slots
//     carry the sequence number; the total is summed in rank order -- the same bits on every GPU, so all GPUs take the
//     same branch.  With world = 1 this is exactly the single-GPU barrier.
//
// Memory: each process allocates one arena (cudaMalloc), exports it with cudaIpcGetMemHandle and opens the others'
// (elph_shard_p2p_*).  Arena = R, P0, P1 as [Lmax slices] (Lmax = ceil(Lglob / world), same layout on every rank), six
// tagged halo rows (R, P0, P1 x lo, hi; 16 bytes per site), the mailboxes and the partials.  Sequence numbers / tags
// increase monotonically over the life of the handle (all ranks execute the same number of barriers), mailbox slots
// and partials alternate by parity, nothing is ever reset.
// Holstein on periodic square lattices (the register tiles of mtm_square.cu); the reference has no counterpart.
#include "square_tiles.cuh"

#include <algorithm>
#include <cstring>

namespace {

using namespace sqt;

constexpr int kMaxWorld = 16;
constexpr unsigned int kSpinLimit = 1u << 25;   // ~ 20 s: a dead peer ends the solve with an error instead of a hang

struct P2pParams {
    const double* __restrict__ D;   // expnV with halos: slice index -1 .. L valid
    const double* __restrict__ b;   // [L][N] right-hand side (initial guess is zero)
    double* __restrict__ x;         // [L][N] out
    double* R;                      // [L][N] own slices in the arena
    double* P0;                     // p buffers: P0 and P0 + Lmax * N, alternating by iteration parity
    // tagged halo rows, [N][2] 64-bit words each, six per arena in the order R lo, R hi, P0 lo, P0 hi, P1 lo, P1 hi:
    // own arena (written by the neighbours) and the two neighbours' arenas (peer memory, written by this GPU: the hi rows
    // of the left neighbour, the lo rows of the right neighbour)
    const unsigned long long* my_halo;
    unsigned long long* left_halo;
    unsigned long long* right_halo;
    unsigned long long* mbox[kMaxWorld];   // mailbox base of every rank (own included): [2 parities][world][2 words]
    double* partial;                // [2 parities][Lmax] per-CTA partials of this GPU
    unsigned int* bar;              // arrival counter of this GPU (monotonic over the launch, zeroed by the host)
    CgScalars* S;                   // in: tol, kappa_max, maxiter; out: iter, eps, normb, done (2 = peer timeout)
    unsigned int seq_base;          // sequence number of the last barrier of the previous solve
    int L, Lmax, Ly, rank, world, tau0, Lglob;
    double c0, s0, c1, s1, c2, s2, c3, s3;
};

__device__ __forceinline__ void st_ll(unsigned long long* p, unsigned long long w0, unsigned long long w1) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ void ld_ll(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
// one double + tag as two self-validating words (a torn 16-byte store cannot pass for a complete one)
__device__ __forceinline__ void push_ll(unsigned long long* p, double v, unsigned int tag) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    st_ll(p, (bits & 0xffffffffull) | ((unsigned long long)tag << 32), (bits >> 32) | ((unsigned long long)tag << 32));
}
__device__ __forceinline__ double unpack_ll(unsigned long long a, unsigned long long b) {
    return __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Read one tagged halo tile (PY x NSEG elements per thread): all loads are issued before any tag is checked, so a tile
// that has already arrived costs one L2 round trip.  Returns 0 on a timeout.
template <int NSEG, int PY, typename Idx>
__device__ __forceinline__ int read_halo(const unsigned long long* row, unsigned int tag, Idx eidx, double (&out)[PY][NSEG]) {
    unsigned long long wa[PY][NSEG], wb[PY][NSEG];
    unsigned int spins = 0;
    bool all;
    do {
        all = true;
#pragma unroll
        for (int rr = 0; rr < PY; ++rr)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) ld_ll(row + 2 * eidx(rr, q), wa[rr][q], wb[rr][q]);
#pragma unroll
        for (int rr = 0; rr < PY; ++rr)
#pragma unroll
            for (int q = 0; q < NSEG; ++q)
                all = all && ((unsigned int)(wa[rr][q] >> 32) == tag) && ((unsigned int)(wb[rr][q] >> 32) == tag);
    } while (!all && ++spins < kSpinLimit);
#pragma unroll
    for (int rr = 0; rr < PY; ++rr)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) out[rr][q] = unpack_ll(wa[rr][q], wb[rr][q]);
    return all ? 1 : 0;
}

// Barrier over all CTAs of all GPUs fused with the sum of one double per CTA.  `nbar` counts this GPU's barriers of the
// launch from 1 (local arrival target = nbar * gridDim.x), seq is the global sequence number.  Returns false on a
// timeout.  The caller must have a __syncthreads between the CTA's global writes and this call.  The all-reduce carries
// no cross-GPU ordering obligation: remote data is only ever read through the self-validating halo rows.
__device__ __forceinline__ bool global_sum(double block_value, const P2pParams& P, unsigned int nbar, unsigned int seq, double* red,
                                           int* flag, double& out) {
    const int nb = gridDim.x;
    double* partial = P.partial + (size_t)(seq & 1u) * P.Lmax;
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = block_value;
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(P.bar), "r"(1u) : "memory");
        const unsigned int target = nbar * (unsigned int)nb;
        unsigned int spins = 0;
        int ok = 1;
        while (ld_acquire_gpu(P.bar) < target)
            if (++spins > kSpinLimit) { ok = 0; break; }
        flag[0] = ok;
    }
    __syncthreads();
    // every CTA folds this GPU's partials in the same fixed order (thread-strided, shuffle tree, warps in order)
    double s = 0.0;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) s += __ldcg(partial + k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
    bool good = (flag[0] != 0);
    if (P.world > 1) {
        __syncthreads();   // red is reused below
        if (threadIdx.x == 0) {
            const size_t par = (size_t)(seq & 1u) * P.world * 2;
            if (blockIdx.x == 0)
                for (int q = 1; q < P.world; ++q) push_ll(P.mbox[(P.rank + q) % P.world] + par + 2 * P.rank, t, seq);
            const unsigned long long* mine = P.mbox[P.rank] + par;
            unsigned long long wa[kMaxWorld], wb[kMaxWorld];
            unsigned int spins = 0;
            bool all;
            do {
                all = true;
                for (int g = 0; g < P.world; ++g)
                    if (g != P.rank) ld_ll(mine + 2 * g, wa[g], wb[g]);
                for (int g = 0; g < P.world; ++g)
                    if (g != P.rank) all = all && ((unsigned int)(wa[g] >> 32) == seq) && ((unsigned int)(wb[g] >> 32) == seq);
            } while (!all && ++spins < kSpinLimit);
            double tot = 0.0;
            for (int g = 0; g < P.world; ++g) tot += (g == P.rank) ? t : unpack_ll(wa[g], wb[g]);   // rank order: same bits everywhere
            red[0] = tot;
            flag[1] = all ? 1 : 0;
        }
        __syncthreads();
        t = red[0];
        good = good && (flag[1] != 0);
    }
    out = t;
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
    // 128-register cap at 512 threads: x, D(tau), D(tau+1) live in shared memory there (one CTA per SM anyway)
    constexpr bool XS = (MAXT > 256);
    extern __shared__ __align__(16) double strips[];   // 2 x [nwarps][4][LX]; XS: + x, D(tau), D(tau+1) [3][N]
    __shared__ double red[32];
    __shared__ int flag[2];   // barrier completed without a timeout: [0] local counter, [1] peers' mailbox words
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int L = P.L, N = LX * P.Ly;
    const int tau = blockIdx.x;                      // local slice; tau-1 = -1 and tau+1 = L live in the tagged halo rows
    const bool first = (tau == 0), last = (tau == L - 1);
    const size_t tile_off = (size_t)warp * PY * LX;
    auto eidx = [&](int r, int q) -> size_t { return tile_off + r * LX + 32 * q + lane; };
    const long long row = (long long)tau * N, rowm = row - N, rowp = row + N;
    // tags: r written in iteration j (0 = the initial residual) carries base + 2j + 1, p written in iteration j >= 1
    // carries base + 2j; unique over the life of the handle because the host advances base by 2 * iterations + 2
    const unsigned int base = P.seq_base;
    const size_t vstride = (size_t)P.Lmax * N, hrow = 2 * (size_t)N;
    auto halo = [&](const unsigned long long* b, int vec, int side) { return b + (size_t)(2 * vec + side) * hrow; };
    auto halo_w = [&](unsigned long long* b, int vec, int side) { return b + (size_t)(2 * vec + side) * hrow; };

    Tile<NSEG, PY> r, pc, t1, t2;   // pc: p_{j-1} on entry to iteration j, p_j after its first loop
    Tile<NSEG, PY> xr, Dcr, Dnr;   // registers when !XS (dead otherwise)
    double* xs = strips + 2ull * nwarps * 4 * LX;
    double* dcs = xs + N;
    double* dns = dcs + N;
    auto sidx = [&](int rr, int q) -> int { return (rr * NSEG + q) * (int)blockDim.x + (int)threadIdx.x; };
    auto X = [&](int rr, int q) -> double& { if constexpr (XS) return xs[sidx(rr, q)]; else return xr.a[rr][q]; };
    auto DC = [&](int rr, int q) -> double& { if constexpr (XS) return dcs[sidx(rr, q)]; else return Dcr.a[rr][q]; };
    auto DN = [&](int rr, int q) -> double& { if constexpr (XS) return dns[sidx(rr, q)]; else return Dnr.a[rr][q]; };
    double accb = 0.0;
#pragma unroll
    for (int rr = 0; rr < PY; ++rr)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const size_t e = eidx(rr, q);
            const double bv = P.b[row + e];          // x0 = 0: r0 = b
            X(rr, q) = 0.0;
            r.a[rr][q] = bv;
            pc.a[rr][q] = 0.0;
            DC(rr, q) = P.D[row + e];
            DN(rr, q) = P.D[rowp + e];
            P.R[row + e] = bv;
            if (first) push_ll(halo_w(P.left_halo, 0, 1) + 2 * e, bv, base + 1u);   // r0 into the neighbours' halo rows
            if (last) push_ll(halo_w(P.right_halo, 0, 0) + 2 * e, bv, base + 1u);
            accb = fma(bv, bv, accb);
        }
    const double tol = P.S->tol, kappa_max = P.S->kappa_max;
    const long long maxiter = P.S->maxiter;
    unsigned int nbar = 0, seq = base;
    double rdotr;
    bool alive = global_sum(tile_sum<NSEG, PY>(accb, red, lane, warp, nwarps), P, ++nbar, ++seq, red, flag, rdotr);
    const double normb = sqrt(rdotr), eps0 = 1.0;
    double beta = 0.0, kmin = 0.0, eps = eps0;
    long long j = 0;
    int xbuf = 0;
    int pb = 0;   // p_j goes to buffer pb, p_{j-1} is in buffer pb ^ 1 (never read in the first iteration: beta = 0)
    const int tg = P.tau0 + tau;                       // global slice index
    const bool wrap_c = (tg == 0);
    const bool wrap_n = (tg + 1 == P.Lglob);

    while (alive && j < maxiter) {
        ++j;
        const unsigned int tag_r_prev = base + 2u * (unsigned int)(j - 1) + 1u;   // r_{j-1}
        const unsigned int tag_p_prev = base + 2u * (unsigned int)(j - 1);        // p_{j-1} (j > 1)
        const unsigned int tag_p = base + 2u * (unsigned int)j;
        const unsigned int tag_r = tag_p + 1u;
        int halo_ok = 1;
        double* Pnew = P.P0 + (pb ? vstride : 0);
        const double* Pold = P.P0 + (pb ? 0 : vstride);
        // p_j(tau-1) = r_{j-1}(tau-1) + beta p_{j-1}(tau-1): neighbour rows of this GPU, or the tagged halo row
        if (first) {
            double hr[PY][NSEG];
            halo_ok &= read_halo<NSEG, PY>(halo(P.my_halo, 0, 0), tag_r_prev, eidx, hr);
#pragma unroll
            for (int rr = 0; rr < PY; ++rr)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) t1.a[rr][q] = hr[rr][q];
            if (j > 1) {
                halo_ok &= read_halo<NSEG, PY>(halo(P.my_halo, 1 + (pb ^ 1), 0), tag_p_prev, eidx, hr);
#pragma unroll
                for (int rr = 0; rr < PY; ++rr)
#pragma unroll
                    for (int q = 0; q < NSEG; ++q) t1.a[rr][q] = fma(beta, hr[rr][q], t1.a[rr][q]);
            }
        } else {
#pragma unroll
            for (int rr = 0; rr < PY; ++rr)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) {
                    const size_t e = eidx(rr, q);
                    const double rm = __ldcg(P.R + rowm + e);
                    t1.a[rr][q] = (j > 1) ? fma(beta, __ldcg(Pold + rowm + e), rm) : rm;
                }
        }
#pragma unroll
        for (int rr = 0; rr < PY; ++rr)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const size_t e = eidx(rr, q);
                const double pcv = fma(beta, pc.a[rr][q], r.a[rr][q]);
                pc.a[rr][q] = pcv;
                Pnew[row + e] = pcv;
                if (first) push_ll(halo_w(P.left_halo, 1 + pb, 1) + 2 * e, pcv, tag_p);
                if (last) push_ll(halo_w(P.right_halo, 1 + pb, 0) + 2 * e, pcv, tag_p);
                t1.a[rr][q] = DC(rr, q) * t1.a[rr][q];
                t2.a[rr][q] = DN(rr, q) * pcv;
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
        // p_j(tau+1), same rule
        double pn[PY][NSEG];
        if (last) {
            halo_ok &= read_halo<NSEG, PY>(halo(P.my_halo, 0, 1), tag_r_prev, eidx, pn);
            if (j > 1) {
                double hp[PY][NSEG];
                halo_ok &= read_halo<NSEG, PY>(halo(P.my_halo, 1 + (pb ^ 1), 1), tag_p_prev, eidx, hp);
#pragma unroll
                for (int rr = 0; rr < PY; ++rr)
#pragma unroll
                    for (int q = 0; q < NSEG; ++q) pn[rr][q] = fma(beta, hp[rr][q], pn[rr][q]);
            }
        } else {
#pragma unroll
            for (int rr = 0; rr < PY; ++rr)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) {
                    const size_t e = eidx(rr, q);
                    const double rp = __ldcg(P.R + rowp + e);
                    pn[rr][q] = (j > 1) ? fma(beta, __ldcg(Pold + rowp + e), rp) : rp;
                }
        }
        double acc = 0.0;
#pragma unroll
        for (int rr = 0; rr < PY; ++rr)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const double wc = wrap_c ? (pc.a[rr][q] + t1.a[rr][q]) : (pc.a[rr][q] - t1.a[rr][q]);
                const double wn = wrap_n ? (pn[rr][q] + t2.a[rr][q]) : (pn[rr][q] - t2.a[rr][q]);
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
        if (!__syncthreads_and(halo_ok)) { alive = false; break; }   // a neighbour's tile never arrived
        double pAp;
        alive = global_sum(tile_sum<NSEG, PY>(acc, red, lane, warp, nwarps), P, ++nbar, ++seq, red, flag, pAp);
        if (!alive) break;
        const double alpha = rdotr / pAp;
        double accr = 0.0;
#pragma unroll
        for (int rr = 0; rr < PY; ++rr)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const size_t e = eidx(rr, q);
                const double du = DN(rr, q) * t2.a[rr][q];
                const double z = wrap_n ? (t1.a[rr][q] + du) : (t1.a[rr][q] - du);
                X(rr, q) = fma(alpha, pc.a[rr][q], X(rr, q));
                const double rv = fma(-alpha, z, r.a[rr][q]);
                r.a[rr][q] = rv;
                P.R[row + e] = rv;
                if (first) push_ll(halo_w(P.left_halo, 0, 1) + 2 * e, rv, tag_r);
                if (last) push_ll(halo_w(P.right_halo, 0, 0) + 2 * e, rv, tag_r);
                accr = fma(rv, rv, accr);
            }
        double rrn;
        alive = global_sum(tile_sum<NSEG, PY>(accr, red, lane, warp, nwarps), P, ++nbar, ++seq, red, flag, rrn);
        if (!alive) break;
        eps = sqrt(rrn) / normb;
        const double lg = log(2.0 * eps0 / eps);
        const double qq = 2.0 * (double)j / lg;
        const double kap = qq * qq;
        if (kap > kmin) kmin = kap;
        if (eps < tol || kmin > kappa_max) break;
        beta = rrn / rdotr;
        rdotr = rrn;
        pb ^= 1;
    }
#pragma unroll
    for (int rr = 0; rr < PY; ++rr)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) P.x[row + eidx(rr, q)] = X(rr, q);
    if (!alive && threadIdx.x == 0) P.S->done = 2;   // any CTA that saw a timeout marks the solve as failed
    if (alive && blockIdx.x == 0 && threadIdx.x == 0) {
        P.S->iter = j;
        P.S->eps = eps;
        P.S->normb = normb;
        P.S->kappa_min = kmin;
        P.S->done = 1;
    }
}

template <int NSEG, int PY, int MAXT>
size_t p2p_smem(const elph_handle* h, int nwarps) {
    return (2ull * nwarps * 4 * (32 * NSEG) + (MAXT > 256 ? 3ull * h->N : 0)) * sizeof(double);
}

// all slices of the slab must be co-resident (cooperative launch, one CTA per slice)
template <int NSEG, int PY, int MAXT>
bool fits_p2p(elph_handle* h, int nwarps) {
    auto kern = cg_p2p_kernel<NSEG, PY, MAXT>;
    elph_enable_smem(h, kern);
    int per_sm = 0;
    ELPH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nwarps * 32, p2p_smem<NSEG, PY, MAXT>(h, nwarps)));
    return (long long)per_sm * h->sm_count >= h->L;
}

template <int NSEG, int PY, int MAXT>
bool launch_p2p(elph_handle* h, P2pParams& P, int nwarps) {
    if (!fits_p2p<NSEG, PY, MAXT>(h, nwarps)) return false;
    auto kern = cg_p2p_kernel<NSEG, PY, MAXT>;
    ELPH_CUDA(cudaMemsetAsync(h->d_bar, 0, sizeof(unsigned int), h->stream));
    void* args[] = {&P};
    ELPH_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(h->L), dim3(nwarps * 32), args, p2p_smem<NSEG, PY, MAXT>(h, nwarps),
                                          h->stream));
    h->launches++;
    return true;
}

// kernel variant for this handle: 0 = none applies
int p2p_variant(const elph_handle* h, int& nwarps) {
    const int Lx = h->sq.Lx, Ly = h->sq.Ly;
    const int PY = (Lx == 32) ? 8 : 4;
    if (Ly % PY) return 0;
    nwarps = Ly / PY;
    if (nwarps < 2 || nwarps > 32) return 0;
    if (Lx == 32 && nwarps * 32 <= 256) return 1;
    if (Lx == 64 && nwarps * 32 <= 512) return 2;
    return 0;
}

// ---- arena layout (identical on every rank) ------------------------------------------------------------------------
size_t arena_vec_doubles(const elph_handle* h) { return (size_t)h->p2p.Lmax * h->N; }
size_t arena_halo_words(const elph_handle* h) { return 2 * (size_t)h->N; }   // one tagged row
size_t arena_mbox_words() { return 2ull * kMaxWorld * 2; }
size_t arena_bytes(const elph_handle* h) {
    return 3 * arena_vec_doubles(h) * sizeof(double) + 6 * arena_halo_words(h) * sizeof(unsigned long long) +
           arena_mbox_words() * sizeof(unsigned long long) + 2 * (size_t)h->p2p.Lmax * sizeof(double) + 256;
}
double* arena_vec(const elph_handle* h, void* base, int which) {   // R (0), P0 (1), P1 (2)
    return reinterpret_cast<double*>(base) + which * arena_vec_doubles(h);
}
// tagged halo rows: vec = R (0), P0 (1), P1 (2); side = 0 (lo: slice -1), 1 (hi: slice L)
unsigned long long* arena_halo(const elph_handle* h, void* base, int vec, int side) {
    return reinterpret_cast<unsigned long long*>(reinterpret_cast<double*>(base) + 3 * arena_vec_doubles(h)) +
           (size_t)(2 * vec + side) * arena_halo_words(h);
}
unsigned long long* arena_mbox(const elph_handle* h, void* base) { return arena_halo(h, base, 3, 0); }
double* arena_partial(const elph_handle* h, void* base) { return reinterpret_cast<double*>(arena_mbox(h, base) + arena_mbox_words()); }

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
    int nwarps = 0;
    const int variant = p2p_variant(h, nwarps);
    const bool fits = (variant == 1) ? fits_p2p<1, 8, 256>(h, nwarps) : (variant == 2) ? fits_p2p<2, 4, 512>(h, nwarps) : false;
    ELPH_REQUIRE(fits, ELPH_ERR_UNSUPPORTED,
                 "peer-memory CG: the slab's time slices are not all co-resident on this GPU (or unsupported lattice)");
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
    int nwarps = 0;
    const int variant = p2p_variant(h, nwarps);
    if (!variant) return false;
    const int Ly = h->sq.Ly;
    cudaStream_t st = h->stream;
    const int left = (A.rank + A.world - 1) % A.world, right = (A.rank + 1) % A.world;
    P2pParams P;
    P.D = h->d_D; P.b = b_own; P.x = x_own;
    P.R = arena_vec(h, A.arena, 0); P.P0 = arena_vec(h, A.arena, 1);
    P.my_halo = arena_halo(h, A.arena, 0, 0);
    P.left_halo = arena_halo(h, A.peer[left], 0, 0);    // this GPU's first slice is the left neighbour's slice L (hi rows)
    P.right_halo = arena_halo(h, A.peer[right], 0, 0);  // its last slice is the right neighbour's slice -1 (lo rows)
    for (int q = 0; q < kMaxWorld; ++q) P.mbox[q] = (q < A.world) ? arena_mbox(h, A.peer[q]) : nullptr;
    P.partial = arena_partial(h, A.arena); P.bar = h->d_bar; P.S = h->d_cg;
    P.seq_base = A.seq;
    P.L = h->L; P.Lmax = A.Lmax; P.Ly = Ly; P.rank = A.rank; P.world = A.world; P.tau0 = h->shard_tau0; P.Lglob = h->shard_Lglob;
    P.c0 = h->sq.c[0]; P.s0 = h->sq.s[0]; P.c1 = h->sq.c[1]; P.s1 = h->sq.s[1];
    P.c2 = h->sq.c[2]; P.s2 = h->sq.s[2]; P.c3 = h->sq.c[3]; P.s3 = h->sq.s[3];
    CgScalars init = {};
    init.tol = tol; init.kappa_max = h->cg_kappa_max; init.maxiter = maxiter;
    *h->h_cg = init;
    ELPH_CUDA(cudaMemcpyAsync(h->d_cg, h->h_cg, sizeof(CgScalars), cudaMemcpyHostToDevice, st));
    bool ok = false;
    if (variant == 1) ok = launch_p2p<1, 8, 256>(h, P, nwarps);
    else if (variant == 2) ok = launch_p2p<2, 4, 512>(h, P, nwarps);
    if (!ok) return false;
    ELPH_CUDA(cudaMemcpyAsync(h->h_cg, h->d_cg, sizeof(CgScalars), cudaMemcpyDeviceToHost, st));
    ELPH_CUDA(cudaStreamSynchronize(st));
    ELPH_REQUIRE(h->h_cg->done == 1, ELPH_ERR_STATE, "peer-memory CG: a peer GPU did not reach a barrier or deliver a halo tile (timeout)");
    A.seq += 2u + 2u * (unsigned int)h->h_cg->iter;   // barriers executed: 1 + 2 per iteration; tags used: up to base + 2 iter + 1
    if (iters) *iters = h->h_cg->iter;
    if (eps) *eps = h->h_cg->eps;
    return true;
}
