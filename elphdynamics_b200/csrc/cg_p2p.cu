// Single-reduction conjugate gradient on A = M^T M as ONE persistent cooperative kernel per GPU, for one lattice on one
// GPU or tau-sharded over several GPUs with the collectives fused into the kernel over NVLink peer memory.
//
// rank g owns a contiguous slab of time slices.  A CG iteration needs the neighbouring slices across the
// slab boundary and the all-reduce of its scalars.  With NCCL calls between kernel launches that is ~250 us per
// iteration at config E; a grid barrier costs ~3 us on one GPU and ~5 us across GPUs while the arithmetic of an
// iteration is ~1.5 us.  So the loop of src/IterativeSolvers.jl:239-314 is reorganised to need ONE barrier per iteration
// (Chronopoulos & Gear 1989: the same Krylov iterates, alpha/beta from recurrences):
//
//     gamma_k = (r_k, r_k),  delta_k = (r_k, A r_k) = |M r_k|^2        -> one all-reduce of two doubles
//     beta_k = gamma_k / gamma_{k-1},  alpha_k = gamma_k / (delta_k - beta_k gamma_k / alpha_{k-1})
//     p_k = r_k + beta_k p_{k-1};  s_k = w_k + beta_k s_{k-1}  (= A p_k);  x += alpha_k p_k;  r_{k+1} = r_k - alpha_k s_k
//     w_{k+1} = A r_{k+1}
//
// A is a 3-point stencil in tau, so CTA tau (one time slice, state in registers for the whole solve) needs r_{k+1} of the
// slices tau-1 and tau+1 -- and rebuilds them itself from the neighbours' r_k, w_k, s_{k-1} (written before the last
// barrier) and the scalars alpha_k, beta_k that every CTA holds: 6 neighbour rows read per iteration instead of a
// second barrier.  Stop rule, iteration numbering and the kappa bound are those of the reference; the iterates agree
// with the two-reduction form to rounding (iteration counts within +-2 in all parity tests; 948 vs 948 at config B).
//
// Multi-GPU (world > 1), both collectives inside the kernel:
//   * halo: the CTAs of the first / last slice of a slab PUSH their new w tile straight into the neighbour GPU's halo
//     row (only w crosses NVLink: the receiving CTA keeps its own copies of the neighbour's r and s rows and advances
//     them with the owner's exact operations -- 3x fewer bytes than pushing r, w, s; measured at 64x64, 2 GPUs:
//     16.4 -> see DESIGN.md us/iteration) as SELF-VALIDATING words (the "LL" idea of NCCL): every double travels as one 16-byte store of two
//     64-bit words, each carrying half of the double and a 32-bit tag that names the iteration that produced it.  The
//     reader spins on the element itself until both tags match -- no flag, no system-scope fence, and the posted stores
//     overlap the barrier that follows them.  (First version: plain rows + fence.acq_rel.sys + a flag per side; the
//     fence sat on the critical path of the boundary CTAs: 13.0 us/iteration at config B.)
//   * all-reduce + barrier: on each GPU the CTAs arrive on a counter and every CTA reads back the GPU's partials in
//     index order (the barrier of cg_persistent.cu); CTA 0 then writes {sums, sequence number} into a mailbox slot on
//     every OTHER GPU (same encoding) and each CTA polls its own GPU's mailbox until the world-1 remote slots carry the
//     sequence number; the total is summed in rank order -- the same bits on every GPU, so all GPUs take the same
//     branch.  With world = 1 the ring closes inside the GPU (plain loads of the wrapped rows, no mailbox).
//
// Memory: one arena per handle (cudaMalloc; exported with cudaIpcGetMemHandle and opened by the other ranks through
// elph_shard_p2p_*).  Arena = r, w, s double-buffered by iteration parity ([2][3][Lmax slices], Lmax = ceil(Lglob /
// world), same layout on every rank), tagged halo rows (2 parities x lo, hi for w, plus the right-hand-side rows of the
// set-up; 16 bytes per site), the boundary CTAs' private r / s copies of the neighbour rows,
// the mailboxes and the partials.  Sequence numbers / tags increase monotonically over the life of the handle (all
// ranks execute the same number of barriers), nothing is ever reset.
// Holstein on periodic square lattices (the register tiles of mtm_square.cu), any world size; SSH on 32-wide square
// lattices on one GPU (tables of slices tau, tau+1 resident in shared memory).  The reference has no counterpart.
#include "ll_words.cuh"
#include "square_tiles.cuh"

#include <algorithm>
#include <cstring>

namespace {

using namespace sqt;

constexpr int kMaxWorld = 16;
constexpr unsigned int kSpinLimit = 1u << 25;   // ~ 20 s: a dead peer ends the solve with an error instead of a hang

struct P2pParams {
    const double* __restrict__ D;   // Holstein expnV: sharded handle slices -1 .. L valid (halos), else [L][N] and tau
                                    // wraps; SSH: exp(dtau mu) [N]
    const double2* __restrict__ tab; // SSH: (cosh, sinh) [L][2][N] in the tile layout of ssh_square.cu (else unused)
    const double* __restrict__ r0;  // [L][N] initial residual (= b when the initial guess is zero)
    double* __restrict__ x;         // [L][N] in (if x0_given) / out
    double* V;                      // arena vectors [2 parities][3: r, w, s][Lmax][N]
    // tagged halo rows, [N][2] 64-bit words each, [2 parities][3 vectors][2 sides: lo, hi] per arena: own arena (written
    // by the neighbours) and the two neighbours' arenas (peer memory, written by this GPU: the hi rows of the left
    // neighbour, the lo rows of the right neighbour)
    const unsigned long long* my_halo;
    unsigned long long* left_halo;
    unsigned long long* right_halo;
    unsigned long long* mbox[kMaxWorld];   // mailbox base of every rank (own included): [2 parities][world][4 words]
    double* partial;                // [2 parities][2 values][Lmax] per-CTA partials of this GPU
    double* ghost;                  // [2 sides][2: r, s][N] the boundary CTAs' own copies of the neighbour GPU's r and s rows
    unsigned int* bar;              // arrival counter of this GPU (monotonic over the launch, zeroed by the host)
    CgScalars* S;                   // in: tol, kappa_max, maxiter, normb (0: x0 = 0, |b| = |r0|); out: iter, eps, done
    unsigned int seq_base;          // sequence number of the last barrier of the previous solve
    int L, Lmax, Ly, rank, world, tau0, Lglob;
    int d_halo;                     // D has halo slices (sharded handle)
    int x0_given;                   // x holds the initial guess (r0 = b - A x0 computed by the caller)
    double c0, s0, c1, s1, c2, s2, c3, s3;
};

__device__ __forceinline__ void ld_ll(const unsigned long long* p, unsigned long long& a, unsigned long long& b) { ll::ld2(p, a, b); }
__device__ __forceinline__ void push_ll(unsigned long long* p, double v, unsigned int tag) { ll::push(p, v, tag); }
__device__ __forceinline__ double unpack_ll(unsigned long long a, unsigned long long b) { return ll::unpack(a, b); }
__device__ __forceinline__ bool tag_ok(unsigned long long a, unsigned long long b, unsigned int tag) { return ll::tag_ok(a, b, tag); }
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
            for (int q = 0; q < NSEG; ++q) all = all && tag_ok(wa[rr][q], wb[rr][q], tag);
    } while (!all && ++spins < kSpinLimit);
#pragma unroll
    for (int rr = 0; rr < PY; ++rr)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) out[rr][q] = unpack_ll(wa[rr][q], wb[rr][q]);
    return all ? 1 : 0;
}

// Barrier over all CTAs of all GPUs fused with the sums of two doubles per CTA.  `nbar` counts this GPU's barriers of
// the launch from 1 (local arrival target = nbar * gridDim.x), seq is the global sequence number.  Returns false on a
// timeout.  The all-reduce carries no cross-GPU ordering obligation: remote data is only ever read through the
// self-validating halo rows.  v0, v1: per-thread contributions.
__device__ __forceinline__ bool global_sum2(double v0, double v1, const P2pParams& P, unsigned int nbar, unsigned int seq,
                                            double* red, int* flag, double& out0, double& out1) {
    const int nb = gridDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    // CTA sums (fixed order)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, o);
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
    }
    __syncthreads();   // also orders the CTA's global writes of this iteration before the arrival below
    if (lane == 0) { red[warp] = v0; red[32 + warp] = v1; }
    __syncthreads();
    double* part0 = P.partial + (size_t)(seq & 1u) * 2 * P.Lmax;
    double* part1 = part0 + P.Lmax;
    if (threadIdx.x == 0) {
        double b0 = 0.0, b1 = 0.0;
        for (int k = 0; k < nwarps; ++k) { b0 += red[k]; b1 += red[32 + k]; }
        part0[blockIdx.x] = b0;
        part1[blockIdx.x] = b1;
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
    double s0 = 0.0, s1 = 0.0;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) { s0 += __ldcg(part0 + k); s1 += __ldcg(part1 + k); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (lane == 0) { red[warp] = s0; red[32 + warp] = s1; }
    __syncthreads();
    double t0 = 0.0, t1 = 0.0;
    for (int k = 0; k < nwarps; ++k) { t0 += red[k]; t1 += red[32 + k]; }
    bool good = (flag[0] != 0);
    if (P.world > 1) {
        __syncthreads();   // red is reused below
        if (threadIdx.x == 0) {
            const size_t par = (size_t)(seq & 1u) * P.world * 4;
            if (blockIdx.x == 0)
                for (int q = 1; q < P.world; ++q) {
                    unsigned long long* slot = P.mbox[(P.rank + q) % P.world] + par + 4 * P.rank;
                    push_ll(slot, t0, seq);
                    push_ll(slot + 2, t1, seq);
                }
            const unsigned long long* mine = P.mbox[P.rank] + par;
            unsigned long long w[kMaxWorld][4];
            unsigned int spins = 0;
            bool all;
            do {
                all = true;
                for (int g = 0; g < P.world; ++g)
                    if (g != P.rank) { ld_ll(mine + 4 * g, w[g][0], w[g][1]); ld_ll(mine + 4 * g + 2, w[g][2], w[g][3]); }
                for (int g = 0; g < P.world; ++g)
                    if (g != P.rank) all = all && tag_ok(w[g][0], w[g][1], seq) && tag_ok(w[g][2], w[g][3], seq);
            } while (!all && ++spins < kSpinLimit);
            double a0 = 0.0, a1 = 0.0;   // rank order: same bits everywhere
            for (int g = 0; g < P.world; ++g) {
                a0 += (g == P.rank) ? t0 : unpack_ll(w[g][0], w[g][1]);
                a1 += (g == P.rank) ? t1 : unpack_ll(w[g][2], w[g][3]);
            }
            red[0] = a0;
            red[1] = a1;
            flag[1] = all ? 1 : 0;
        }
        __syncthreads();
        t0 = red[0];
        t1 = red[1];
        good = good && (flag[1] != 0);
    }
    out0 = t0;
    out1 = t1;
    __syncthreads();   // red / flag are reused by the caller
    return good;
}

template <int NSEG, int PY, int MAXT, bool SSH>
__global__ void __launch_bounds__(MAXT) cg_p2p_kernel(P2pParams P) {
    constexpr int LX = 32 * NSEG;
    // 128-register cap at 512 threads: x, p, s, D(tau), D(tau+1) live in shared memory there (one CTA per SM anyway)
    constexpr bool XS = (MAXT > 256);
    extern __shared__ __align__(16) double strips[];   // 2 x [nwarps][4][LX]; XS: + x, p, s, D(tau), D(tau+1) [5][N];
                                                       // SSH: + the tables of slices tau and tau+1 (fixed during a solve)
    __shared__ double red[64];
    __shared__ int flag[2];   // barrier completed without a timeout: [0] local counter, [1] peers' mailbox words
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int L = P.L, N = LX * P.Ly;
    const int tau = blockIdx.x;
    const bool multi = (P.world > 1);
    // neighbours: inside the slab plain rows; across the slab boundary the tagged halo rows (world > 1) or the wrapped
    // rows of this GPU (world = 1)
    const bool first = multi && (tau == 0), last = multi && (tau == L - 1);
    const size_t tile_off = (size_t)warp * PY * LX;
    auto eidx = [&](int r, int q) -> size_t { return tile_off + r * LX + 32 * q + lane; };
    const long long row = (long long)tau * N;
    const long long rowm = (long long)((tau == 0) ? L - 1 : tau - 1) * N;      // unused when `first`
    const long long rowp = (long long)((tau == L - 1) ? 0 : tau + 1) * N;      // unused when `last`
    const long long rowDn = P.d_halo ? row + N : rowp;
    // tags: rows written in iteration k+1 (r_{k+1}, w_{k+1}, s_k; parity (k+1)&1) carry base + k + 2; the set-up rows
    // (r_0, w_0; parity 0) and the pushed right-hand side (parity 1) carry base + 1.  The host advances base by
    // iterations + 2 per solve.
    const unsigned int base = P.seq_base;
    const size_t vstride = (size_t)P.Lmax * N, hrow = 2 * (size_t)N;
    auto vec = [&](int par, int which) -> double* { return P.V + (size_t)(par * 3 + which) * vstride; };
    auto halo = [&](const unsigned long long* b, int par, int which, int side) {
        return b + (size_t)((par * 3 + which) * 2 + side) * hrow;
    };
    auto halo_w = [&](unsigned long long* b, int par, int which, int side) {
        return b + (size_t)((par * 3 + which) * 2 + side) * hrow;
    };

    Tile<NSEG, PY> r, w, t1, t2;
    Tile<NSEG, PY> xr, pr, sr, Dcr, Dnr;   // registers when !XS (dead otherwise)
    double* xs = strips + 2ull * nwarps * 4 * LX;
    double* ps = xs + N;
    double* ss = ps + N;
    double* dcs = ss + N;
    double* dns = dcs + N;
    auto sidx = [&](int rr, int q) -> int { return (rr * NSEG + q) * (int)blockDim.x + (int)threadIdx.x; };
    auto X = [&](int rr, int q) -> double& { if constexpr (XS) return xs[sidx(rr, q)]; else return xr.a[rr][q]; };
    auto PP = [&](int rr, int q) -> double& { if constexpr (XS) return ps[sidx(rr, q)]; else return pr.a[rr][q]; };
    auto SS = [&](int rr, int q) -> double& { if constexpr (XS) return ss[sidx(rr, q)]; else return sr.a[rr][q]; };
    auto DC = [&](int rr, int q) -> double& { if constexpr (XS) return dcs[sidx(rr, q)]; else return Dcr.a[rr][q]; };
    auto DN = [&](int rr, int q) -> double& { if constexpr (XS) return dns[sidx(rr, q)]; else return Dnr.a[rr][q]; };

    // SSH: K(tau) and K(tau+1) stay in shared memory for the whole solve (single GPU only: tau wraps)
    const double2* txc = nullptr; const double2* tyc = nullptr; const double2* hyc = nullptr;
    const double2* txn = nullptr; const double2* tyn = nullptr; const double2* hyn = nullptr;
    if constexpr (SSH) {
        static_assert(!XS || !SSH, "SSH tables and the shared-memory state do not fit together");
        double2* tabc = reinterpret_cast<double2*>(strips + 2ull * nwarps * 4 * LX);
        double2* tabn = tabc + 2 * N;
        const int taup_w = (tau == L - 1) ? 0 : tau + 1;
        for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) {
            tabc[i] = P.tab[(size_t)tau * 2 * N + i];
            tabn[i] = P.tab[(size_t)taup_w * 2 * N + i];
        }
        __syncthreads();
        const size_t halo_off = (size_t)((warp * PY + P.Ly - 1) % P.Ly) * LX;
        txc = tabc + tile_off; tyc = tabc + N + tile_off; hyc = tabc + N + halo_off;
        txn = tabn + tile_off; tyn = tabn + N + tile_off; hyn = tabn + N + halo_off;
    }
    const int tg = P.tau0 + tau;                       // global slice index
    const bool wrap_c = (tg == 0);
    const bool wrap_n = (tg + 1 == P.Lglob);
    int xbuf = 0;
    // w(tau) = (M^T M v)(tau) from t1 = v(tau-1), r = v(tau), t2 = v(tau+1) given through `load_next` (called after the
    // first sweeps, so that its loads overlap them); returns this thread's share of |(M v)(tau)|^2
    auto apply_A = [&](auto&& load_next) -> double {
#pragma unroll
        for (int rr = 0; rr < PY; ++rr)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                t1.a[rr][q] = DC(rr, q) * t1.a[rr][q];
                t2.a[rr][q] = DN(rr, q) * r.a[rr][q];
            }
        if constexpr (SSH) {
            g0_tab(t1, txc, lane);
            g0_tab(t2, txn, lane);
            g1_tab(t1, txc, lane);
            g1_tab(t2, txn, lane);
            g2_tab(t1, tyc, lane);
            g2_tab(t2, tyn, lane);
        } else {
            g0_x_even(t1, P.c0, P.s0);
            g0_x_even(t2, P.c0, P.s0);
            g1_x_odd(t1, P.c1, P.s1, lane);
            g1_x_odd(t2, P.c1, P.s1, lane);
            g2_y_even(t1, P.c2, P.s2);
            g2_y_even(t2, P.c2, P.s2);
        }
        {
            double a1[NSEG], a2[NSEG], b1[NSEG], b2[NSEG];
            exchange_edges2(t1, t2, strips + (size_t)xbuf * nwarps * 4 * LX, warp, nwarps, lane, a1, a2, b1, b2);
            xbuf ^= 1;
            if constexpr (SSH) {
                g3_tab(t1, tyc, hyc, lane, a1, b1);
                g3_tab(t2, tyn, hyn, lane, a2, b2);
            } else {
                g3_y_odd(t1, P.c3, P.s3, a1, b1);
                g3_y_odd(t2, P.c3, P.s3, a2, b2);
            }
        }
        double vn[PY][NSEG];
        load_next(vn);
        double acc = 0.0;
#pragma unroll
        for (int rr = 0; rr < PY; ++rr)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const double wc = wrap_c ? (r.a[rr][q] + t1.a[rr][q]) : (r.a[rr][q] - t1.a[rr][q]);
                const double wn = wrap_n ? (vn[rr][q] + t2.a[rr][q]) : (vn[rr][q] - t2.a[rr][q]);
                t1.a[rr][q] = wc;
                t2.a[rr][q] = wn;
                acc = fma(wc, wc, acc);
            }
        {
            double ab[NSEG], be[NSEG];
            exchange_edges1(t2, strips + (size_t)xbuf * nwarps * 4 * LX, warp, nwarps, lane, ab, be);
            xbuf ^= 1;
            if constexpr (SSH) g3_tab(t2, tyn, hyn, lane, ab, be);
            else g3_y_odd(t2, P.c3, P.s3, ab, be);
        }
        if constexpr (SSH) {
            g2_tab(t2, tyn, lane);
            g1_tab(t2, txn, lane);
            g0_tab(t2, txn, lane);
        } else {
            g2_y_even(t2, P.c2, P.s2);
            g1_x_odd(t2, P.c1, P.s1, lane);
            g0_x_even(t2, P.c0, P.s0);
        }
#pragma unroll
        for (int rr = 0; rr < PY; ++rr)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const double du = DN(rr, q) * t2.a[rr][q];
                w.a[rr][q] = wrap_n ? (t1.a[rr][q] + du) : (t1.a[rr][q] - du);
            }
        return acc;
    };

    // ---- set-up: r_0, w_0 = A r_0, gamma_0, delta_0 ----------------------------------------------------------------
    int halo_ok = 1;
    double accg = 0.0;
#pragma unroll
    for (int rr = 0; rr < PY; ++rr)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const size_t e = eidx(rr, q);
            const double bv = P.r0[row + e];
            X(rr, q) = P.x0_given ? P.x[row + e] : 0.0;
            PP(rr, q) = 0.0;
            SS(rr, q) = 0.0;
            r.a[rr][q] = bv;
            DC(rr, q) = SSH ? P.D[e] : P.D[row + e];
            DN(rr, q) = SSH ? P.D[e] : P.D[rowDn + e];
            if (first) push_ll(halo_w(P.left_halo, 1, 0, 1) + 2 * e, bv, base + 1u);   // r_0 rows for the neighbours' set-up
            if (last) push_ll(halo_w(P.right_halo, 1, 0, 0) + 2 * e, bv, base + 1u);
            accg = fma(bv, bv, accg);
        }
    double* ghost_r_lo = P.ghost;
    double* ghost_s_lo = P.ghost + N;
    double* ghost_r_hi = P.ghost + 2 * (size_t)N;
    double* ghost_s_hi = P.ghost + 3 * (size_t)N;
    if (first) {
        double h[PY][NSEG];
        halo_ok &= read_halo<NSEG, PY>(halo(P.my_halo, 1, 0, 0), base + 1u, eidx, h);
#pragma unroll
        for (int rr = 0; rr < PY; ++rr)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                t1.a[rr][q] = h[rr][q];
                ghost_r_lo[eidx(rr, q)] = h[rr][q];
            }
    } else {
#pragma unroll
        for (int rr = 0; rr < PY; ++rr)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) t1.a[rr][q] = P.r0[rowm + eidx(rr, q)];
    }
    double accd = apply_A([&](double (&vn)[PY][NSEG]) {
        if (last) {
            halo_ok &= read_halo<NSEG, PY>(halo(P.my_halo, 1, 0, 1), base + 1u, eidx, vn);
#pragma unroll
            for (int rr = 0; rr < PY; ++rr)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) ghost_r_hi[eidx(rr, q)] = vn[rr][q];
        } else {
#pragma unroll
            for (int rr = 0; rr < PY; ++rr)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) vn[rr][q] = P.r0[rowp + eidx(rr, q)];
        }
    });
    {
        double* R0 = vec(0, 0);
        double* W0 = vec(0, 1);
#pragma unroll
        for (int rr = 0; rr < PY; ++rr)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const size_t e = eidx(rr, q);
                R0[row + e] = r.a[rr][q];
                W0[row + e] = w.a[rr][q];
                if (first) push_ll(halo_w(P.left_halo, 0, 1, 1) + 2 * e, w.a[rr][q], base + 1u);
                if (last) push_ll(halo_w(P.right_halo, 0, 1, 0) + 2 * e, w.a[rr][q], base + 1u);
            }
    }
    const double tol = P.S->tol, kappa_max = P.S->kappa_max;
    const long long maxiter = P.S->maxiter;
    const double normb_in = P.S->normb;
    unsigned int nbar = 0, seq = base;
    double gamma, delta;
    bool alive = (__syncthreads_and(halo_ok) != 0);
    if (alive) alive = global_sum2(accg, accd, P, ++nbar, ++seq, red, flag, gamma, delta);
    const double normb = (normb_in > 0.0) ? normb_in : sqrt(gamma);
    const double eps0 = sqrt(gamma) / normb;
    double alpha = gamma / delta, beta = 0.0, kmin = 0.0, eps = eps0;
    long long j = 0;

    // ---- iterations ------------------------------------------------------------------------------------------------
    while (alive && j < maxiter) {
        const int rd = (int)(j & 1), wr = rd ^ 1;
        const unsigned int tag_rd = base + (unsigned int)j + 1u, tag_wr = tag_rd + 1u;
        ++j;
        const bool have_s = (j > 1);                     // s_{-1} = 0 (beta_0 = 0): nothing to read
        const double* Rr = vec(rd, 0);
        const double* Wr = vec(rd, 1);
        const double* Sr = vec(rd, 2);
        double* Rw = vec(wr, 0);
        double* Ww = vec(wr, 1);
        double* Sw = vec(wr, 2);
        const double mab = -alpha * beta;
        // r_new(tau -+ 1) = r - alpha (w + beta s) of the neighbour slice, rebuilt here
        auto neighbour = [&](bool edge, int side, long long nrow, double (&out)[PY][NSEG]) {
            if (edge) {
                // only w crosses NVLink: this CTA advances its own copies of the neighbour's s and r with the owner's
                // exact operations (s_k = w_k + beta s_{k-1}, r_{k+1} = r_k - alpha s_k), each thread its own elements
                double* gr = side ? ghost_r_hi : ghost_r_lo;
                double* gs = side ? ghost_s_hi : ghost_s_lo;
                halo_ok &= read_halo<NSEG, PY>(halo(P.my_halo, rd, 1, side), tag_rd, eidx, out);
#pragma unroll
                for (int rr = 0; rr < PY; ++rr)
#pragma unroll
                    for (int q = 0; q < NSEG; ++q) {
                        const size_t e = eidx(rr, q);
                        const double sv = have_s ? fma(beta, gs[e], out[rr][q]) : out[rr][q];
                        const double rv = fma(-alpha, sv, gr[e]);
                        gs[e] = sv;
                        gr[e] = rv;
                        out[rr][q] = rv;
                    }
            } else {
#pragma unroll
                for (int rr = 0; rr < PY; ++rr)
#pragma unroll
                    for (int q = 0; q < NSEG; ++q) {
                        const size_t e = eidx(rr, q);
                        double v = fma(-alpha, __ldcg(Wr + nrow + e), __ldcg(Rr + nrow + e));
                        if (have_s) v = fma(mab, __ldcg(Sr + nrow + e), v);
                        out[rr][q] = v;
                    }
            }
        };
        {
            double rm[PY][NSEG];
            neighbour(first, 0, rowm, rm);
#pragma unroll
            for (int rr = 0; rr < PY; ++rr)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) t1.a[rr][q] = rm[rr][q];
        }
        double accr = 0.0;
#pragma unroll
        for (int rr = 0; rr < PY; ++rr)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const size_t e = eidx(rr, q);
                const double pv = fma(beta, PP(rr, q), r.a[rr][q]);
                const double sv = fma(beta, SS(rr, q), w.a[rr][q]);
                PP(rr, q) = pv;
                SS(rr, q) = sv;
                X(rr, q) = fma(alpha, pv, X(rr, q));
                const double rv = fma(-alpha, sv, r.a[rr][q]);
                r.a[rr][q] = rv;
                Rw[row + e] = rv;
                Sw[row + e] = sv;
                accr = fma(rv, rv, accr);
            }
        const double accw = apply_A([&](double (&vn)[PY][NSEG]) { neighbour(last, 1, rowp, vn); });
#pragma unroll
        for (int rr = 0; rr < PY; ++rr)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const size_t e = eidx(rr, q);
                Ww[row + e] = w.a[rr][q];
                if (first) push_ll(halo_w(P.left_halo, wr, 1, 1) + 2 * e, w.a[rr][q], tag_wr);
                if (last) push_ll(halo_w(P.right_halo, wr, 1, 0) + 2 * e, w.a[rr][q], tag_wr);
            }
        if (!__syncthreads_and(halo_ok)) { alive = false; break; }   // a neighbour's tile never arrived
        double gnew, dnew;
        alive = global_sum2(accr, accw, P, ++nbar, ++seq, red, flag, gnew, dnew);
        if (!alive) break;
        eps = sqrt(gnew) / normb;
        const double lg = log(2.0 * eps0 / eps);
        const double qq = 2.0 * (double)j / lg;
        const double kap = qq * qq;
        if (kap > kmin) kmin = kap;
        if (eps < tol || kmin > kappa_max) break;
        beta = gnew / gamma;
        alpha = gnew / (dnew - beta * gnew / alpha);
        gamma = gnew;
    }
#pragma unroll
    for (int rr = 0; rr < PY; ++rr)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) P.x[row + eidx(rr, q)] = X(rr, q);
    if (!alive && threadIdx.x == 0) P.S->done = 2;   // any CTA that saw a timeout marks the solve as failed
    if (alive && blockIdx.x == 0 && threadIdx.x == 0) {
        P.S->iter = j;
        P.S->eps = eps;
        P.S->eps0 = eps0;
        P.S->normb = normb;
        P.S->kappa_min = kmin;
        P.S->done = 1;
    }
}

template <int NSEG, int PY, int MAXT, bool SSH>
size_t p2p_smem(const elph_handle* h, int nwarps) {
    return (2ull * nwarps * 4 * (32 * NSEG) + (MAXT > 256 ? 5ull * h->N : 0)) * sizeof(double) +
           (SSH ? 2ull * 2 * h->N * sizeof(double2) : 0);
}

// all slices of the slab must be co-resident (cooperative launch, one CTA per slice)
template <int NSEG, int PY, int MAXT, bool SSH = false>
bool fits_p2p(elph_handle* h, int nwarps) {
    auto kern = cg_p2p_kernel<NSEG, PY, MAXT, SSH>;
    const size_t smem = p2p_smem<NSEG, PY, MAXT, SSH>(h, nwarps);
    if (smem > h->smem_optin) return false;
    elph_enable_smem(h, kern);
    int per_sm = 0;
    ELPH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nwarps * 32, smem));
    return (long long)per_sm * h->sm_count >= h->L;
}

template <int NSEG, int PY, int MAXT, bool SSH = false>
bool launch_p2p(elph_handle* h, P2pParams& P, int nwarps) {
    if (!fits_p2p<NSEG, PY, MAXT, SSH>(h, nwarps)) return false;
    auto kern = cg_p2p_kernel<NSEG, PY, MAXT, SSH>;
    ELPH_CUDA(cudaMemsetAsync(h->d_bar, 0, sizeof(unsigned int), h->stream));
    void* args[] = {&P};
    ELPH_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(h->L), dim3(nwarps * 32), args, p2p_smem<NSEG, PY, MAXT, SSH>(h, nwarps),
                                          h->stream));
    h->launches++;
    return true;
}

// kernel variant for this handle: 0 = none applies
int p2p_variant(const elph_handle* h, int& nwarps) {
    const bool ssh = (h->model == ELPH_MODEL_SSH);
    if (!(ssh ? h->ssq.enabled : h->sq.enabled)) return 0;
    const int Lx = ssh ? h->ssq.Lx : h->sq.Lx, Ly = ssh ? h->ssq.Ly : h->sq.Ly;
    const int PY = (Lx == 32) ? 8 : 4;
    if (Ly % PY) return 0;
    nwarps = Ly / PY;
    if (nwarps < 2 || nwarps > 32) return 0;
    if (ssh) return (Lx == 32 && nwarps * 32 <= 256 && !h->sharded) ? 3 : 0;   // single GPU only
    if (Lx == 32 && nwarps * 32 <= 256) return 1;
    if (Lx == 64 && nwarps * 32 <= 512) return 2;
    return 0;
}

bool p2p_fits(elph_handle* h) {
    int nwarps = 0;
    const int variant = p2p_variant(h, nwarps);
    return (variant == 1) ? fits_p2p<1, 8, 256>(h, nwarps) : (variant == 2) ? fits_p2p<2, 4, 512>(h, nwarps)
         : (variant == 3) ? fits_p2p<1, 8, 256, true>(h, nwarps) : false;
}

// ---- arena layout (identical on every rank) ------------------------------------------------------------------------
size_t arena_vec_doubles(const elph_handle* h) { return 6 * (size_t)h->p2p.Lmax * h->N; }       // [2][3][Lmax][N]
size_t arena_halo_words(const elph_handle* h) { return 12 * 2 * (size_t)h->N; }                  // [2][3][2][N][2]
size_t arena_mbox_words() { return 2ull * kMaxWorld * 4; }
size_t arena_p2p_bytes(const elph_handle* h) {
    const size_t b = arena_vec_doubles(h) * sizeof(double) + arena_halo_words(h) * sizeof(unsigned long long) +
                     arena_mbox_words() * sizeof(unsigned long long) + (4 * (size_t)h->p2p.Lmax + 4 * (size_t)h->N) * sizeof(double) + 256;
    return (b + 255) & ~size_t(255);
}
// the region of the pipelined kernel (cg_pipe.cu) follows
size_t arena_bytes(const elph_handle* h) { return arena_p2p_bytes(h) + elph_pipe_arena_bytes(h->N, h->p2p.Lmax); }
double* arena_vec(const elph_handle* h, void* base) { return reinterpret_cast<double*>(base); }
unsigned long long* arena_halo(const elph_handle* h, void* base) {
    return reinterpret_cast<unsigned long long*>(reinterpret_cast<double*>(base) + arena_vec_doubles(h));
}
unsigned long long* arena_mbox(const elph_handle* h, void* base) { return arena_halo(h, base) + arena_halo_words(h); }
double* arena_partial(const elph_handle* h, void* base) { return reinterpret_cast<double*>(arena_mbox(h, base) + arena_mbox_words()); }
double* arena_ghost(const elph_handle* h, void* base) { return arena_partial(h, base) + 4 * (size_t)h->p2p.Lmax; }

void alloc_arena(elph_handle* h, int rank, int world, int Lglob) {
    auto& A = h->p2p;
    const int Lmax = (Lglob + world - 1) / world;
    if (A.arena && A.rank == rank && A.world == world && A.Lmax == Lmax) return;
    if (A.arena) elph_shard_p2p_close_impl(h);   // geometry changed (e.g. elph_set_shard after a single-GPU solve): start over
    A.rank = rank;
    A.world = world;
    A.Lmax = (Lglob + world - 1) / world;
    ELPH_REQUIRE(h->L <= A.Lmax, ELPH_ERR_INVALID, "slab longer than ceil(Lglob / world)");
    ELPH_CUDA(cudaMalloc(&A.arena, arena_bytes(h)));
    ELPH_CUDA(cudaMemset(A.arena, 0, arena_bytes(h)));
    ELPH_CUDA(cudaDeviceSynchronize());
    A.pipe_off = arena_p2p_bytes(h);
    A.seq = 0;
    A.pipe_seq = 0;
    A.hx_seq = 0;
    A.failed = false;
    A.pipe_failed = false;
    if (!h->h_hx_flag) ELPH_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h->h_hx_flag), sizeof(unsigned int), cudaHostAllocMapped));
    *h->h_hx_flag = 0u;
}

// run one solve on an opened arena; r0/x as in P2pParams.  Returns false if the kernel does not apply.
// scalars_on_device: tol, kappa_max, maxiter, normb were left in h->d_cg by cg_init_kernel (general initial guess).
bool run_p2p(elph_handle* h, const double* r0, double* x, bool x0_given, bool scalars_on_device, double tol, int64_t maxiter) {
    auto& A = h->p2p;
    int nwarps = 0;
    const int variant = p2p_variant(h, nwarps);
    if (!variant) return false;
    ELPH_REQUIRE(!A.failed, ELPH_ERR_STATE, "peer-memory CG: an earlier solve on this arena timed out; re-open the peer arenas");
    cudaStream_t st = h->stream;
    const int left = (A.rank + A.world - 1) % A.world, right = (A.rank + 1) % A.world;
    P2pParams P;
    P.D = h->d_D; P.tab = (variant == 3) ? h->ssq.d_tab : nullptr; P.r0 = r0; P.x = x;
    P.V = arena_vec(h, A.arena);
    P.my_halo = arena_halo(h, A.arena);
    P.left_halo = arena_halo(h, A.peer[left]);    // this GPU's first slice is the left neighbour's slice L (hi rows)
    P.right_halo = arena_halo(h, A.peer[right]);  // its last slice is the right neighbour's slice -1 (lo rows)
    for (int q = 0; q < kMaxWorld; ++q) P.mbox[q] = (q < A.world) ? arena_mbox(h, A.peer[q]) : nullptr;
    P.partial = arena_partial(h, A.arena); P.ghost = arena_ghost(h, A.arena); P.bar = h->d_bar; P.S = h->d_cg;
    P.seq_base = A.seq;
    P.L = h->L; P.Lmax = A.Lmax; P.Ly = (variant == 3) ? h->ssq.Ly : h->sq.Ly; P.rank = A.rank; P.world = A.world;
    P.tau0 = h->sharded ? h->shard_tau0 : 0;
    P.Lglob = h->sharded ? h->shard_Lglob : h->L;
    P.d_halo = h->sharded ? 1 : 0;
    P.x0_given = x0_given ? 1 : 0;
    P.c0 = h->sq.c[0]; P.s0 = h->sq.s[0]; P.c1 = h->sq.c[1]; P.s1 = h->sq.s[1];
    P.c2 = h->sq.c[2]; P.s2 = h->sq.s[2]; P.c3 = h->sq.c[3]; P.s3 = h->sq.s[3];
    if (!scalars_on_device) {
        CgScalars init = {};
        init.tol = tol; init.kappa_max = h->cg_kappa_max; init.maxiter = maxiter; init.normb = 0.0;   // |b| = |r0|
        *h->h_cg = init;
        ELPH_CUDA(cudaMemcpyAsync(h->d_cg, h->h_cg, sizeof(CgScalars), cudaMemcpyHostToDevice, st));
    }
    bool ok = false;
    if (variant == 1) ok = launch_p2p<1, 8, 256>(h, P, nwarps);
    else if (variant == 2) ok = launch_p2p<2, 4, 512>(h, P, nwarps);
    else if (variant == 3) ok = launch_p2p<1, 8, 256, true>(h, P, nwarps);
    if (!ok) return false;
    ELPH_CUDA(cudaMemcpyAsync(h->h_cg, h->d_cg, sizeof(CgScalars), cudaMemcpyDeviceToHost, st));
    ELPH_CUDA(cudaStreamSynchronize(st));
    if (h->h_cg->done != 1) {
        // tags up to base + j + 1 of an unknown j may sit in the peers' arenas and the ranks may disagree on the count:
        // the arena is unusable until every rank re-opens it (elph_shard_p2p_export / _open re-zero it and restart the tags)
        A.failed = true;
        ELPH_REQUIRE(false, ELPH_ERR_STATE, "peer-memory CG: a peer GPU did not reach a barrier or deliver a halo tile (timeout)");
    }
    // barriers executed: 1 + iterations (sequence numbers base + 1 .. base + 1 + iterations); tags used: up to base +
    // iterations + 1.  Advancing by iterations + 3 puts the first barrier of the next solve on the OTHER mailbox parity than
    // the last barrier of this one, so a rank that has already relaunched cannot overwrite a slot that a lagging CTA of this
    // solve is still polling.
    A.seq += 3u + (unsigned int)h->h_cg->iter;
    return true;
}

}  // namespace

// Allocate the arena (once) and export its IPC handle (64 bytes).
void elph_shard_p2p_export_impl(elph_handle* h, int rank, int world, unsigned char* handle_out) {
    ELPH_REQUIRE(h->sharded, ELPH_ERR_STATE, "elph_set_shard has not been called");
    ELPH_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, ELPH_ERR_INVALID, "bad rank / world");
    ELPH_REQUIRE(h->sq.enabled && h->model == ELPH_MODEL_HOLSTEIN, ELPH_ERR_UNSUPPORTED,
                 "the peer-memory CG needs the Holstein square-lattice register kernels");
    auto& A = h->p2p;
    const bool had = (A.arena != nullptr);
    alloc_arena(h, rank, world, h->shard_Lglob);
    if (had) {   // re-export (every rank does this together, e.g. after a timeout): stale tags must not survive
        ELPH_CUDA(cudaStreamSynchronize(h->stream));
        ELPH_CUDA(cudaMemset(A.arena, 0, arena_bytes(h)));
        ELPH_CUDA(cudaDeviceSynchronize());
        A.seq = 0; A.pipe_seq = 0; A.hx_seq = 0; A.failed = false; A.pipe_failed = false;
        *h->h_hx_flag = 0u;
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
    A.opened = true;   // the halo exchange of the products works from here on; the persistent CG needs co-resident slabs on top
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

// the pipelined kernel (cg_pipe.cu) is the default; forcing one of the older forms with tuning key 7 selects that form
static bool want_pipeline(const elph_handle* h) {
    return h->cg_pipeline == 1 || (h->cg_pipeline < 0 && h->cg_single_reduction < 0);
}

bool elph_shard_cg_available_impl(elph_handle* h) { return h->p2p.opened && (p2p_fits(h) || elph_cg_pipe_fits(h)); }

// Solve A x = b with x0 = 0 on the slab owned by this rank; every rank of the ring must make the same call.
// b_own / x_own: [L][N] own slices (engine layout).  Returns false if the kernel does not apply (caller falls back).
bool elph_shard_cg_p2p_impl(elph_handle* h, const double* b_own, double* x_own, double tol, int64_t maxiter, int64_t* iters,
                            double* eps) {
    auto& A = h->p2p;
    ELPH_REQUIRE(A.opened, ELPH_ERR_STATE, "elph_shard_p2p_open has not been called");
    if (tol == 0.0) tol = h->cg_tol;
    if (maxiter == 0) maxiter = h->cg_maxiter;
    const bool piped = want_pipeline(h) && elph_cg_pipe_run(h, b_own, x_own, false, false, tol, maxiter);
    if (!piped && !run_p2p(h, b_own, x_own, false, false, tol, maxiter)) return false;
    if (iters) *iters = h->h_cg->iter;
    if (eps) *eps = h->h_cg->eps;
    return true;
}

// Single-GPU use on an unsharded handle (elph_cg_device): r0 in h->d_r, scalars in h->d_cg (cg_init_kernel), x_dev holds
// the initial guess.  The ring closes inside the GPU; the arena is private.  Returns false if the kernel does not apply.
bool elph_cg_single_reduction(elph_handle* h, double* x_dev) {
    if (h->sharded || h->sq_disable || h->L < 2) return false;
    if (!((h->model == ELPH_MODEL_SSH) ? h->ssq.enabled : h->sq.enabled)) return false;
    auto& A = h->p2p;
    if (!A.arena) {
        alloc_arena(h, 0, 1, h->L);
        A.peer.assign(1, A.arena);
        A.peer_L.assign(1, h->L);
        A.opened = true;
    }
    // pipelined kernel: the reduction is off the critical path and a slice may be split over several SMs
    if (want_pipeline(h) && elph_cg_pipe_run(h, h->d_r, x_dev, true, true, 0.0, 0)) return true;
    if (h->cg_single_reduction == 0) return false;
    int nwarps = 0;
    const int variant = p2p_variant(h, nwarps);
    if (!variant) return false;
    // measured on B200 (scripts/bench_cg1r.py): 32x32xL200 6.93 -> 5.85 us/iteration; 64-wide lattices keep more state in
    // shared memory under the 128-register cap and lose (9.7 -> 10.6 us), so they stay on the two-reduction kernel
    if (h->cg_single_reduction < 0 && variant == 2) return false;
    if (!p2p_fits(h)) return false;
    return run_p2p(h, h->d_r, x_dev, true, true, 0.0, 0);
}
