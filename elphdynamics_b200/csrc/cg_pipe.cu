// Pipelined conjugate gradient on A = M^T M as ONE persistent kernel per GPU: the all-reduce of an iteration travels
// while the next product is computed.  One lattice on one GPU, or tau-sharded over several GPUs with halo and all-reduce
// inside the kernel over NVLink peer memory.  Replaces solve! of src/IterativeSolvers.jl:239-314 (same
// Krylov iterates, stop rule and iteration numbering) for periodic square lattices (register tiles of square_tiles.cuh).
//
// Why: a CG iteration of one lattice is latency bound -- ~1.5 us of arithmetic against a 2.6 us grid barrier (5 us
// across GPUs).  cg_p2p.cu needs one barrier per iteration; here the barrier disappears from the critical path
// (Ghysels & Vanroose 2014, "pipelined CG"):
//
//     gamma_j = (r_j, r_j), delta_j = (w_j, r_j)      -> posted, reduced by a dedicated CTA while ...
//     q_j = A w_j                                     ... the product runs
//     beta_j = gamma_j / gamma_{j-1},  alpha_j = gamma_j / (delta_j - beta_j gamma_j / alpha_{j-1})
//     z_j = q_j + beta_j z_{j-1};  s_j = w_j + beta_j s_{j-1};  p_j = r_j + beta_j p_{j-1}
//     x_{j+1} = x_j + alpha_j p_j;  r_{j+1} = r_j - alpha_j s_j;  w_{j+1} = w_j - alpha_j z_j     (w = A r, z = A s)
//
// Iteration counts equal the reference recurrences' at every configuration measured on the CPU (scripts/cg_variants_study.py:
// 607/607, 963/964 at 32x32xL200; 1094/1094, 2037/2037 at 64x64xL400) and the parity tests require +-2.
//
// Decomposition: a CTA owns Ly/YS rows of ONE time slice (state in registers / shared memory for the whole solve); the
// YS CTAs of a slice form a thread-block cluster and swap the rows at their common edges through distributed shared
// memory (two cluster barriers per product).  So a 64x64 slice is swept by up to 4 SMs instead of one.
//
// No grid barrier, no flags, no fences.  Everything that crosses CTAs travels as self-validating words (ll_words.cuh):
//   * neighbour slices: CTA tau publishes q_j(tau) (one row, 16 bytes per site) and keeps private "ghost" copies of
//     w(tau-1), z(tau-1), w(tau+1), z(tau+1) which it advances with the owner's exact operations -- the ghost of
//     w_{j+1}(tau+-1) is what the product needs.  The row of a slice on the neighbour GPU is pushed straight into that
//     GPU's arena over NVLink: intra- and inter-GPU neighbours are the same code.  A row published during iteration j is
//     read during iteration j+1, a whole reduction latency later: the wait is normally already satisfied.
//   * all-reduce: every CTA posts its two partial sums into its slot; one extra CTA per GPU (the reducer) polls the slots,
//     adds them in index order, swaps the GPU totals with the other GPUs through mailboxes, adds those in rank order and
//     publishes {gamma, delta, status} in a broadcast slot that the working CTAs poll -- identical bits everywhere, so all
//     CTAs of all GPUs take the same branch.  The reducer replays the scalar recurrences to know when to stop.
// Tags grow monotonically over the life of the arena (2^32 steps); a timeout anywhere raises the arena's abort word, the
// reducer turns it into status = 1 and every CTA leaves at the same iteration (ELPH_ERR_STATE on the host, re-open needed).
#include "ll_words.cuh"
#include "square_tiles.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <cstring>

namespace cgx = cooperative_groups;

namespace {

using namespace sqt;

constexpr int kMaxWorld = 16;
constexpr unsigned int kSpinLimit = 1u << 25;   // per wait, ~10-20 s: a dead peer ends the solve with an error, not a hang
constexpr int kMaxYS = 8;
constexpr int kBcastCopies = 8;      // the broadcast message is replicated: CTA c polls copy c % 8 (spreads the pollers over L2 slices)
constexpr int kBcastStride = 64;     // words between copies (512 bytes)

// where the state tiles live (bit mask)
enum : int {
    PL_XP_SMEM = 1,     // x, p in shared memory
    PL_SZ_SMEM = 2,     // s, z in shared memory
    PL_D_SMEM = 4,      // D(tau), D(tau+1) in shared memory
    PL_XP_GLOBAL = 8,   // x in the output vector, p in a scratch vector (L2)
    PL_GHOST_GLOBAL = 16,
    PL_R_SMEM = 32,     // r in shared memory
    PL_PREFETCH = 64,   // the neighbours' rows are copied into shared memory with cp.async while the update runs
};

struct PipeParams {
    const double* __restrict__ D;     // Holstein expnV (sharded handle: halo slices around the own ones); SSH: exp(dtau mu) [N]
    const double2* __restrict__ tab;  // SSH: (cosh, sinh) [L][2][N] in the tile layout of ssh_square.cu
    const double* __restrict__ r0;    // [L][N] initial residual
    double* x;                        // [L][N] in (x0_given) / out
    unsigned long long* rows;         // own arena: [2 parities][Lmax + 2][N][2 words]; row 0 / 1 = lo / hi halo, 2 + tau = own
    unsigned long long* left_rows;    // the left / right neighbour GPU's rows (peer memory): this GPU writes the hi halo of the
    unsigned long long* right_rows;   // left one and the lo halo of the right one
    unsigned long long* part;         // [2 parities][maxcta][4 words] partial sums of the CTAs
    unsigned long long* bcast;        // [2 parities][kBcastCopies][kBcastStride words] {alpha, beta, flags}
    unsigned long long* mbox[kMaxWorld];   // mailbox of every rank: [2 parities][kMaxWorld][4 words]
    unsigned int* abort_word;
    double* pg;                       // [Lmax][N] p when PL_XP_GLOBAL
    double* ghost;                    // [4][Lmax][N] when PL_GHOST_GLOBAL
    double* state;                    // [6][Lmax][N] r, w, s, z, q, p in multi-slice mode
    CgScalars* S;
    unsigned long long* prof;         // development aid (tuning key 12): [cta][8] cycles per phase, summed over the iterations
    unsigned int base;                // tags of this solve: publication k -> base + 1 + k, reduction n -> base + 1 + n
    int sync_mode;
    int L, Lmax, Ly, ys, spc, maxcta, rank, world, tau0, Lglob, d_halo, x0_given;
    double c0, s0, c1, s1, c2, s2, c3, s3;
};

// scalar recurrences + stop rule (src/IterativeSolvers.jl:198-231), evaluated ONCE per reduction by the reducer CTA
struct Rec {
    double normb, eps0, eps, kmin, alpha, beta, gam_old, alpha_old;
    long long j;
};
// alpha_j, beta_j from the sums of reduction j (three dependent divisions: on the critical path of every iteration)
__device__ __forceinline__ void rec_coeff(Rec& R, double gam, double del) {
    if (R.j == 0) {
        R.beta = 0.0;
        R.alpha = gam / del;
    } else {
        R.beta = gam / R.gam_old;
        R.alpha = gam / (del - R.beta * gam / R.alpha_old);
    }
    R.gam_old = gam;
    R.alpha_old = R.alpha;
}
// the stop rule for j completed iterations (sqrt, log: off the critical path, published one slot later)
__device__ __forceinline__ bool rec_stop(Rec& R, double gam, double tol, double kappa_max, long long maxiter, double normb_in) {
    if (R.j == 0) {
        R.normb = (normb_in > 0.0) ? normb_in : sqrt(gam);
        R.eps0 = sqrt(gam) / R.normb;
        R.eps = R.eps0;
        R.kmin = 0.0;
    } else {
        R.eps = sqrt(gam) / R.normb;
        const double lg = log(2.0 * R.eps0 / R.eps);
        const double qq = 2.0 * (double)R.j / lg;
        const double kap = qq * qq;
        if (kap > R.kmin) R.kmin = kap;
        if (R.eps < tol || R.kmin > kappa_max) return true;
    }
    return R.j >= maxiter;
}

__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- the reducer CTA ------------------------------------------------------------------------------------------------
// Measured on B200 (scripts/micro/pingpong_bench.cu): a word stored by one SM is seen by a polling thread of another SM after
// ~860 cycles (0.45 us), whatever the number of pollers; release/acquire pairs cost 1500.  The all-reduce is two such hops
// (partial sums -> reducer -> message) plus the reducer's arithmetic.  A one-hop variant (every CTA gathers all partial sums
// itself) was tried and lost: 5.4 against 3.8 us per iteration at 32x32xL200 -- 200 CTAs polling 200 slots each slow the
// products of the CTAs that share their SM.
__device__ void reducer_loop(const PipeParams& P, int ncta, double (*rsum)[8], double* rtot) {
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const double tol = P.S->tol, kappa_max = P.S->kappa_max, normb_in = P.S->normb;
    const long long maxiter = P.S->maxiter;
    Rec R = {};                       // advanced by thread 0 only
    bool failed = false, stop_prev = false, abort_seen = false;
    unsigned long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool prof = (P.prof != nullptr) && tid == 0;
    long long tk = prof ? clock64() : 0;
    auto tick = [&](int k) { if (prof) { const long long now = clock64(); pc[k] += (unsigned long long)(now - tk); tk = now; } };
    for (unsigned int n = 0;; ++n) {
        const unsigned int tag = P.base + 1u + n;
        const unsigned long long* slots = P.part + (size_t)(tag & 1u) * P.maxcta * 4;
        double s0 = 0.0, s1 = 0.0;
        for (int k0 = tid; k0 < ncta; k0 += 4 * T) {       // four slots per thread and round, slots in index order
            unsigned long long w[4][4];
            unsigned int spins = 0;
            bool all;
            do {
                all = true;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = k0 + u * T;
                    if (k < ncta) {
                        ll::ld2(slots + 4 * (size_t)k, w[u][0], w[u][1]);
                        ll::ld2(slots + 4 * (size_t)k + 2, w[u][2], w[u][3]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (k0 + u * T < ncta) all = all && ll::tag_ok(w[u][0], w[u][1], tag) && ll::tag_ok(w[u][2], w[u][3], tag);
                if (!all && ((++spins & 1023u) == 0u) && (spins > kSpinLimit || ld_volatile_u32(P.abort_word) != 0u)) break;
            } while (!all);
            if (!all) failed = true;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (k0 + u * T < ncta) { s0 += ll::unpack(w[u][0], w[u][1]); s1 += ll::unpack(w[u][2], w[u][3]); }
        }
        tick(0);     // thread 0's own slots have arrived
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        if (lane == 0) { rsum[0][warp] = s0; rsum[1][warp] = s1; }
        const int any_failed = __syncthreads_or(failed ? 1 : 0);
        tick(1);     // everybody's slots have arrived
        if (tid == 0) {
            double t0 = 0.0, t1 = 0.0;
            for (int k = 0; k < nwarps; ++k) { t0 += rsum[0][k]; t1 += rsum[1][k]; }
            bool bad = (any_failed != 0) || abort_seen;
            if (P.world > 1) {
                const size_t par = (size_t)(tag & 1u) * kMaxWorld * 4;
                for (int q = 1; q < P.world; ++q) {
                    unsigned long long* slot = P.mbox[(P.rank + q) % P.world] + par + 4 * P.rank;
                    ll::push(slot, t0, tag);
                    ll::push(slot + 2, t1, tag);
                }
                const unsigned long long* mine = P.mbox[P.rank] + par;
                unsigned long long w[kMaxWorld][4];
                unsigned int spins = 0;
                bool all;
                do {
                    all = true;
                    for (int g = 0; g < P.world; ++g)
                        if (g != P.rank) { ll::ld2(mine + 4 * g, w[g][0], w[g][1]); ll::ld2(mine + 4 * g + 2, w[g][2], w[g][3]); }
                    for (int g = 0; g < P.world; ++g)
                        if (g != P.rank) all = all && ll::tag_ok(w[g][0], w[g][1], tag) && ll::tag_ok(w[g][2], w[g][3], tag);
                    if (!all && ((++spins & 1023u) == 0u) && (spins > kSpinLimit || ld_volatile_u32(P.abort_word) != 0u)) break;
                } while (!all);
                if (!all) bad = true;
                double a0 = 0.0, a1 = 0.0;   // rank order: the same bits on every GPU
                for (int g = 0; g < P.world; ++g) {
                    a0 += (g == P.rank) ? t0 : ll::unpack(w[g][0], w[g][1]);
                    a1 += (g == P.rank) ? t1 : ll::unpack(w[g][2], w[g][3]);
                }
                t0 = a0;
                t1 = a1;
            }
            tick(2);     // cross-GPU exchange
            // message n = {alpha_n, beta_n, flags}: bit 0 = abort, bit 1 = "n-1 iterations were enough" (the stop rule of
            // reduction n-1, evaluated AFTER message n-1 had left: sqrt, log and a division stay off the critical path; the
            // CTAs commit x_n only when message n tells them to go on).  Copies in kBcastCopies lines = L2 slices.
            rec_coeff(R, t0, t1);
            tick(3);
            const double flags = (bad ? 1.0 : 0.0) + (stop_prev ? 2.0 : 0.0);
            unsigned long long* bc = P.bcast + (size_t)(tag & 1u) * kBcastCopies * kBcastStride;
            for (int c = 0; c < kBcastCopies; ++c) {
                ll::push(bc + c * kBcastStride, R.alpha, tag);
                ll::push(bc + c * kBcastStride + 2, R.beta, tag);
                ll::push(bc + c * kBcastStride + 4, flags, tag);
            }
            tick(4);     // message pushed
            const bool leave = bad || stop_prev;
            if (!leave) {
                stop_prev = rec_stop(R, t0, tol, kappa_max, maxiter, normb_in);
                if (!stop_prev) ++R.j;
                abort_seen = (ld_volatile_u32(P.abort_word) != 0u);     // reported with the next message
            }
            tick(5);
            rtot[0] = leave ? 1.0 : 0.0; rtot[1] = bad ? 1.0 : 0.0;
        }
        __syncthreads();
        const bool leave = (rtot[0] != 0.0);
        if (rtot[1] != 0.0) failed = true;
        __syncthreads();
        if (leave) break;
    }
    if (prof)
        for (int k = 0; k < 8; ++k) P.prof[(size_t)blockIdx.x * 8 + k] = pc[k];
    if (tid == 0) {
        failed = (rtot[1] != 0.0) || failed;
        P.S->iter = R.j;
        P.S->eps = R.eps;
        P.S->eps0 = R.eps0;
        P.S->normb = R.normb;
        P.S->kappa_min = R.kmin;
        P.S->done = failed ? 2 : 1;
    }
}

// MS (multi-slice, "streaming"): a CTA owns P.spc CONSECUTIVE time slices of its rows and keeps all vectors in global memory
// (L2 resident at the slab sizes this is for: 64x64xL200 per GPU = 46 MB of state); only the two outer neighbours of the
// chunk need ghosts and rows.  The slab sizes whose state does not fit registers + shared memory run this way.
template <int NSEG, int PY, int MAXT, int MINB, int PLACE, bool SSH, bool MS = false>
__global__ void __launch_bounds__(MAXT, MINB) cgpipe_kernel(PipeParams P) {
    constexpr int LX = 32 * NSEG;
    constexpr bool XPS = (PLACE & PL_XP_SMEM) != 0, SZS = (PLACE & PL_SZ_SMEM) != 0, DS = (PLACE & PL_D_SMEM) != 0;
    constexpr bool XPG = (PLACE & PL_XP_GLOBAL) != 0, GG = (PLACE & PL_GHOST_GLOBAL) != 0, RS = (PLACE & PL_R_SMEM) != 0;
    constexpr bool PF = (PLACE & PL_PREFETCH) != 0;
    static_assert(!(XPS && XPG), "x, p: shared memory or global, not both");
    static_assert(!MS || (GG && !XPS && !SZS && !DS && !RS && !SSH), "multi-slice mode: all state in global memory, Holstein");
    extern __shared__ __align__(16) double smem[];
    __shared__ double red[2][8];
    __shared__ double cf[4];
    const int YS = P.ys;
    const int c = blockIdx.x / YS, yb = blockIdx.x - c * YS;
    const int L = P.L;
    const int SPC = MS ? P.spc : 1;                       // slices per CTA
    const int nchunk = (L + SPC - 1) / SPC;
    if (c >= nchunk) {                  // the extra cluster: its first CTA is the reducer of this GPU
        if (yb == 0) reducer_loop(P, nchunk * YS, red, cf);
        return;
    }
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, NW = T >> 5;
    const int N = LX * P.Ly;
    const int NB = PY * NW * LX;        // sites of this CTA
    const int tau = c * SPC;                              // first own slice
    const int ns = (L - tau < SPC) ? L - tau : SPC;       // own slices: tau .. taul
    const int taul = tau + ns - 1;
    const bool multi = (P.world > 1);

    // ---- shared memory carve-up ------------------------------------------------------------------------------------
    double* strips = smem;                                   // [2][NW][4][LX]
    double* sp = smem + 2ull * NW * 4 * LX;
    double *gwl = nullptr, *gzl = nullptr, *gwh = nullptr, *gzh = nullptr;
    if constexpr (!GG) { gwl = sp; gzl = sp + NB; gwh = sp + 2 * NB; gzh = sp + 3 * NB; sp += 4 * NB; }
    double *xs = nullptr, *ps = nullptr, *ss = nullptr, *zs = nullptr, *dcs = nullptr, *dns = nullptr, *rs = nullptr;
    if constexpr (XPS) { xs = sp; ps = sp + NB; sp += 2 * NB; }
    if constexpr (SZS) { ss = sp; zs = sp + NB; sp += 2 * NB; }
    if constexpr (DS) { dcs = sp; dns = sp + NB; sp += 2 * NB; }
    if constexpr (RS) { rs = sp; sp += NB; }
    ulonglong2* pre = nullptr;                               // [2 sides][PY * NSEG][T] 16-byte entries
    if constexpr (PF) { pre = reinterpret_cast<ulonglong2*>(sp); sp += 4 * NB; }

    const size_t tile_off = (size_t)((yb * NW + warp) * PY) * LX;      // first site of this warp's rows within a slice
    auto eidx = [&](int rr, int q) -> size_t { return tile_off + rr * LX + 32 * q + lane; };
    auto sidx = [&](int rr, int q) -> int { return (rr * NSEG + q) * T + tid; };
    const long long row = (long long)tau * N;
    const long long rowm = (long long)((tau == 0) ? L - 1 : tau - 1) * N;
    const long long rowp = (long long)((tau == L - 1) ? 0 : tau + 1) * N;
    const long long rowDn = P.d_halo ? row + N : rowp;
    const bool first = multi && (tau == 0), last = multi && (taul == L - 1);
    if constexpr (GG) {
        // ghosts of the slice below the first own one are kept at the first own slice's position, those of the slice above the
        // last own one at the last own slice's position (indexed by gidx below)
        const size_t gs = (size_t)P.Lmax * N;
        gwl = P.ghost + row + tile_off - tid; gzl = gwl + gs;
        gwh = P.ghost + 2 * gs + (long long)taul * N + tile_off - tid; gzh = gwh + gs;
    }
    // ghost element index: shared memory -> sidx, global -> position in the slice
    auto gidx = [&](int rr, int q) -> size_t {
        if constexpr (GG) return (size_t)tid + rr * LX + 32 * q + lane; else return (size_t)sidx(rr, q);
    };

    Tile<NSEG, PY> w, t1, t2;
    Tile<NSEG, PY> rr_, xr, pr, sr, zr, Dcr, Dnr;     // registers unless placed elsewhere (dead then)
    double* pgl = XPG ? P.pg + row : nullptr;
    double* xgl = XPG ? P.x + row : nullptr;
    auto X = [&](int a, int q) -> double& { if constexpr (XPS) return xs[sidx(a, q)]; else if constexpr (XPG) return xgl[eidx(a, q)]; else return xr.a[a][q]; };
    auto PP = [&](int a, int q) -> double& { if constexpr (XPS) return ps[sidx(a, q)]; else if constexpr (XPG) return pgl[eidx(a, q)]; else return pr.a[a][q]; };
    auto SS = [&](int a, int q) -> double& { if constexpr (SZS) return ss[sidx(a, q)]; else return sr.a[a][q]; };
    auto ZZ = [&](int a, int q) -> double& { if constexpr (SZS) return zs[sidx(a, q)]; else return zr.a[a][q]; };
    auto DC = [&](int a, int q) -> double& { if constexpr (DS) return dcs[sidx(a, q)]; else return Dcr.a[a][q]; };
    auto DN = [&](int a, int q) -> double& { if constexpr (DS) return dns[sidx(a, q)]; else return Dnr.a[a][q]; };
    auto RR = [&](int a, int q) -> double& { if constexpr (RS) return rs[sidx(a, q)]; else return rr_.a[a][q]; };

    // ---- y-edge exchange: inside the CTA through the strips, across the CTAs of the slice through DSMEM -----------
    cgx::cluster_group cluster = cgx::this_cluster();
    const double* strips_up = strips;      // the CTA holding the rows above / below (periodic in y)
    const double* strips_dn = strips;
    if (YS > 1) {
        strips_up = cluster.map_shared_rank(strips, (yb + YS - 1) % YS);
        strips_dn = cluster.map_shared_rank(strips, (yb + 1) % YS);
    }
    int xbuf = 0;
    bool dead = false;                     // a row never arrived: stop waiting; the reducer turns the abort word into an abort message
    // cluster.sync() = MEMBAR.ALL.GPU + barrier: the fence waits for every outstanding global store (the 16-byte row pushes) and
    // costs more than the barrier.  Only the shared-memory strips have to be ordered here: they live in the owning SM, a
    // CTA-scope fence completes this CTA's stores to them before it arrives, and the partners read them after the wait.
    auto edge_sync = [&]() {
        if (YS > 1) {
            if (P.sync_mode == 0) cluster.sync();
            else if (P.sync_mode == 1) {
                __threadfence_block();
                asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
                asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
            } else if (P.sync_mode == 2) {
                asm volatile("fence.acq_rel.cluster;" ::: "memory");
                asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
                asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
            } else if (P.sync_mode == 3) {
                asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
                asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
            } else if (P.sync_mode == 4) {
                __threadfence();
                asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
                asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
            } else if (P.sync_mode == 5) {
                __syncthreads();
                asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
                asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
            } else {
                __threadfence_block();
                asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
                asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
                __syncthreads();
            }
        } else {
            __syncthreads();
        }
    };
    auto exchange2 = [&](const Tile<NSEG, PY>& a, const Tile<NSEG, PY>& b, double (&a_ab)[NSEG], double (&b_ab)[NSEG],
                         double (&a_be)[NSEG], double (&b_be)[NSEG]) {
        const size_t boff = (size_t)xbuf * NW * 4 * LX;
        double* mine = strips + boff + (size_t)warp * 4 * LX;
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            mine[0 * LX + 32 * q + lane] = a.a[0][q];
            mine[1 * LX + 32 * q + lane] = b.a[0][q];
            mine[2 * LX + 32 * q + lane] = a.a[PY - 1][q];
            mine[3 * LX + 32 * q + lane] = b.a[PY - 1][q];
        }
        edge_sync();
        const double* up = (warp == 0) ? strips_up + boff + (size_t)(NW - 1) * 4 * LX : strips + boff + (size_t)(warp - 1) * 4 * LX;
        const double* dn = (warp == NW - 1) ? strips_dn + boff : strips + boff + (size_t)(warp + 1) * 4 * LX;
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            a_ab[q] = up[2 * LX + 32 * q + lane];
            b_ab[q] = up[3 * LX + 32 * q + lane];
            a_be[q] = dn[0 * LX + 32 * q + lane];
            b_be[q] = dn[1 * LX + 32 * q + lane];
        }
        xbuf ^= 1;
    };
    auto exchange1 = [&](const Tile<NSEG, PY>& a, double (&a_ab)[NSEG], double (&a_be)[NSEG]) {
        const size_t boff = (size_t)xbuf * NW * 4 * LX;
        double* mine = strips + boff + (size_t)warp * 4 * LX;
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            mine[0 * LX + 32 * q + lane] = a.a[0][q];
            mine[2 * LX + 32 * q + lane] = a.a[PY - 1][q];
        }
        edge_sync();
        const double* up = (warp == 0) ? strips_up + boff + (size_t)(NW - 1) * 4 * LX : strips + boff + (size_t)(warp - 1) * 4 * LX;
        const double* dn = (warp == NW - 1) ? strips_dn + boff : strips + boff + (size_t)(warp + 1) * 4 * LX;
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            a_ab[q] = up[2 * LX + 32 * q + lane];
            a_be[q] = dn[0 * LX + 32 * q + lane];
        }
        xbuf ^= 1;
    };

    // SSH: the tables of slices tau and tau+1 (rows of this CTA + the y-table row above them) stay in shared memory
    const double2* txc = nullptr; const double2* tyc = nullptr; const double2* hyc = nullptr;
    const double2* txn = nullptr; const double2* tyn = nullptr; const double2* hyn = nullptr;
    if constexpr (SSH) {
        // per slice: x-table [RB][LX], y-table [RB][LX], y-halo row [LX]
        double2* tb = reinterpret_cast<double2*>(sp);
        const int RBL = PY * NW * LX;
        const int taup_w = (tau == L - 1) ? 0 : tau + 1;
        const size_t cta_off = (size_t)(yb * NW * PY) * LX;
        const size_t halo_row = (size_t)((yb * NW * PY + P.Ly - 1) % P.Ly) * LX;
        for (int s = 0; s < 2; ++s) {
            const double2* src = P.tab + (size_t)(s ? taup_w : tau) * 2 * N;
            double2* dst = tb + (size_t)s * (2 * RBL + LX);
            for (int i = tid; i < RBL; i += T) {
                dst[i] = src[cta_off + i];
                dst[RBL + i] = src[N + cta_off + i];
            }
            for (int i = tid; i < LX; i += T) dst[2 * RBL + i] = src[N + halo_row + i];
        }
        __syncthreads();
        const size_t wo = (size_t)warp * PY * LX;
        const double2* t0 = tb;
        const double2* t1n = tb + (2 * RBL + LX);
        txc = t0 + wo; tyc = t0 + RBL + wo; hyc = (warp == 0) ? t0 + 2 * RBL : t0 + RBL + wo - LX;
        txn = t1n + wo; tyn = t1n + RBL + wo; hyn = (warp == 0) ? t1n + 2 * RBL : t1n + RBL + wo - LX;
    }

    // antiperiodic wrap: the sign flips on GLOBAL slice 0 (re-set per slice in multi-slice mode)
    bool wrap_c = (P.tau0 + tau == 0);
    bool wrap_n = (P.tau0 + tau + 1 == P.Lglob);
    // t1 <- (M^T M v)(tau) from t1 = v(tau-1), vc = v(tau) and v(tau+1) delivered by load_next (after the first sweeps)
    auto apply_A = [&](const Tile<NSEG, PY>& vc, auto&& load_next) {
#pragma unroll
        for (int a = 0; a < PY; ++a)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                t1.a[a][q] = DC(a, q) * t1.a[a][q];
                t2.a[a][q] = DN(a, q) * vc.a[a][q];
            }
        if constexpr (SSH) {
            g0_tab(t1, txc, lane); g0_tab(t2, txn, lane);
            g1_tab(t1, txc, lane); g1_tab(t2, txn, lane);
            g2_tab(t1, tyc, lane); g2_tab(t2, tyn, lane);
        } else {
            g0_x_even(t1, P.c0, P.s0); g0_x_even(t2, P.c0, P.s0);
            g1_x_odd(t1, P.c1, P.s1, lane); g1_x_odd(t2, P.c1, P.s1, lane);
            g2_y_even(t1, P.c2, P.s2); g2_y_even(t2, P.c2, P.s2);
        }
        {
            double a1[NSEG], a2[NSEG], b1[NSEG], b2[NSEG];
            exchange2(t1, t2, a1, a2, b1, b2);
            if constexpr (SSH) { g3_tab(t1, tyc, hyc, lane, a1, b1); g3_tab(t2, tyn, hyn, lane, a2, b2); }
            else { g3_y_odd(t1, P.c3, P.s3, a1, b1); g3_y_odd(t2, P.c3, P.s3, a2, b2); }
        }
        double vn[PY][NSEG];
        load_next(vn);
#pragma unroll
        for (int a = 0; a < PY; ++a)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const double wc = wrap_c ? (vc.a[a][q] + t1.a[a][q]) : (vc.a[a][q] - t1.a[a][q]);
                const double wn = wrap_n ? (vn[a][q] + t2.a[a][q]) : (vn[a][q] - t2.a[a][q]);
                t1.a[a][q] = wc;
                t2.a[a][q] = wn;
            }
        {
            double ab[NSEG], be[NSEG];
            exchange1(t2, ab, be);
            if constexpr (SSH) g3_tab(t2, tyn, hyn, lane, ab, be);
            else g3_y_odd(t2, P.c3, P.s3, ab, be);
        }
        if constexpr (SSH) { g2_tab(t2, tyn, lane); g1_tab(t2, txn, lane); g0_tab(t2, txn, lane); }
        else { g2_y_even(t2, P.c2, P.s2); g1_x_odd(t2, P.c1, P.s1, lane); g0_x_even(t2, P.c0, P.s0); }
#pragma unroll
        for (int a = 0; a < PY; ++a)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const double du = DN(a, q) * t2.a[a][q];
                t1.a[a][q] = wrap_n ? (t1.a[a][q] + du) : (t1.a[a][q] - du);
            }
    };

    // ---- rows of self-validating words ---------------------------------------------------------------------------------
    const size_t rstride = (size_t)(P.Lmax + 2) * N * 2;      // words per parity
    auto row_words = [&](unsigned long long* b, unsigned int tag, int r) { return b + (size_t)(tag & 1u) * rstride + (size_t)r * N * 2; };
    const int r_lo = (tau > 0) ? 2 + tau - 1 : (multi ? 0 : 2 + L - 1);
    const int r_hi = (taul < L - 1) ? 2 + taul + 1 : (multi ? 1 : 2);
    bool alive = true;
    // the row of own slice ts; to_left / to_right: it is also the neighbour GPU's halo row
    auto publish_slice = [&](const Tile<NSEG, PY>& v, unsigned int tag, int ts, bool to_left, bool to_right) {
        unsigned long long* own = row_words(P.rows, tag, 2 + ts);
        unsigned long long* pl = to_left ? row_words(P.left_rows, tag, 1) : nullptr;
        unsigned long long* pr2 = to_right ? row_words(P.right_rows, tag, 0) : nullptr;
#pragma unroll
        for (int a = 0; a < PY; ++a)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const size_t e = eidx(a, q);
                if (to_left) ll::push(pl + 2 * e, v.a[a][q], tag);    // remote stores first: they travel furthest
                if (to_right) ll::push(pr2 + 2 * e, v.a[a][q], tag);
                ll::push(own + 2 * e, v.a[a][q], tag);
            }
    };
    auto publish = [&](const Tile<NSEG, PY>& v, unsigned int tag) { publish_slice(v, tag, tau, first, last); };
    auto read_row = [&](int r, unsigned int tag, double (&out)[PY][NSEG]) {
        const unsigned long long* src = row_words(P.rows, tag, r);
        unsigned long long wa[PY][NSEG], wb[PY][NSEG];
        unsigned int spins = 0;
        bool all;
        do {
            all = true;
#pragma unroll
            for (int a = 0; a < PY; ++a)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) ll::ld2(src + 2 * eidx(a, q), wa[a][q], wb[a][q]);
#pragma unroll
            for (int a = 0; a < PY; ++a)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) all = all && ll::tag_ok(wa[a][q], wb[a][q], tag);
        } while (!all && !dead && ++spins < kSpinLimit);
        if (!all) { alive = false; dead = true; st_volatile_u32(P.abort_word, 1u); }
#pragma unroll
        for (int a = 0; a < PY; ++a)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) out[a][q] = ll::unpack(wa[a][q], wb[a][q]);
    };
    // both neighbour rows of one publication: prefetch_rows starts 16-byte asynchronous copies into shared memory (no
    // registers held while the update and the post run), fetch_rows checks the tags and falls back to polling
    auto prefetch_rows = [&](unsigned int tag) {
        if constexpr (PF) {
            const unsigned long long* slo = row_words(P.rows, tag, r_lo);
            const unsigned long long* shi = row_words(P.rows, tag, r_hi);
#pragma unroll
            for (int a = 0; a < PY; ++a)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) {
                    const unsigned int d0 = (unsigned int)__cvta_generic_to_shared(pre + (size_t)(a * NSEG + q) * T + tid);
                    const unsigned int d1 = (unsigned int)__cvta_generic_to_shared(pre + (size_t)((PY + a) * NSEG + q) * T + tid);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0), "l"(slo + 2 * eidx(a, q)) : "memory");
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d1), "l"(shi + 2 * eidx(a, q)) : "memory");
                }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    };
    auto fetch_rows = [&](unsigned int tag, double (&lo)[PY][NSEG], double (&hi)[PY][NSEG]) {
        bool got = false;
        if constexpr (PF) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            got = true;
#pragma unroll
            for (int a = 0; a < PY; ++a)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) {
                    const ulonglong2 u = pre[(size_t)(a * NSEG + q) * T + tid];
                    const ulonglong2 v = pre[(size_t)((PY + a) * NSEG + q) * T + tid];
                    got = got && ll::tag_ok(u.x, u.y, tag) && ll::tag_ok(v.x, v.y, tag);
                    lo[a][q] = ll::unpack(u.x, u.y);
                    hi[a][q] = ll::unpack(v.x, v.y);
                }
        }
        if (!got) {     // not there yet (or no prefetch): poll the rows themselves, all loads of a round in flight together
            const unsigned long long* slo = row_words(P.rows, tag, r_lo);
            const unsigned long long* shi = row_words(P.rows, tag, r_hi);
            unsigned long long wa[PY][NSEG], wb[PY][NSEG], wc[PY][NSEG], wd[PY][NSEG];
            unsigned int spins = 0;
            bool all;
            do {
                all = true;
#pragma unroll
                for (int a = 0; a < PY; ++a)
#pragma unroll
                    for (int q = 0; q < NSEG; ++q) {
                        ll::ld2(slo + 2 * eidx(a, q), wa[a][q], wb[a][q]);
                        ll::ld2(shi + 2 * eidx(a, q), wc[a][q], wd[a][q]);
                    }
#pragma unroll
                for (int a = 0; a < PY; ++a)
#pragma unroll
                    for (int q = 0; q < NSEG; ++q) all = all && ll::tag_ok(wa[a][q], wb[a][q], tag) && ll::tag_ok(wc[a][q], wd[a][q], tag);
            } while (!all && !dead && ++spins < kSpinLimit);
            if (!all) { alive = false; dead = true; st_volatile_u32(P.abort_word, 1u); }
#pragma unroll
            for (int a = 0; a < PY; ++a)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) { lo[a][q] = ll::unpack(wa[a][q], wb[a][q]); hi[a][q] = ll::unpack(wc[a][q], wd[a][q]); }
        }
    };
    // sums over the CTA of two per-thread values, posted into this CTA's slot with the tag of reduction n
    auto post = [&](double v0, double v1, unsigned int tag) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v0 += __shfl_xor_sync(0xffffffffu, v0, o);
            v1 += __shfl_xor_sync(0xffffffffu, v1, o);
        }
        if (lane == 0) { red[0][warp] = v0; red[1][warp] = v1; }
        __syncthreads();
        if (tid == 0) {
            double b0 = 0.0, b1 = 0.0;
            for (int k = 0; k < NW; ++k) { b0 += red[0][k]; b1 += red[1][k]; }
            unsigned long long* slot = P.part + ((size_t)(tag & 1u) * P.maxcta + blockIdx.x) * 4;
            ll::push(slot, b0, tag);
            ll::push(slot + 2, b1, tag);
        }
    };
    // message j of the reducer: alpha_j, beta_j, flags.  Warp 0 polls this CTA's copy (all lanes read the same words), the
    // other warps pick the values up from shared memory.  Returns the flags (bit 0 abort, bit 1 stop), 1 on a timeout.
    auto wait_coeff = [&](unsigned int tag, double& alpha, double& beta) -> int {
        if (warp == 0) {
            const unsigned long long* bc = P.bcast + ((size_t)(tag & 1u) * kBcastCopies + (blockIdx.x % kBcastCopies)) * kBcastStride;
            unsigned long long a0, a1, b0, b1, c0, c1;
            unsigned int spins = 0;
            bool ok;
            do {
                ll::ld2(bc, a0, a1);
                ll::ld2(bc + 2, b0, b1);
                ll::ld2(bc + 4, c0, c1);
                ok = ll::tag_ok(a0, a1, tag) && ll::tag_ok(b0, b1, tag) && ll::tag_ok(c0, c1, tag);
            } while (!ok && ++spins < 8u * kSpinLimit);
            if (lane == 0) { cf[0] = ll::unpack(a0, a1); cf[1] = ll::unpack(b0, b1); cf[2] = ok ? ll::unpack(c0, c1) : 1.0; }
        }
        __syncthreads();
        alpha = cf[0];
        beta = cf[1];
        return (int)cf[2];
    };

    // ---- multi-slice mode ------------------------------------------------------------------------------------------------
    if constexpr (MS) {
        const size_t vs = (size_t)P.Lmax * N;
        double* Rg = P.state;           // r, w, s, z, q, p of the own slices: touched by their owner thread only
        double* Wg = Rg + vs;
        double* Sg = Wg + vs;
        double* Zg = Sg + vs;
        double* Qg = Zg + vs;
        double* Pg = Qg + vs;
        double* Xg = P.x;
        const unsigned int base = P.base;
        auto rowof = [&](int ts) -> long long { return (long long)ts * N; };
        // D(ts) and D(ts+1) of own slice ts into the register tiles apply_A reads; wrap flags of that slice
        auto enter_slice = [&](int ts) {
            const long long rk = rowof(ts);
            const long long rdn = P.d_halo ? rk + N : rowof((ts == L - 1) ? 0 : ts + 1);
#pragma unroll
            for (int a = 0; a < PY; ++a)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) {
                    Dcr.a[a][q] = P.D[rk + eidx(a, q)];
                    Dnr.a[a][q] = P.D[rdn + eidx(a, q)];
                }
            wrap_c = (P.tau0 + ts == 0);
            wrap_n = (P.tau0 + ts + 1 == P.Lglob);
        };
        auto load_tile = [&](const double* src, long long rk, Tile<NSEG, PY>& t) {
#pragma unroll
            for (int a = 0; a < PY; ++a)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) t.a[a][q] = __ldcg(src + rk + eidx(a, q));
        };
        // ---- set-up ------------------------------------------------------------------------------------------------------
        for (int k = 0; k < ns; ++k) {
            const long long rk = rowof(tau + k);
#pragma unroll
            for (int a = 0; a < PY; ++a)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) {
                    const size_t e = rk + eidx(a, q);
                    Rg[e] = P.r0[e];
                    Sg[e] = 0.0; Zg[e] = 0.0; Pg[e] = 0.0;
                    if (!P.x0_given) Xg[e] = 0.0;
                }
        }
#pragma unroll
        for (int a = 0; a < PY; ++a)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) { gzl[gidx(a, q)] = 0.0; gzh[gidx(a, q)] = 0.0; }
        if (first) { load_tile(P.r0, rowof(0), w); publish_slice(w, base + 1u, 0, true, false); }          // r_0 for the neighbour GPUs
        if (last) { load_tile(P.r0, rowof(L - 1), w); publish_slice(w, base + 1u, L - 1, false, true); }
        double accg = 0.0, accd = 0.0;
        for (int k = 0; k < ns; ++k) {
            const int ts = tau + k;
            enter_slice(ts);
            Tile<NSEG, PY> vc;
            load_tile(P.r0, rowof(ts), vc);
            if (first && ts == 0) {
                double hrow[PY][NSEG];
                read_row(0, base + 1u, hrow);
#pragma unroll
                for (int a = 0; a < PY; ++a)
#pragma unroll
                    for (int q = 0; q < NSEG; ++q) t1.a[a][q] = hrow[a][q];
            } else {
                load_tile(P.r0, rowof(ts == 0 ? L - 1 : ts - 1), t1);
            }
            apply_A(vc, [&](double (&vn)[PY][NSEG]) {
                if (last && ts == L - 1) read_row(1, base + 1u, vn);
                else {
                    const long long rn = rowof(ts == L - 1 ? 0 : ts + 1);
#pragma unroll
                    for (int a = 0; a < PY; ++a)
#pragma unroll
                        for (int q = 0; q < NSEG; ++q) vn[a][q] = P.r0[rn + eidx(a, q)];
                }
            });
#pragma unroll
            for (int a = 0; a < PY; ++a)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) {
                    Wg[rowof(ts) + eidx(a, q)] = t1.a[a][q];
                    accg = fma(vc.a[a][q], vc.a[a][q], accg);
                    accd = fma(t1.a[a][q], vc.a[a][q], accd);
                }
            if (k == 0 || k == ns - 1) publish_slice(t1, base + 2u, ts, first && k == 0, last && k == ns - 1);   // w_0 of the edge slices
        }
        post(accg, accd, base + 1u);
        {
            double lo[PY][NSEG], hi[PY][NSEG];
            read_row(r_lo, base + 2u, lo);
            read_row(r_hi, base + 2u, hi);
#pragma unroll
            for (int a = 0; a < PY; ++a)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) { gwl[gidx(a, q)] = lo[a][q]; gwh[gidx(a, q)] = hi[a][q]; }
        }
        // q = A w for every own slice from the w of its neighbours: own slices from global memory, outer ones from the ghosts
        auto products = [&](unsigned int tag) {
            for (int k = 0; k < ns; ++k) {
                const int ts = tau + k;
                enter_slice(ts);
                Tile<NSEG, PY> vc;
                load_tile(Wg, rowof(ts), vc);
                if (k == 0) {
#pragma unroll
                    for (int a = 0; a < PY; ++a)
#pragma unroll
                        for (int q = 0; q < NSEG; ++q) t1.a[a][q] = gwl[gidx(a, q)];
                } else {
                    load_tile(Wg, rowof(ts - 1), t1);
                }
                apply_A(vc, [&](double (&vn)[PY][NSEG]) {
#pragma unroll
                    for (int a = 0; a < PY; ++a)
#pragma unroll
                        for (int q = 0; q < NSEG; ++q) vn[a][q] = (k == ns - 1) ? gwh[gidx(a, q)] : __ldcg(Wg + rowof(ts + 1) + eidx(a, q));
                });
#pragma unroll
                for (int a = 0; a < PY; ++a)
#pragma unroll
                    for (int q = 0; q < NSEG; ++q) __stcg(Qg + rowof(ts) + eidx(a, q), t1.a[a][q]);
                if (k == 0 || k == ns - 1) publish_slice(t1, tag, ts, first && k == 0, last && k == ns - 1);
            }
        };
        products(base + 3u);                                                         // q_0
        // ---- iterations --------------------------------------------------------------------------------------------------
        double alpha_prev = 0.0;
        for (unsigned int j = 0;; ++j) {
            double alpha, beta;
            const int flags = wait_coeff(base + 1u + j, alpha, beta);
            if (flags & 1) { alive = false; break; }
            if (flags & 2) break;
            prefetch_rows(base + 3u + j);
            accg = 0.0; accd = 0.0;
            for (int k = 0; k < ns; ++k) {
                const long long rk = rowof(tau + k);
                // all 7 x (PY x NSEG) loads of the slice in flight together: the phase is bound by memory-level parallelism
                double rj[PY][NSEG], wj[PY][NSEG], sj[PY][NSEG], zj[PY][NSEG], qj[PY][NSEG], pj[PY][NSEG], xj[PY][NSEG];
#pragma unroll
                for (int a = 0; a < PY; ++a)
#pragma unroll
                    for (int q = 0; q < NSEG; ++q) {
                        const size_t e = rk + eidx(a, q);
                        rj[a][q] = __ldcg(Rg + e); wj[a][q] = __ldcg(Wg + e); sj[a][q] = __ldcg(Sg + e); zj[a][q] = __ldcg(Zg + e);
                        qj[a][q] = __ldcg(Qg + e); pj[a][q] = __ldcg(Pg + e); xj[a][q] = __ldcg(Xg + e);
                    }
#pragma unroll
                for (int a = 0; a < PY; ++a)
#pragma unroll
                    for (int q = 0; q < NSEG; ++q) {
                        const size_t e = rk + eidx(a, q);
                        if (j > 0) __stcg(Xg + e, fma(alpha_prev, pj[a][q], xj[a][q]));   // commit x_j = x_{j-1} + alpha_{j-1} p_{j-1}
                        const double zv = fma(beta, zj[a][q], qj[a][q]);
                        const double sv = fma(beta, sj[a][q], wj[a][q]);
                        __stcg(Zg + e, zv);
                        __stcg(Sg + e, sv);
                        __stcg(Pg + e, fma(beta, pj[a][q], rj[a][q]));
                        const double rv = fma(-alpha, sv, rj[a][q]);
                        const double wv = fma(-alpha, zv, wj[a][q]);
                        __stcg(Rg + e, rv);
                        __stcg(Wg + e, wv);
                        accg = fma(rv, rv, accg);
                        accd = fma(wv, rv, accd);
                    }
            }
            alpha_prev = alpha;
            post(accg, accd, base + 2u + j);
            {
                double lo[PY][NSEG], hi[PY][NSEG];
                fetch_rows(base + 3u + j, lo, hi);
#pragma unroll
                for (int a = 0; a < PY; ++a)
#pragma unroll
                    for (int q = 0; q < NSEG; ++q) {
                        const size_t g = gidx(a, q);
                        const double zl = fma(beta, gzl[g], lo[a][q]);
                        const double zh = fma(beta, gzh[g], hi[a][q]);
                        gzl[g] = zl;
                        gzh[g] = zh;
                        gwl[g] = fma(-alpha, zl, gwl[g]);
                        gwh[g] = fma(-alpha, zh, gwh[g]);
                    }
            }
            products(base + 4u + j);                                                 // q_{j+1}
        }
        if (!alive && tid == 0) st_volatile_u32(P.abort_word, 1u);
        if (YS > 1) cluster.sync();
        return;
    }

    // ---- set-up: r_0, w_0 = A r_0, q_0 = A w_0 -------------------------------------------------------------------------
    const unsigned int base = P.base;
#pragma unroll
    for (int a = 0; a < PY; ++a)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const size_t e = eidx(a, q);
            RR(a, q) = P.r0[row + e];
            if constexpr (XPG) { if (!P.x0_given) xgl[e] = 0.0; pgl[e] = 0.0; }
            else { X(a, q) = P.x0_given ? P.x[row + e] : 0.0; PP(a, q) = 0.0; }
            SS(a, q) = 0.0;
            ZZ(a, q) = 0.0;
            DC(a, q) = SSH ? P.D[e] : P.D[row + e];
            DN(a, q) = SSH ? P.D[e] : P.D[rowDn + e];
            gzl[gidx(a, q)] = 0.0;
            gzh[gidx(a, q)] = 0.0;
        }
    {
        Tile<NSEG, PY> vc;
#pragma unroll
        for (int a = 0; a < PY; ++a)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) vc.a[a][q] = RR(a, q);
        if (first || last) publish(vc, base + 1u);              // r_0 of the slab's edge slices for the neighbour GPUs
        if (first) {
            double hrow[PY][NSEG];
            read_row(0, base + 1u, hrow);
#pragma unroll
            for (int a = 0; a < PY; ++a)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) t1.a[a][q] = hrow[a][q];
        } else {
#pragma unroll
            for (int a = 0; a < PY; ++a)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) t1.a[a][q] = P.r0[rowm + eidx(a, q)];
        }
        apply_A(vc, [&](double (&vn)[PY][NSEG]) {
            if (last) read_row(1, base + 1u, vn);
            else {
#pragma unroll
                for (int a = 0; a < PY; ++a)
#pragma unroll
                    for (int q = 0; q < NSEG; ++q) vn[a][q] = P.r0[rowp + eidx(a, q)];
            }
        });
        double accg = 0.0, accd = 0.0;
#pragma unroll
        for (int a = 0; a < PY; ++a)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                w.a[a][q] = t1.a[a][q];
                accg = fma(vc.a[a][q], vc.a[a][q], accg);
                accd = fma(w.a[a][q], vc.a[a][q], accd);
            }
        publish(w, base + 2u);                                   // w_0: the neighbours' ghosts start from it
        post(accg, accd, base + 1u);                             // reduction 0
        double lo[PY][NSEG], hi[PY][NSEG];
        {
            // (no prefetch here: the rows are being written right now)
            read_row(r_lo, base + 2u, lo);
            read_row(r_hi, base + 2u, hi);
        }
#pragma unroll
        for (int a = 0; a < PY; ++a)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                gwl[gidx(a, q)] = lo[a][q];
                gwh[gidx(a, q)] = hi[a][q];
                t1.a[a][q] = lo[a][q];
            }
        apply_A(w, [&](double (&vn)[PY][NSEG]) {
#pragma unroll
            for (int a = 0; a < PY; ++a)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) vn[a][q] = hi[a][q];
        });
        publish(t1, base + 3u);                                  // q_0
    }
    // ---- iterations ----------------------------------------------------------------------------------------------------
    unsigned long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool prof = (P.prof != nullptr);
    long long tk = prof ? clock64() : 0;
    auto tick = [&](int k) { if (prof) { const long long now = clock64(); pc[k] += (unsigned long long)(now - tk); tk = now; } };
    double alpha_prev = 0.0;
    for (unsigned int j = 0;; ++j) {
        double alpha, beta;
        const int flags = wait_coeff(base + 1u + j, alpha, beta);
        if (flags & 1) { alive = false; break; }                                 // aborted (a timeout somewhere)
        if (flags & 2) break;                                                    // j-1 iterations were enough: x holds x_{j-1}
        tick(0);
        prefetch_rows(base + 3u + j);                                            // q_j of both neighbours -> shared memory
        if (j > 0) {                                                             // commit x_j = x_{j-1} + alpha_{j-1} p_{j-1}
#pragma unroll
            for (int a = 0; a < PY; ++a)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) {
                    if constexpr (XPG) { const size_t e = eidx(a, q); xgl[e] = fma(alpha_prev, pgl[e], xgl[e]); }
                    else X(a, q) = fma(alpha_prev, PP(a, q), X(a, q));
                }
        }
        alpha_prev = alpha;
        double accg = 0.0, accd = 0.0;
#pragma unroll
        for (int a = 0; a < PY; ++a)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const double rj = RR(a, q);
                const double zv = fma(beta, ZZ(a, q), t1.a[a][q]);          // t1 holds q_j
                const double sv = fma(beta, SS(a, q), w.a[a][q]);
                ZZ(a, q) = zv;
                SS(a, q) = sv;
                const double rv = fma(-alpha, sv, rj);
                const double wv = fma(-alpha, zv, w.a[a][q]);
                RR(a, q) = rv;
                w.a[a][q] = wv;
                accg = fma(rv, rv, accg);
                accd = fma(wv, rv, accd);
                if constexpr (!XPG) PP(a, q) = fma(beta, PP(a, q), rj);     // x_{j+1} = x_j + alpha p_j waits for the stop check
                else t2.a[a][q] = rj;                                        // x, p live in L2: advanced after the post
            }
        tick(1);
        post(accg, accd, base + 2u + j);                                    // reduction j+1 travels from here on
        tick(2);
        if constexpr (XPG) {
#pragma unroll
            for (int a = 0; a < PY; ++a)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) {
                    const size_t e = eidx(a, q);
                    pgl[e] = fma(beta, pgl[e], t2.a[a][q]);
                }
        }
        // ghosts of the neighbour slices: z_j = q_j + beta z_{j-1}, w_{j+1} = w_j - alpha z_j with the owners' operations
        {
            double lo[PY][NSEG], hi[PY][NSEG];
            fetch_rows(base + 3u + j, lo, hi);
            tick(3);
#pragma unroll
            for (int a = 0; a < PY; ++a)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) {
                    const size_t g = gidx(a, q);
                    const double zl = fma(beta, gzl[g], lo[a][q]);
                    const double zh = fma(beta, gzh[g], hi[a][q]);
                    gzl[g] = zl;
                    gzh[g] = zh;
                    const double wl = fma(-alpha, zl, gwl[g]);
                    const double wh = fma(-alpha, zh, gwh[g]);
                    gwl[g] = wl;
                    gwh[g] = wh;
                    t1.a[a][q] = wl;
                }
        }
        apply_A(w, [&](double (&vn)[PY][NSEG]) {
#pragma unroll
            for (int a = 0; a < PY; ++a)
#pragma unroll
                for (int q = 0; q < NSEG; ++q) vn[a][q] = gwh[gidx(a, q)];
        });
        tick(4);
        publish(t1, base + 4u + j);                                         // q_{j+1}
        tick(5);
    }
    if (prof && tid == 0)
        for (int k = 0; k < 8; ++k) P.prof[(size_t)blockIdx.x * 8 + k] = pc[k];
    if constexpr (!XPG) {
#pragma unroll
        for (int a = 0; a < PY; ++a)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) P.x[row + eidx(a, q)] = X(a, q);
    }
    if (!alive && tid == 0) st_volatile_u32(P.abort_word, 1u);
    if (YS > 1) cluster.sync();     // no CTA leaves while a neighbour may still read its strips
}

// ---- host side -------------------------------------------------------------------------------------------------------
struct PipeLayout {   // offsets (bytes) inside the pipe region of an arena; identical on every rank
    size_t rows, part, bcast, mbox, abort_word, pg, ghost, state, hx, hx_flag, total;
};
PipeLayout pipe_layout(int N, int Lmax) {
    PipeLayout Y;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 255) & ~size_t(255); return at; };
    Y.rows = take(2ull * (Lmax + 2) * N * 2 * sizeof(unsigned long long));
    Y.part = take(2ull * Lmax * kMaxYS * 4 * sizeof(unsigned long long));
    Y.bcast = take(2ull * kBcastCopies * kBcastStride * sizeof(unsigned long long));
    Y.mbox = take(2ull * kMaxWorld * 4 * sizeof(unsigned long long));
    Y.abort_word = take(sizeof(unsigned int));
    Y.pg = take((size_t)Lmax * N * sizeof(double));
    Y.ghost = take(4ull * Lmax * N * sizeof(double));
    Y.state = take(6ull * Lmax * N * sizeof(double));
    Y.hx = take(2ull * 2 * N * 2 * sizeof(unsigned long long));      // halo exchange of the products: [2 parities][lo, hi][N][2 words]
    Y.hx_flag = take(sizeof(unsigned int));
    Y.total = o;
    return Y;
}

struct PipeConfig {
    int variant = 0;     // index into the instantiation table below, 0 = none applies
    int py = 0, nw = 0, ys = 1, spc = 1;
};

template <int NSEG, int PY, int MAXT, int MINB, int PLACE, bool SSH>
size_t pipe_smem(int N_cta_sites, int nw) {
    const int LX = 32 * NSEG;
    size_t d = 2ull * nw * 4 * LX;
    if (!(PLACE & PL_GHOST_GLOBAL)) d += 4ull * N_cta_sites;
    if (PLACE & PL_XP_SMEM) d += 2ull * N_cta_sites;
    if (PLACE & PL_SZ_SMEM) d += 2ull * N_cta_sites;
    if (PLACE & PL_D_SMEM) d += 2ull * N_cta_sites;
    if (PLACE & PL_R_SMEM) d += 1ull * N_cta_sites;
    if (PLACE & PL_PREFETCH) d += 4ull * N_cta_sites;
    size_t bytes = d * sizeof(double);
    if (SSH) bytes += 2ull * (2ull * N_cta_sites + LX) * sizeof(double2);
    return bytes;
}

// co-residency of the whole grid (the CTAs spin on each other): occupancy query with the cluster shape of the launch
template <typename K>
bool pipe_fits(elph_handle* h, K kern, int grid, int threads, int ys, size_t smem) {
    if (smem > h->smem_optin) return false;
    elph_enable_smem(h, kern);
    if (ys > 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(threads);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = ys; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return false; }
        return (long long)nclusters * ys >= grid;
    }
    int per_sm = 0;
    ELPH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    return (long long)per_sm * h->sm_count >= grid;
}

template <typename K>
bool pipe_launch(elph_handle* h, K kern, PipeParams& P, int grid, int threads, int ys, size_t smem) {
    if (!pipe_fits(h, kern, grid, threads, ys, smem)) return false;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute at[2];
    int na = 0;
    at[na].id = cudaLaunchAttributeCooperative;      // co-residency enforced by the driver as well
    at[na].val.cooperative = 1;
    ++na;
    if (ys > 1) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = ys; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = at;
    cfg.numAttrs = na;
    void* args[] = {&P};
    cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)kern, args);
    if (e != cudaSuccess && ys > 1) {
        // cooperative + cluster not accepted by this driver: the occupancy query above already guarantees co-residency
        cudaGetLastError();
        cfg.attrs = at + 1;
        cfg.numAttrs = na - 1;
        e = cudaLaunchKernelExC(&cfg, (const void*)kern, args);
    }
    ELPH_CUDA(e);
    h->launches++;
    return true;
}

// The variants: (id, lattice width / 32, rows per warp, max threads, min CTAs per SM, placement, SSH).
// ys = CTAs per slice, nw = warps per CTA: Ly = PY * nw * ys.
//   1: Lx 32, 8 rows / warp, <= 4 warps, all state in registers (<= 255 registers)
//   2: Lx 32, 4 rows / warp, <= 4 warps (slices cut in two or more), 3 CTAs / SM
//   7: Lx 32, 4 rows / warp, 8 warps: the whole slice in one CTA with 4 elements per thread (<= 128 registers)
//   3: Lx 64, 4 rows / warp, <= 4 warps, all state in registers
//   4: Lx 64, 4 rows / warp, x and p in shared memory, 3 CTAs / SM (64x64xL400 on 4 GPUs)
//   9: as 4 without the prefetch buffers: 64 KB of shared memory, three CTAs per SM
//   8: Lx 64, 2 rows / warp, 8 warps (<= 128 registers)
//   5: Lx 64, 4 rows / warp, 256 threads, s, z in shared memory, x, p and the ghosts in L2, 3 CTAs / SM (64x64xL400 on 2 GPUs)
//   6: Lx 32, 8 rows / warp, SSH (tables of two slices resident in shared memory), single GPU
//  10: Lx 64, multi-slice: 2..8 consecutive slices per CTA, all vectors in L2 (64x64xL400 on 1 or 2 GPUs);  11: the same for Lx 32
#define PIPE_VARIANTS(V)                                                    \
    V(1, 1, 8, 128, 2, PL_PREFETCH, false, false)                           \
    V(2, 1, 4, 128, 3, PL_PREFETCH, false, false)                           \
    V(7, 1, 4, 256, 2, PL_PREFETCH, false, false)                           \
    V(3, 2, 4, 128, 2, PL_PREFETCH, false, false)                           \
    V(4, 2, 4, 128, 3, PL_XP_SMEM | PL_PREFETCH, false, false)              \
    V(9, 2, 4, 128, 3, PL_XP_SMEM, false, false)                            \
    V(8, 2, 2, 256, 2, PL_PREFETCH, false, false)                           \
    V(5, 2, 4, 256, 3, PL_SZ_SMEM | PL_XP_GLOBAL | PL_GHOST_GLOBAL, false, false) \
    V(6, 1, 8, 128, 2, 0, true, false)                                      \
    V(10, 2, 4, 128, 2, PL_GHOST_GLOBAL | PL_PREFETCH, false, true)         \
    V(11, 1, 8, 128, 2, PL_GHOST_GLOBAL | PL_PREFETCH, false, true)

int pipe_variant_py(int variant) {
    switch (variant) {
#define V(id, nseg, py, maxt, minb, pl, ssh, ms) case id: return py;
        PIPE_VARIANTS(V)
#undef V
        default: return 0;
    }
}
int pipe_variant_maxw(int variant) {
    switch (variant) {
#define V(id, nseg, py, maxt, minb, pl, ssh, ms) case id: return maxt / 32;
        PIPE_VARIANTS(V)
#undef V
        default: return 0;
    }
}
bool pipe_variant_ms(int variant) {
    switch (variant) {
#define V(id, nseg, py, maxt, minb, pl, ssh, ms) case id: return ms;
        PIPE_VARIANTS(V)
#undef V
        default: return false;
    }
}

// spc: time slices per CTA (1 except in the multi-slice variants)
bool pipe_try(elph_handle* h, PipeParams& P, int variant, int nw, int ys, int spc, bool launch) {
    const int nchunk = (h->L + spc - 1) / spc;
    const int grid = (nchunk + 1) * ys, threads = nw * 32;
    const int Lx = (h->model == ELPH_MODEL_SSH) ? h->ssq.Lx : h->sq.Lx;
    auto go = [&](auto kern, size_t smem) {
        if (!launch) return pipe_fits(h, kern, grid, threads, ys, smem);
        return pipe_launch(h, kern, P, grid, threads, ys, smem);
    };
    const int nb = pipe_variant_py(variant) * nw * Lx;
    switch (variant) {
#define V(id, nseg, py, maxt, minb, pl, ssh, ms) \
        case id: return threads <= maxt && go(cgpipe_kernel<nseg, py, maxt, minb, pl, ssh, ms>, pipe_smem<nseg, py, maxt, minb, pl, ssh>(nb, nw));
        PIPE_VARIANTS(V)
#undef V
        default: return false;
    }
}

// candidate (variant, nw, ys, spc) tuples for this handle, best first
int pipe_candidates(const elph_handle* h, int (&cand)[64][4]) {
    const bool ssh = (h->model == ELPH_MODEL_SSH);
    if (!(ssh ? h->ssq.enabled : h->sq.enabled) || h->sq_disable) return 0;
    const int Lx = ssh ? h->ssq.Lx : h->sq.Lx, Ly = ssh ? h->ssq.Ly : h->sq.Ly;
    int n = 0;
    auto add = [&](int variant, int ys, int spc = 1) {
        const int py = pipe_variant_py(variant);
        if (h->pipe_variant > 0 && variant != h->pipe_variant) return;    // tuning key 13
        if (ys > kMaxYS || Ly % (py * ys)) return;
        const int nw = Ly / (py * ys);
        if (nw < 1 || nw > pipe_variant_maxw(variant) || (ys == 1 && nw < 2) || n >= 64) return;
        cand[n][0] = variant; cand[n][1] = nw; cand[n][2] = ys; cand[n][3] = spc; ++n;
    };
    // all placements for one way of cutting a slice, the one with most state in registers first
    auto add_ys = [&](int ys) {
        if (ssh) { if (Lx == 32 && !h->sharded && ys == 1) add(6, 1); return; }
        if (Lx == 32) { add(1, ys); add(7, ys); add(2, ys); }
        // (variant 5 -- most of the state in L2, one slice per CTA -- is slower than the launch-per-iteration path: 43 against
        // 25 us per iteration at 64x64xL200; it stays selectable with tuning key 13 only)
        else if (Lx == 64) { add(3, ys); add(8, ys); add(4, ys); add(9, ys); if (h->pipe_variant == 5) add(5, ys); }
    };
    if (h->pipe_ys > 0) add_ys(h->pipe_ys);    // tuning key 11 first; whatever does not fit falls through to the automatic order
    if (Lx == 32) for (int ys = 1; ys <= kMaxYS; ys *= 2) add_ys(ys);
    else { add_ys(4); add_ys(8); add_ys(2); add_ys(1); }
    // slabs whose state does not fit on chip: several slices per CTA, vectors streamed from L2 / HBM (tuning key 14 = slices per
    // CTA).  Measured at 64x64: 25 us per iteration for 200 slices, the same as the launch-per-iteration path of one GPU -- so it is
    // chosen only where that path does not exist (a lattice sharded over several GPUs) or on request (tuning key 13).
    const bool ms_ok = (h->sharded && h->p2p.world > 1) || h->pipe_variant == 10 || h->pipe_variant == 11;
    if (!ssh && ms_ok) {
        const int v = (Lx == 64) ? 10 : ((Lx == 32) ? 11 : 0), ys0 = (Lx == 64) ? 4 : 1;
        if (v)
            for (int spc = (h->pipe_spc > 0 ? h->pipe_spc : 2); spc <= (h->pipe_spc > 0 ? h->pipe_spc : 8); ++spc)
                add(v, (h->pipe_ys > 0) ? h->pipe_ys : ys0, spc);
    }
    return n;
}

bool pipe_select(elph_handle* h, PipeParams& P, PipeConfig& C, bool launch) {
    int cand[64][4];
    const int n = pipe_candidates(h, cand);
    for (int k = 0; k < n; ++k)
        if (pipe_try(h, P, cand[k][0], cand[k][1], cand[k][2], cand[k][3], false)) {
            C.variant = cand[k][0]; C.nw = cand[k][1]; C.ys = cand[k][2]; C.spc = cand[k][3];
            if (!launch) return true;
            P.ys = C.ys;
            P.spc = C.spc;
            P.maxcta = h->p2p.Lmax * kMaxYS;
            return pipe_try(h, P, C.variant, C.nw, C.ys, C.spc, true);
        }
    return false;
}

// Halo exchange of one halo'd vector ([-1] and [L] around the own slices) through peer memory: every rank pushes its first own
// slice into the left neighbour's hi row and its last own slice into the right neighbour's lo row as self-validating words,
// then unpacks its own two rows into the vector.  One launch, no NCCL call, no host synchronisation; the product that
// follows in the stream finds its halos in place.  Both directions always travel: completing exchange k then implies that
// both neighbours have consumed exchange k-1, which makes two buffers (tag parity) enough.
__global__ void __launch_bounds__(256) halo_exchange_kernel(double* v_own, int L, int N, unsigned long long* mine, unsigned long long* left,
                                                            unsigned long long* right, unsigned int tag, unsigned int* fail_flag) {
    const size_t side = (size_t)N * 2, par = (size_t)(tag & 1u) * 2 * side;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < N; e += gridDim.x * blockDim.x) {
        ll::push(left + par + side + 2 * (size_t)e, v_own[e], tag);                         // first slice -> left neighbour's hi row
        ll::push(right + par + 2 * (size_t)e, v_own[(size_t)(L - 1) * N + e], tag);          // last slice -> right neighbour's lo row
    }
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < N; e += gridDim.x * blockDim.x) {
        unsigned long long a0, a1, b0, b1;
        unsigned int spins = 0;
        bool ok;
        do {
            ll::ld2(mine + par + 2 * (size_t)e, a0, a1);
            ll::ld2(mine + par + side + 2 * (size_t)e, b0, b1);
            ok = ll::tag_ok(a0, a1, tag) && ll::tag_ok(b0, b1, tag);
        } while (!ok && ++spins < kSpinLimit);
        if (!ok) st_volatile_u32(fail_flag, 1u);
        v_own[-(ptrdiff_t)N + e] = ll::unpack(a0, a1);
        v_own[(size_t)L * N + e] = ll::unpack(b0, b1);
    }
}

}  // namespace

void elph_shard_halo_impl(elph_handle* h, double* v_own) {
    auto& A = h->p2p;
    ELPH_REQUIRE(A.arena && A.opened, ELPH_ERR_STATE, "elph_shard_p2p_open has not been called");
    const PipeLayout Y = pipe_layout(h->N, A.Lmax);
    auto at = [&](void* base, size_t off) { return reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(base) + A.pipe_off + off); };
    const int left = (A.rank + A.world - 1) % A.world, right = (A.rank + 1) % A.world;
    // the failure flag lives in mapped page-locked host memory: a timed-out kernel writes it directly, the host reads it at the
    // next call -- no copy, nothing in the stream between the exchange and the product
    unsigned int* flag = nullptr;
    ELPH_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&flag), h->h_hx_flag, 0));
    ELPH_REQUIRE(*h->h_hx_flag == 0u, ELPH_ERR_STATE, "halo exchange: a neighbour GPU did not deliver its slice in time (timeout)");
    const int threads = 256, blocks = std::min(h->sm_count, (h->N + threads - 1) / threads);
    halo_exchange_kernel<<<blocks, threads, 0, h->stream>>>(v_own, h->L, h->N, at(A.arena, Y.hx), at(A.peer[left], Y.hx), at(A.peer[right], Y.hx),
                                                            ++A.hx_seq, flag);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}

// the same exchange as arguments for a product kernel that performs it itself (mtm_square.cu, HALO); `advance` consumes a tag
HaloArgs elph_shard_halo_args(elph_handle* h, double* v_own, bool advance) {
    auto& A = h->p2p;
    ELPH_REQUIRE(A.arena && A.opened, ELPH_ERR_STATE, "elph_shard_p2p_open has not been called");
    const PipeLayout Y = pipe_layout(h->N, A.Lmax);
    auto at = [&](void* base, size_t off) { return reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(base) + A.pipe_off + off); };
    const int left = (A.rank + A.world - 1) % A.world, right = (A.rank + 1) % A.world;
    HaloArgs H;
    ELPH_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&H.fail), h->h_hx_flag, 0));
    ELPH_REQUIRE(*h->h_hx_flag == 0u, ELPH_ERR_STATE, "halo exchange: a neighbour GPU did not deliver its slice in time (timeout)");
    H.enabled = true;
    H.mine = at(A.arena, Y.hx); H.left = at(A.peer[left], Y.hx); H.right = at(A.peer[right], Y.hx);
    H.v_out = v_own;
    H.tag = advance ? ++A.hx_seq : A.hx_seq + 1;
    return H;
}

size_t elph_pipe_arena_bytes(int N, int Lmax) { return pipe_layout(N, Lmax).total; }

bool elph_cg_pipe_fits(elph_handle* h) {
    PipeParams P = {};
    PipeConfig C;
    return h->L >= 2 && pipe_select(h, P, C, false);
}

// One solve on the handle's arena (h->p2p: allocated and, for several GPUs, opened by cg_p2p.cu; the pipe region starts at
// p2p.pipe_off).  r0: initial residual, x: initial guess (x0_given) / solution.  Returns false if no variant applies.
bool elph_cg_pipe_run(elph_handle* h, const double* r0, double* x, bool x0_given, bool scalars_on_device, double tol, int64_t maxiter) {
    auto& A = h->p2p;
    if (h->L < 2 || !A.arena || !A.opened) return false;
    ELPH_REQUIRE(!A.pipe_failed, ELPH_ERR_STATE, "pipelined CG: an earlier solve on this arena timed out; re-open the peer arenas");
    const PipeLayout Y = pipe_layout(h->N, A.Lmax);
    auto at = [&](void* base, size_t off) { return reinterpret_cast<char*>(base) + A.pipe_off + off; };
    const int left = (A.rank + A.world - 1) % A.world, right = (A.rank + 1) % A.world;
    PipeParams P = {};
    const bool ssh = (h->model == ELPH_MODEL_SSH);
    P.D = h->d_D; P.tab = ssh ? h->ssq.d_tab : nullptr; P.r0 = r0; P.x = x;
    P.rows = reinterpret_cast<unsigned long long*>(at(A.arena, Y.rows));
    P.left_rows = reinterpret_cast<unsigned long long*>(at(A.peer[left], Y.rows));
    P.right_rows = reinterpret_cast<unsigned long long*>(at(A.peer[right], Y.rows));
    P.part = reinterpret_cast<unsigned long long*>(at(A.arena, Y.part));
    P.bcast = reinterpret_cast<unsigned long long*>(at(A.arena, Y.bcast));
    for (int q = 0; q < kMaxWorld; ++q) P.mbox[q] = (q < A.world) ? reinterpret_cast<unsigned long long*>(at(A.peer[q], Y.mbox)) : nullptr;
    P.abort_word = reinterpret_cast<unsigned int*>(at(A.arena, Y.abort_word));
    P.pg = reinterpret_cast<double*>(at(A.arena, Y.pg));
    P.ghost = reinterpret_cast<double*>(at(A.arena, Y.ghost));
    P.state = reinterpret_cast<double*>(at(A.arena, Y.state));
    P.spc = 1;
    P.sync_mode = h->pipe_sync_mode;
    P.S = h->d_cg;
    P.prof = h->pipe_prof ? h->pipe_prof_buf : nullptr;
    P.base = A.pipe_seq;
    P.L = h->L; P.Lmax = A.Lmax; P.Ly = ssh ? h->ssq.Ly : h->sq.Ly; P.rank = A.rank; P.world = A.world;
    P.tau0 = h->sharded ? h->shard_tau0 : 0;
    P.Lglob = h->sharded ? h->shard_Lglob : h->L;
    P.d_halo = h->sharded ? 1 : 0;
    P.x0_given = x0_given ? 1 : 0;
    P.c0 = h->sq.c[0]; P.s0 = h->sq.s[0]; P.c1 = h->sq.c[1]; P.s1 = h->sq.s[1];
    P.c2 = h->sq.c[2]; P.s2 = h->sq.s[2]; P.c3 = h->sq.c[3]; P.s3 = h->sq.s[3];
    cudaStream_t st = h->stream;
    if (!scalars_on_device) {
        CgScalars init = {};
        init.tol = tol; init.kappa_max = h->cg_kappa_max; init.maxiter = maxiter; init.normb = 0.0;   // |b| = |r0|
        *h->h_cg = init;
        ELPH_CUDA(cudaMemcpyAsync(h->d_cg, h->h_cg, sizeof(CgScalars), cudaMemcpyHostToDevice, st));
    }
    PipeConfig C;
    if (!pipe_select(h, P, C, true)) return false;
    ELPH_CUDA(cudaMemcpyAsync(h->h_cg, h->d_cg, sizeof(CgScalars), cudaMemcpyDeviceToHost, st));
    ELPH_CUDA(cudaStreamSynchronize(st));
    if (h->h_cg->done != 1) {
        A.pipe_failed = true;
        ELPH_REQUIRE(false, ELPH_ERR_STATE, "pipelined CG: a peer GPU or CTA did not deliver its words in time (timeout)");
    }
    // the CTAs run one iteration more than the count (they learn of the stop with the next message): publications use the
    // tags up to base + iter + 4, reductions up to base + iter + 2.  base + iter + 4 keeps the publication tags consecutive and
    // puts the first reduction of the next solve (base' + 1) on the other parity than the last one of this solve.
    A.pipe_seq += 4u + (unsigned int)h->h_cg->iter;
    h->pipe_last_variant = C.variant * 100 + C.ys * 10 + C.nw;
    h->pipe_last_spc = C.spc;
    return true;
}
