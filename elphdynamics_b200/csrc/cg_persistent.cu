// Persistent cooperative conjugate gradient on A = M^T M for periodic square lattices (unpreconditioned path).
//
// Same algorithm and stop rule as cg.cu / the reference (src/IterativeSolvers.jl:239-314).  Motivation (measured on
// B200, scripts/micro/xr_bench.cu): a dependent kernel launch costs >= 4 us whatever the kernel does, so a CG iteration
// made of two launches cannot go below ~10 us while its arithmetic at config B (32x32xL200, L2 resident) is ~2 us.
// Here the WHOLE solve is one cooperative launch: CTA tau owns time slice tau; x(tau), r(tau), p(tau) live in registers
// for the entire solve; an iteration costs two grid barriers (each fused with a scalar reduction) and no launch.
//
// Per iteration, CTA tau:
//   p_k(tau-1), p_k(tau+1) are rebuilt from the neighbours' r_k and p_{k-1} (global, written before the last barrier):
//   p_k = r_k + beta p_{k-1} -- two extra FMAs per point instead of a third barrier for the halo of p_k;
//   w(tau) = p(tau) -/+ K D(tau) p(tau-1),  w(tau+1) = p(tau+1) -/+ K D(tau+1) p(tau)   (both sweeps share one barrier)
//   z(tau) = w(tau) -/+ D(tau+1) K^T w(tau+1);   p.Ap = sum_tau |w(tau)|^2  -> grid barrier + fixed-order reduction
//   x += alpha p, r -= alpha z, |r|^2 -> grid barrier + fixed-order reduction -> stop rule (evaluated identically by every CTA)
// Reductions are index-ordered, so the result is bit-reproducible and independent of CTA scheduling.
#include "square_tiles.cuh"

#include <algorithm>

namespace {

using namespace sqt;

struct PcgParams {
    const double* __restrict__ D;    // Holstein: expnV [L][N]; SSH: exp(dtau mu) [N]
    const double2* __restrict__ tab; // SSH: (cosh, sinh) [L][2][N] in the tile layout of ssh_square.cu
    double* __restrict__ x;          // [L][N] in: initial guess, out: solution
    double* R;                       // [L][N] residual (in: r0), updated every iteration
    double* P0;                      // [L][N] p buffers (double-buffered); P1 must hold zeros on entry
    double* P1;
    double* partialA;                // [L][2] barrier slots (see grid_sum)
    double* partialB;                // [L][2]
    unsigned int* bar;               // monotonically increasing arrival counter (zero on entry)
    CgScalars* S;                    // in: normb, eps0, rdotz (= r0.r0), tol, kappa_max, maxiter ; out: iter, eps, done
    int L, Ly;
    long long vstride;               // batch: right-hand side k = blockIdx.y lives at x/R/P0/P1 + k*vstride,
    int pstride;                     //        partialA/B + k*pstride, bar + k, S + k
    double c0, s0, c1, s1, c2, s2, c3, s3;
};

// batch offsets (same field names in both parameter structs)
template <typename Params>
__device__ __forceinline__ void select_rhs(Params& P) {
    const size_t k = blockIdx.y;
    P.x += k * P.vstride; P.R += k * P.vstride; P.P0 += k * P.vstride; P.P1 += k * P.vstride;
    P.partialA += k * P.pstride; P.partialB += k * P.pstride; P.bar += k; P.S += k;
}

// Grid barrier fused with a sum over all CTAs; every CTA returns the same bits (fixed order: lane-strided, then tree).
// One arrival counter (monotonically increasing, target = seq * nb) polled by one thread per CTA, then the partials are
// read back.  (Tried and rejected, measured on B200: LL-style flagged slots -- value and sequence number in one 16-byte
// store, every CTA's warp 0 spinning on all nb slots -- cost 14.3 us/iteration at 200 CTAs against 8.1 us for the
// counter: O(nb^2) polling traffic in L2.)  The caller must have a __syncthreads between the CTA's global writes and
// this call.  `seq` numbers the barriers of the launch from 1.
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ double grid_sum(double block_value, double* partial, unsigned int* bar, unsigned int seq, int nb,
                                           double* bcast) {
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = block_value;
        // release-arrive without waiting for the atomic's return value
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
        const unsigned int target = seq * (unsigned int)nb;
        while (ld_acquire(bar) < target) {}
    }
    __syncthreads();
    // read-back with the whole CTA: one L2 round trip (scripts/micro/barrier_bench.cu: a single warp striding over 200
    // partials costs 1.35 us, the bare barrier 1.25 us), then a fixed-order tree: thread-strided, shuffle, warps in order
    double s = 0.0;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) s += __ldcg(partial + k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) bcast[threadIdx.x >> 5] = s;
    __syncthreads();
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += bcast[k];
    return t;   // same bits on every thread of every CTA
}

template <int NSEG, int PY>
__device__ __forceinline__ double tile_block_sum(double v, double* red, int lane, int warp, int nwarps) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int k = 0; k < nwarps; ++k) t += red[k];
    return t;   // same value on every thread
}

// MINB: resident CTAs per SM the register allocation must allow (3 x 128 threads -> 168 registers: 444 CTAs on 148 SMs,
// i.e. two right-hand sides of config B at once)
// LAT: 0 = square lattice, one (cosh, sinh) per colour; 1 = square lattice, SSH tables; 2 = honeycomb lattice 32 cells wide
// (NSEG = 2: the two orbitals of a cell, see the hc tiles of square_tiles.cuh)
template <int NSEG, int PY, int MAXT, int LAT, int MINB = 1>
__global__ void __launch_bounds__(MAXT, MINB) cg_persistent_kernel(PcgParams P) {
    constexpr int LX = 32 * NSEG;
    constexpr bool SSH = (LAT == 1);
    constexpr bool HC = (LAT == 2);
    static_assert(!HC || NSEG == 2, "honeycomb tiles hold the two orbitals of a cell");
    extern __shared__ __align__(16) double strips[];   // 2 x [nwarps][4][LX]; SSH: + the tables of slices tau and tau+1
    __shared__ double red[32];
    __shared__ double bcast[32];
    select_rhs(P);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int L = P.L, N = LX * P.Ly, nb = gridDim.x;
    const int tau = blockIdx.x;
    const int taum = (tau == 0) ? L - 1 : tau - 1;
    const int taup = (tau == L - 1) ? 0 : tau + 1;
    const size_t tile_off = (size_t)warp * PY * LX;
    auto eidx = [&](int r, int q) -> size_t { return HC ? tile_off + r * LX + 2 * lane + q : tile_off + r * LX + 32 * q + lane; };

    Tile<NSEG, PY> x, r, pprev, pc, Dc, Dn, t1, t2;
#pragma unroll
    for (int rr = 0; rr < PY; ++rr)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const size_t e = eidx(rr, q);
            x.a[rr][q] = P.x[(size_t)tau * N + e];
            r.a[rr][q] = P.R[(size_t)tau * N + e];
            pprev.a[rr][q] = 0.0;
            Dc.a[rr][q] = SSH ? P.D[e] : P.D[(size_t)tau * N + e];
            Dn.a[rr][q] = SSH ? P.D[e] : P.D[(size_t)taup * N + e];
        }
    // SSH: K(tau) and K(tau+1) stay in shared memory for the whole solve (the field is fixed during a solve)
    const double2* txc = nullptr; const double2* tyc = nullptr; const double2* hyc = nullptr;
    const double2* txn = nullptr; const double2* tyn = nullptr; const double2* hyn = nullptr;
    if constexpr (SSH) {
        double2* tabc = reinterpret_cast<double2*>(strips + 2ull * nwarps * 4 * LX);
        double2* tabn = tabc + 2 * N;
        for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) {
            tabc[i] = P.tab[(size_t)tau * 2 * N + i];
            tabn[i] = P.tab[(size_t)taup * 2 * N + i];
        }
        __syncthreads();
        const size_t halo_off = (size_t)((warp * PY + P.Ly - 1) % P.Ly) * LX;
        txc = tabc + tile_off; tyc = tabc + N + tile_off; hyc = tabc + N + halo_off;
        txn = tabn + tile_off; tyn = tabn + N + tile_off; hyn = tabn + N + halo_off;
    }
    const double normb = P.S->normb, eps0 = P.S->eps0, tol = P.S->tol, kappa_max = P.S->kappa_max;
    const long long maxiter = P.S->maxiter;
    double rdotr = P.S->rdotz, beta = 0.0, kmin = 0.0, eps = eps0;
    long long j = 0;
    int xbuf = 0;
    double* Pold = P.P1;   // holds zeros on entry: p_0 = r_0 + 0 * p_{-1}
    double* Pnew = P.P0;
    const bool wrap_c = (tau == 0);          // w(tau)   uses '+' on global slice 0
    const bool wrap_n = (taup == 0);         // w(tau+1) and the M^T closure use '+' when tau+1 wraps to 0

    while (j < maxiter) {
        ++j;
        // ---- p_k on slices tau-1 (folded straight into t1), tau, tau+1 (kept in t2 until the sweep) ----------
#pragma unroll
        for (int rr = 0; rr < PY; ++rr)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const size_t e = eidx(rr, q);
                const double pm = fma(beta, __ldcg(Pold + (size_t)taum * N + e), __ldcg(P.R + (size_t)taum * N + e));
                const double pcv = fma(beta, pprev.a[rr][q], r.a[rr][q]);
                pc.a[rr][q] = pcv;
                Pnew[(size_t)tau * N + e] = pcv;
                t1.a[rr][q] = Dc.a[rr][q] * pm;          // D(tau) p(tau-1)
                t2.a[rr][q] = Dn.a[rr][q] * pcv;         // D(tau+1) p(tau)
            }
        // ---- K sweep on both tiles (one barrier) ----------------------------------------------------------------
        if constexpr (HC) {
            hc0_cell(t1, P.c0, P.s0);
            hc0_cell(t2, P.c0, P.s0);
            hc1_lane(t1, P.c1, P.s1, lane);
            hc1_lane(t2, P.c1, P.s1, lane);
            double a1, b1, a2, b2;
            exchange_hc2(t1, t2, strips + (size_t)xbuf * nwarps * 4 * LX, warp, nwarps, lane, a1, b1, a2, b2);
            xbuf ^= 1;
            hc2_row(t1, P.c2, P.s2, a1, b1);
            hc2_row(t2, P.c2, P.s2, a2, b2);
        } else if constexpr (SSH) {
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
        if constexpr (!HC) {
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
        // w(tau) -> t1 ; w(tau+1) -> t2 ; partial p.Ap = |w(tau)|^2
        double acc = 0.0;
#pragma unroll
        for (int rr = 0; rr < PY; ++rr)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const size_t e = eidx(rr, q);
                const double pn = fma(beta, __ldcg(Pold + (size_t)taup * N + e), __ldcg(P.R + (size_t)taup * N + e));
                const double wc = wrap_c ? (pc.a[rr][q] + t1.a[rr][q]) : (pc.a[rr][q] - t1.a[rr][q]);
                const double wn = wrap_n ? (pn + t2.a[rr][q]) : (pn - t2.a[rr][q]);
                t1.a[rr][q] = wc;
                t2.a[rr][q] = wn;
                acc = fma(wc, wc, acc);
            }
        // ---- u = K^T w(tau+1) (in t2): g3, g2, g1, g0 ----------------------------------------------------------
        if constexpr (HC) {
            double ab, be;
            exchange_hc1(t2, strips + (size_t)xbuf * nwarps * 4 * LX, warp, nwarps, lane, ab, be);
            xbuf ^= 1;
            hc2_row(t2, P.c2, P.s2, ab, be);
            hc1_lane(t2, P.c1, P.s1, lane);
            hc0_cell(t2, P.c0, P.s0);
        } else {
            double ab[NSEG], be[NSEG];
            exchange_edges1(t2, strips + (size_t)xbuf * nwarps * 4 * LX, warp, nwarps, lane, ab, be);
            xbuf ^= 1;
            if constexpr (SSH) g3_tab(t2, tyn, hyn, lane, ab, be);
            else g3_y_odd(t2, P.c3, P.s3, ab, be);
        }
        if constexpr (HC) {
        } else if constexpr (SSH) {
            g2_tab(t2, tyn, lane);
            g1_tab(t2, txn, lane);
            g0_tab(t2, txn, lane);
        } else {
            g2_y_even(t2, P.c2, P.s2);
            g1_x_odd(t2, P.c1, P.s1, lane);
            g0_x_even(t2, P.c0, P.s0);
        }
        // ---- alpha ------------------------------------------------------------------------------------------------
        const double blockA = tile_block_sum<NSEG, PY>(acc, red, lane, warp, nwarps);
        const double pAp = grid_sum(blockA, P.partialA, P.bar, 2u * (unsigned int)j - 1u, nb, bcast);
        const double alpha = rdotr / pAp;
        // ---- x += alpha p ; r -= alpha z, z = w(tau) -/+ D(tau+1) u ------------------------------------------------
        double accr = 0.0;
#pragma unroll
        for (int rr = 0; rr < PY; ++rr)
#pragma unroll
            for (int q = 0; q < NSEG; ++q) {
                const double du = Dn.a[rr][q] * t2.a[rr][q];
                const double z = wrap_n ? (t1.a[rr][q] + du) : (t1.a[rr][q] - du);
                x.a[rr][q] = fma(alpha, pc.a[rr][q], x.a[rr][q]);
                const double rv = fma(-alpha, z, r.a[rr][q]);
                r.a[rr][q] = rv;
                P.R[(size_t)tau * N + eidx(rr, q)] = rv;
                accr = fma(rv, rv, accr);
                pprev.a[rr][q] = pc.a[rr][q];
            }
        const double blockB = tile_block_sum<NSEG, PY>(accr, red, lane, warp, nwarps);
        const double rrn = grid_sum(blockB, P.partialB, P.bar, 2u * (unsigned int)j, nb, bcast);
        // ---- stop rule (src/IterativeSolvers.jl:287-301), identical on every CTA -----------------------------------
        eps = sqrt(rrn) / normb;
        const double lg = log(2.0 * eps0 / eps);
        const double qq = 2.0 * (double)j / lg;
        const double kap = qq * qq;
        if (kap > kmin) kmin = kap;
        if (eps < tol || kmin > kappa_max) break;
        beta = rrn / rdotr;
        rdotr = rrn;
        double* tmp = Pold;
        Pold = Pnew;
        Pnew = tmp;
    }
#pragma unroll
    for (int rr = 0; rr < PY; ++rr)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) P.x[(size_t)tau * N + eidx(rr, q)] = x.a[rr][q];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        P.S->iter = j;
        P.S->eps = eps;
        P.S->kappa_min = kmin;
        P.S->done = 1;
    }
}

// Cooperative launch of nrhs independent solves, grid (L, g): as many right-hand sides at once as stay co-resident.
template <typename Params, typename Kern>
bool launch_groups(elph_handle* h, Kern kern, const Params& P, int threads, size_t smem, int nrhs) {
    elph_enable_smem(h, kern);
    int per_sm = 0;
    ELPH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    const int cap = (int)std::min<long long>((long long)per_sm * h->sm_count / h->L, 65535);   // all CTAs of a group co-resident
    if (cap < 1) return false;
    for (int k0 = 0; k0 < nrhs; k0 += cap) {
        const int g = std::min(cap, nrhs - k0);
        Params Q = P;
        Q.x += k0 * P.vstride; Q.R += k0 * P.vstride; Q.P0 += k0 * P.vstride; Q.P1 += k0 * P.vstride;
        Q.partialA += (size_t)k0 * P.pstride; Q.partialB += (size_t)k0 * P.pstride; Q.bar += k0; Q.S += k0;
        ELPH_CUDA(cudaMemsetAsync(Q.bar, 0, g * sizeof(unsigned int), h->stream));
        void* args[] = {&Q};
        ELPH_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(h->L, g), dim3(threads), args, smem, h->stream));
        h->launches++;
    }
    return true;
}

template <int NSEG, int PY, int MAXT, int LAT, int MINB = 1>
bool launch_persistent(elph_handle* h, const PcgParams& P, int nwarps, int nrhs) {
    constexpr int LX = 32 * NSEG;
    const size_t smem = 2ull * nwarps * 4 * LX * sizeof(double) + (LAT == 1 ? 2ull * 2 * h->N * sizeof(double2) : 0);
    if (smem > h->smem_optin) return false;
    return launch_groups(h, cg_persistent_kernel<NSEG, PY, MAXT, LAT, MINB>, P, nwarps * 32, smem, nrhs);
}

// ---- any lattice (Holstein): one time slice per CTA in shared memory ------------------------------------------------
// Same iteration as above for arbitrary bond lists (honeycomb, triangular, chains, small test lattices -- config D):
// the CTA keeps its two working slices D(tau) p(tau-1) and D(tau+1) p(tau) in shared memory together with the whole
// bond list (sites + cosh/sinh, loaded once per solve), sweeps the colours there (one __syncthreads per colour, both
// slices per pass), and holds x, r, p, D of its slice in registers (element i = tid + k*T).
struct GcgParams {
    const double* __restrict__ D;     // expnV [L][N]
    double* __restrict__ x;
    double* R;
    double* P0;
    double* P1;
    double* partialA;
    double* partialB;
    unsigned int* bar;
    CgScalars* S;
    const int2* __restrict__ bonds;   // [Nb] colour-ordered
    const double2* __restrict__ cs;   // [Nb]
    const int* __restrict__ goff;     // [ngroups+1]
    int N, L, Nb, ngroups;
    long long vstride;
    int pstride;
};

__device__ __forceinline__ double block_sum_all(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
    return t;   // same value on every thread
}

template <int EPT, int MAXT>
__global__ void __launch_bounds__(MAXT) cg_persistent_generic_kernel(GcgParams P) {
    extern __shared__ __align__(16) unsigned char gsm[];
    __shared__ double red[32];
    __shared__ double bcast[32];
    select_rhs(P);
    const int N = P.N, L = P.L, nb = gridDim.x, T = blockDim.x, tid = threadIdx.x;
    double* A1 = reinterpret_cast<double*>(gsm);
    double* A2 = A1 + N;
    double2* scs = reinterpret_cast<double2*>(A2 + N);   // 2N doubles = 16N bytes: 16-byte aligned
    int2* sb = reinterpret_cast<int2*>(scs + P.Nb);
    // colour offsets in shared memory too: the acquire loads of the grid barrier invalidate L1 (CCTL.IVALL), so a global
    // goff[g] inside the sweep loops would cost an L2 round trip per colour and iteration
    int* sgoff = reinterpret_cast<int*>(sb + P.Nb);
    const int tau = blockIdx.x;
    const int taum = (tau == 0) ? L - 1 : tau - 1;
    const int taup = (tau == L - 1) ? 0 : tau + 1;
    for (int b = tid; b < P.Nb; b += T) {
        sb[b] = P.bonds[b];
        scs[b] = P.cs[b];
    }
    for (int g = tid; g <= P.ngroups; g += T) sgoff[g] = P.goff[g];
    double x[EPT], r[EPT], pprev[EPT], pc[EPT], Dc[EPT], Dn[EPT], pn[EPT];
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int i = tid + k * T;
        const bool in = i < N;
        x[k] = in ? P.x[(size_t)tau * N + i] : 0.0;
        r[k] = in ? P.R[(size_t)tau * N + i] : 0.0;
        Dc[k] = in ? P.D[(size_t)tau * N + i] : 0.0;
        Dn[k] = in ? P.D[(size_t)taup * N + i] : 0.0;
        pprev[k] = 0.0;
        pc[k] = 0.0;
    }
    const double normb = P.S->normb, eps0 = P.S->eps0, tol = P.S->tol, kappa_max = P.S->kappa_max;
    const long long maxiter = P.S->maxiter;
    double rdotr = P.S->rdotz, beta = 0.0, kmin = 0.0, eps = eps0;
    long long j = 0;
    double* Pold = P.P1;   // zeros on entry
    double* Pnew = P.P0;
    const bool wrap_c = (tau == 0), wrap_n = (taup == 0);
    __syncthreads();

    while (j < maxiter) {
        ++j;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int i = tid + k * T;
            if (i < N) {
                // both neighbour slices are fetched together: one L2 round trip per iteration
                const double pm = fma(beta, __ldcg(Pold + (size_t)taum * N + i), __ldcg(P.R + (size_t)taum * N + i));
                pn[k] = fma(beta, __ldcg(Pold + (size_t)taup * N + i), __ldcg(P.R + (size_t)taup * N + i));
                const double pcv = fma(beta, pprev[k], r[k]);
                pc[k] = pcv;
                Pnew[(size_t)tau * N + i] = pcv;
                A1[i] = Dc[k] * pm;
                A2[i] = Dn[k] * pcv;
            }
        }
        __syncthreads();
        for (int g = 0; g < P.ngroups; ++g) {   // K on both slices
            const int hi = sgoff[g + 1];
            for (int b = sgoff[g] + tid; b < hi; b += T) {
                const int2 ij = sb[b];
                const double2 c = scs[b];
                const double a1 = A1[ij.x], a2 = A1[ij.y], b1 = A2[ij.x], b2 = A2[ij.y];
                A1[ij.x] = c.x * a1 + c.y * a2;
                A1[ij.y] = c.x * a2 + c.y * a1;
                A2[ij.x] = c.x * b1 + c.y * b2;
                A2[ij.y] = c.x * b2 + c.y * b1;
            }
            __syncthreads();
        }
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int i = tid + k * T;
            if (i < N) {
                const double wc = wrap_c ? (pc[k] + A1[i]) : (pc[k] - A1[i]);
                const double wn = wrap_n ? (pn[k] + A2[i]) : (pn[k] - A2[i]);
                A1[i] = wc;
                A2[i] = wn;
                acc = fma(wc, wc, acc);
            }
        }
        __syncthreads();
        for (int g = P.ngroups - 1; g >= 0; --g) {   // K^T on w(tau+1)
            const int hi = sgoff[g + 1];
            for (int b = sgoff[g] + tid; b < hi; b += T) {
                const int2 ij = sb[b];
                const double2 c = scs[b];
                const double b1 = A2[ij.x], b2 = A2[ij.y];
                A2[ij.x] = c.x * b1 + c.y * b2;
                A2[ij.y] = c.x * b2 + c.y * b1;
            }
            __syncthreads();
        }
        const double blockA = block_sum_all(acc, red);
        const double pAp = grid_sum(blockA, P.partialA, P.bar, 2u * (unsigned int)j - 1u, nb, bcast);
        const double alpha = rdotr / pAp;
        double accr = 0.0;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int i = tid + k * T;
            if (i < N) {
                const double du = Dn[k] * A2[i];
                const double z = wrap_n ? (A1[i] + du) : (A1[i] - du);
                x[k] = fma(alpha, pc[k], x[k]);
                const double rv = fma(-alpha, z, r[k]);
                r[k] = rv;
                P.R[(size_t)tau * N + i] = rv;
                accr = fma(rv, rv, accr);
                pprev[k] = pc[k];
            }
        }
        const double blockB = block_sum_all(accr, red);
        const double rrn = grid_sum(blockB, P.partialB, P.bar, 2u * (unsigned int)j, nb, bcast);
        eps = sqrt(rrn) / normb;
        const double lg = log(2.0 * eps0 / eps);
        const double qq = 2.0 * (double)j / lg;
        const double kap = qq * qq;
        if (kap > kmin) kmin = kap;
        if (eps < tol || kmin > kappa_max) break;
        beta = rrn / rdotr;
        rdotr = rrn;
        double* tmp = Pold;
        Pold = Pnew;
        Pnew = tmp;
    }
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int i = tid + k * T;
        if (i < N) P.x[(size_t)tau * N + i] = x[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        P.S->iter = j;
        P.S->eps = eps;
        P.S->kappa_min = kmin;
        P.S->done = 1;
    }
}

template <typename Params>
void fill_io(Params& P, const CgBatchBufs& B, int L) {
    P.x = B.x; P.R = B.R; P.P0 = B.P0; P.P1 = B.P1;
    P.partialA = B.partial; P.partialB = B.partial + L; P.bar = B.bar; P.S = B.S;
    P.vstride = B.vstride; P.pstride = B.pstride;
}


// ---- any lattice, single-reduction form ------------------------------------------------------------------------------
// The loop of cg_p2p.cu (one grid barrier per iteration: gamma = (r, r) and delta = (r, A r) reduced together, alpha / beta
// from the Chronopoulos-Gear recurrences, the neighbour slices' new residuals rebuilt from their r, w, s of the previous
// iteration) with the body of the generic kernel above: slice pair + bond list in shared memory, x, r, p, s, w, D in
// registers.  grid.y = independent right-hand sides.  Measured at honeycomb L = 32 (config D, 20 CTAs per right-hand
// side): the barrier is cheap at this grid size and the shared-memory sweeps dominate, so the gain is small (HMC trajectory
// 16.1 -> 15.5 ms, ten right-hand sides 2.65 -> 2.56 ms); iteration counts unchanged.
struct G1Params {
    const double* __restrict__ D;     // expnV [L][N]
    double* __restrict__ x;           // in: initial guess, out: solution
    const double* __restrict__ R0;    // initial residual
    double* V;                        // [2 parities][3: r, w, s][L][N] per right-hand side
    double* partial;                  // [2 parities][2 values][L] per right-hand side
    unsigned int* bar;
    CgScalars* S;
    const int2* __restrict__ bonds;
    const double2* __restrict__ cs;
    const int* __restrict__ goff;
    int N, L, Nb, ngroups;
    long long vstride, v6stride;      // right-hand side k: x, R0 + k*vstride; V + k*v6stride; partial + k*4L; bar + k; S + k
};

// grid barrier + ordered sums of two doubles per CTA (see grid_sum); part: [2][L] of this barrier's parity
__device__ __forceinline__ void grid_sum2(double v0, double v1, double* part, unsigned int* bar, unsigned int seq, int nb, int L,
                                          double* bcast, double& out0, double& out1) {
    if (threadIdx.x == 0) {
        part[blockIdx.x] = v0;
        part[L + blockIdx.x] = v1;
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
        const unsigned int target = seq * (unsigned int)nb;
        while (ld_acquire(bar) < target) {}
    }
    __syncthreads();
    double s0 = 0.0, s1 = 0.0;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) { s0 += __ldcg(part + k); s1 += __ldcg(part + L + k); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if ((threadIdx.x & 31) == 0) { bcast[threadIdx.x >> 5] = s0; bcast[32 + (threadIdx.x >> 5)] = s1; }
    __syncthreads();
    double t0 = 0.0, t1 = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { t0 += bcast[k]; t1 += bcast[32 + k]; }
    out0 = t0;
    out1 = t1;
    __syncthreads();   // bcast is reused by the next call
}

__device__ __forceinline__ void block_sum_all2(double& v0, double& v1, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, o);
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = v0; red[32 + (threadIdx.x >> 5)] = v1; }
    __syncthreads();
    double t0 = 0.0, t1 = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { t0 += red[k]; t1 += red[32 + k]; }
    v0 = t0;
    v1 = t1;
}

template <int EPT, int MAXT>
__global__ void __launch_bounds__(MAXT) cg1r_generic_kernel(G1Params P) {
    extern __shared__ __align__(16) unsigned char gsm[];
    __shared__ double red[64];
    __shared__ double bcast[64];
    {
        const size_t k = blockIdx.y;
        P.x += k * P.vstride; P.R0 += k * P.vstride; P.V += k * P.v6stride;
        P.partial += k * 4 * (size_t)P.L; P.bar += k; P.S += k;
    }
    const int N = P.N, L = P.L, nb = gridDim.x, T = blockDim.x, tid = threadIdx.x;
    double* A1 = reinterpret_cast<double*>(gsm);
    double* A2 = A1 + N;
    double2* scs = reinterpret_cast<double2*>(A2 + N);
    int2* sb = reinterpret_cast<int2*>(scs + P.Nb);
    // colour offsets in shared memory too: the acquire loads of the grid barrier invalidate L1 (CCTL.IVALL), so a global
    // goff[g] inside the sweep loops would cost an L2 round trip per colour and iteration
    int* sgoff = reinterpret_cast<int*>(sb + P.Nb);
    const int tau = blockIdx.x;
    const int taum = (tau == 0) ? L - 1 : tau - 1;
    const int taup = (tau == L - 1) ? 0 : tau + 1;
    for (int b = tid; b < P.Nb; b += T) {
        sb[b] = P.bonds[b];
        scs[b] = P.cs[b];
    }
    for (int g = tid; g <= P.ngroups; g += T) sgoff[g] = P.goff[g];
    const bool wrap_c = (tau == 0), wrap_n = (taup == 0);
    const size_t vs = (size_t)L * N;
    auto vec = [&](int par, int which) -> double* { return P.V + (size_t)(par * 3 + which) * vs; };
    double x[EPT], r[EPT], p[EPT], s[EPT], w[EPT], Dc[EPT], Dn[EPT], vm[EPT], vp[EPT];

    // w(tau) = (M^T M v)(tau) from vm = v(tau-1), r = v(tau), vp = v(tau+1); returns this thread's share of |(M v)(tau)|^2
    auto apply_A = [&]() -> double {
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int i = tid + k * T;
            if (i < N) {
                A1[i] = Dc[k] * vm[k];
                A2[i] = Dn[k] * r[k];
            }
        }
        __syncthreads();
        for (int g = 0; g < P.ngroups; ++g) {   // K on both slices
            const int hi = sgoff[g + 1];
            for (int b = sgoff[g] + tid; b < hi; b += T) {
                const int2 ij = sb[b];
                const double2 c = scs[b];
                const double a1 = A1[ij.x], a2 = A1[ij.y], b1 = A2[ij.x], b2 = A2[ij.y];
                A1[ij.x] = c.x * a1 + c.y * a2;
                A1[ij.y] = c.x * a2 + c.y * a1;
                A2[ij.x] = c.x * b1 + c.y * b2;
                A2[ij.y] = c.x * b2 + c.y * b1;
            }
            __syncthreads();
        }
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int i = tid + k * T;
            if (i < N) {
                const double wc = wrap_c ? (r[k] + A1[i]) : (r[k] - A1[i]);
                const double wn = wrap_n ? (vp[k] + A2[i]) : (vp[k] - A2[i]);
                A1[i] = wc;
                A2[i] = wn;
                acc = fma(wc, wc, acc);
            }
        }
        __syncthreads();
        for (int g = P.ngroups - 1; g >= 0; --g) {   // K^T on (M v)(tau+1)
            const int hi = sgoff[g + 1];
            for (int b = sgoff[g] + tid; b < hi; b += T) {
                const int2 ij = sb[b];
                const double2 c = scs[b];
                const double b1 = A2[ij.x], b2 = A2[ij.y];
                A2[ij.x] = c.x * b1 + c.y * b2;
                A2[ij.y] = c.x * b2 + c.y * b1;
            }
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int i = tid + k * T;
            if (i < N) {
                const double du = Dn[k] * A2[i];
                w[k] = wrap_n ? (A1[i] + du) : (A1[i] - du);
            }
        }
        __syncthreads();   // A1 / A2 are rewritten by the next call
        return acc;
    };

    // ---- set-up: r_0 (from the caller), w_0 = A r_0, gamma_0, delta_0 ------------------------------------------------
    double accg = 0.0;
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int i = tid + k * T;
        const bool in = i < N;
        x[k] = in ? P.x[(size_t)tau * N + i] : 0.0;
        r[k] = in ? P.R0[(size_t)tau * N + i] : 0.0;
        vm[k] = in ? P.R0[(size_t)taum * N + i] : 0.0;
        vp[k] = in ? P.R0[(size_t)taup * N + i] : 0.0;
        Dc[k] = in ? P.D[(size_t)tau * N + i] : 0.0;
        Dn[k] = in ? P.D[(size_t)taup * N + i] : 0.0;
        p[k] = 0.0;
        s[k] = 0.0;
        w[k] = 0.0;
        accg = fma(r[k], r[k], accg);
    }
    __syncthreads();   // bond list staged
    double accd = apply_A();
    {
        double* R0w = vec(0, 0);
        double* W0w = vec(0, 1);
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int i = tid + k * T;
            if (i < N) {
                R0w[(size_t)tau * N + i] = r[k];
                W0w[(size_t)tau * N + i] = w[k];
            }
        }
    }
    const double normb = P.S->normb, tol = P.S->tol, kappa_max = P.S->kappa_max;
    const long long maxiter = P.S->maxiter;
    unsigned int seq = 0;
    double gamma = accg, delta = accd;
    block_sum_all2(gamma, delta, red);
    ++seq;
    grid_sum2(gamma, delta, P.partial + (size_t)(seq & 1u) * 2 * L, P.bar, seq, nb, L, bcast, gamma, delta);
    const double eps0 = sqrt(gamma) / normb;
    double alpha = gamma / delta, beta = 0.0, kmin = 0.0, eps = eps0;
    long long j = 0;

    while (j < maxiter) {
        const int rd = (int)(j & 1), wr = rd ^ 1;
        ++j;
        const bool have_s = (j > 1);
        const double* Rr = vec(rd, 0);
        const double* Wr = vec(rd, 1);
        const double* Sr = vec(rd, 2);
        double* Rw = vec(wr, 0);
        double* Ww = vec(wr, 1);
        double* Sw = vec(wr, 2);
        const double mab = -alpha * beta;
        double accr = 0.0;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int i = tid + k * T;
            if (i < N) {
                // r_new of the neighbour slices = r - alpha (w + beta s), rebuilt here; both fetched together
                const size_t em = (size_t)taum * N + i, ep = (size_t)taup * N + i;
                double a = fma(-alpha, __ldcg(Wr + em), __ldcg(Rr + em));
                double b = fma(-alpha, __ldcg(Wr + ep), __ldcg(Rr + ep));
                if (have_s) {
                    a = fma(mab, __ldcg(Sr + em), a);
                    b = fma(mab, __ldcg(Sr + ep), b);
                }
                vm[k] = a;
                vp[k] = b;
                const double pv = fma(beta, p[k], r[k]);
                const double sv = fma(beta, s[k], w[k]);
                p[k] = pv;
                s[k] = sv;
                x[k] = fma(alpha, pv, x[k]);
                const double rv = fma(-alpha, sv, r[k]);
                r[k] = rv;
                Rw[(size_t)tau * N + i] = rv;
                Sw[(size_t)tau * N + i] = sv;
                accr = fma(rv, rv, accr);
            }
        }
        double accw = apply_A();
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int i = tid + k * T;
            if (i < N) Ww[(size_t)tau * N + i] = w[k];
        }
        block_sum_all2(accr, accw, red);
        ++seq;
        double gnew, dnew;
        grid_sum2(accr, accw, P.partial + (size_t)(seq & 1u) * 2 * L, P.bar, seq, nb, L, bcast, gnew, dnew);
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
    for (int k = 0; k < EPT; ++k) {
        const int i = tid + k * T;
        if (i < N) P.x[(size_t)tau * N + i] = x[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        P.S->iter = j;
        P.S->eps = eps;
        P.S->kappa_min = kmin;
        P.S->done = 1;
    }
}

template <typename Kern>
bool launch_g1(elph_handle* h, Kern kern, const G1Params& P, int threads, size_t smem, int nrhs) {
    elph_enable_smem(h, kern);
    int per_sm = 0;
    ELPH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    const int cap = (int)std::min<long long>((long long)per_sm * h->sm_count / h->L, 65535);
    if (cap < 1) return false;
    for (int k0 = 0; k0 < nrhs; k0 += cap) {
        const int g = std::min(cap, nrhs - k0);
        G1Params Q = P;
        Q.x += k0 * P.vstride; Q.R0 += k0 * P.vstride; Q.V += k0 * P.v6stride;
        Q.partial += (size_t)k0 * 4 * h->L; Q.bar += k0; Q.S += k0;
        ELPH_CUDA(cudaMemsetAsync(Q.bar, 0, g * sizeof(unsigned int), h->stream));
        void* args[] = {&Q};
        ELPH_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(h->L, g), dim3(threads), args, smem, h->stream));
        h->launches++;
    }
    return true;
}

// work space of the single-reduction generic kernel: 6 vectors + 4 L partials per right-hand side, grow-only
bool cg1r_generic(elph_handle* h, int nrhs, const CgBatchBufs& B) {
    if (h->model != ELPH_MODEL_HOLSTEIN || h->N > 4096 || h->L < 2) return false;
    const size_t smem = (size_t)(2 * h->N + 2) * sizeof(double) + (size_t)h->Nb * (sizeof(double2) + sizeof(int2)) +
                        (size_t)(h->ngroups + 1) * sizeof(int);
    if (smem > h->smem_optin) return false;
    auto& W = h->g1r;
    if (W.cap < nrhs) {
        if (W.V) cudaFree(W.V);
        if (W.partial) cudaFree(W.partial);
        W.V = elph_dalloc<double>((size_t)6 * h->Ndim * nrhs);
        W.partial = elph_dalloc<double>((size_t)4 * h->L * nrhs);
        W.cap = nrhs;
    }
    G1Params P;
    P.D = h->d_D; P.x = B.x; P.R0 = B.R; P.V = W.V; P.partial = W.partial; P.bar = B.bar; P.S = B.S;
    P.bonds = h->d_bonds; P.cs = h->d_cs; P.goff = h->d_goff;
    P.N = h->N; P.L = h->L; P.Nb = h->Nb; P.ngroups = h->ngroups;
    P.vstride = B.vstride; P.v6stride = (long long)6 * h->Ndim;
    const int threads = (h->N <= 64) ? 64 : ((h->N <= 256) ? 256 : 512);
    const int ept = (h->N + threads - 1) / threads;
    if (ept <= 1) return launch_g1(h, cg1r_generic_kernel<1, 512>, P, threads, smem, nrhs);
    if (ept <= 2) return launch_g1(h, cg1r_generic_kernel<2, 512>, P, threads, smem, nrhs);
    if (ept <= 4) return launch_g1(h, cg1r_generic_kernel<4, 512>, P, threads, smem, nrhs);
    return launch_g1(h, cg1r_generic_kernel<8, 512>, P, threads, smem, nrhs);
}

bool cg_persistent_generic(elph_handle* h, int nrhs, const CgBatchBufs& B) {
    if (h->model != ELPH_MODEL_HOLSTEIN || h->N > 4096) return false;
    if (h->cg_single_reduction != 0 && cg1r_generic(h, nrhs, B)) return true;
    const size_t smem = (size_t)(2 * h->N + 2) * sizeof(double) + (size_t)h->Nb * (sizeof(double2) + sizeof(int2)) +
                        (size_t)(h->ngroups + 1) * sizeof(int);
    if (smem > h->smem_optin) return false;
    GcgParams P;
    fill_io(P, B, h->L);
    P.D = h->d_D;
    P.bonds = h->d_bonds; P.cs = h->d_cs; P.goff = h->d_goff;
    P.N = h->N; P.L = h->L; P.Nb = h->Nb; P.ngroups = h->ngroups;
    // measured (honeycomb L=32, N=2048): 512 threads x 4 elements beat 1024 x 2 (cheaper barriers, no spills)
    const int threads = (h->N <= 64) ? 64 : ((h->N <= 256) ? 256 : 512);
    const int ept = (h->N + threads - 1) / threads;
    if (ept <= 1) return launch_groups(h, cg_persistent_generic_kernel<1, 512>, P, threads, smem, nrhs);
    if (ept <= 2) return launch_groups(h, cg_persistent_generic_kernel<2, 512>, P, threads, smem, nrhs);
    if (ept <= 4) return launch_groups(h, cg_persistent_generic_kernel<4, 512>, P, threads, smem, nrhs);
    return launch_groups(h, cg_persistent_generic_kernel<8, 512>, P, threads, smem, nrhs);
}

}  // namespace

// nrhs independent unpreconditioned solves on the same field.  For right-hand side k the caller has set up r0 in
// B.R + k*vstride, the initial guess in B.x + k*vstride, zeros in B.P1 + k*vstride and the scalar block B.S[k]
// (normb, eps0, rdotz = r0.r0, tol, ...) with cg_init_kernel.  B.partial: 2L doubles per right-hand side (pstride),
// B.bar: one counter per right-hand side.  Returns false if the persistent kernels do not apply to this handle.
bool elph_cg_persistent_batch(elph_handle* h, int nrhs, const CgBatchBufs& B) {
    const bool ssh = (h->model == ELPH_MODEL_SSH);
    if (!h->use_persistent || h->sharded || nrhs < 1) return false;
    int dev_coop = 0;
    cudaDeviceGetAttribute(&dev_coop, cudaDevAttrCooperativeLaunch, h->device);
    if (!dev_coop) return false;
    if (!ssh && h->hc.enabled && !h->sq_disable && h->hc_tiles && h->L >= 4) {
        // honeycomb lattice 32 cells wide (config D): register tiles, 4 rows of cells per warp
        const int nwarps = h->hc.L2 / 4;
        if (nwarps >= 2 && nwarps <= 16) {
            PcgParams P;
            fill_io(P, B, h->L);
            P.D = h->d_D; P.tab = nullptr;
            P.L = h->L; P.Ly = h->hc.L2;
            P.c0 = h->hc.c[0]; P.s0 = h->hc.s[0]; P.c1 = h->hc.c[1]; P.s1 = h->hc.s[1];
            P.c2 = h->hc.c[2]; P.s2 = h->hc.s[2]; P.c3 = 1.0; P.s3 = 0.0;
            if (nwarps * 32 <= 256 && launch_persistent<2, 4, 256, 2>(h, P, nwarps, nrhs)) return true;
            if (nwarps * 32 > 256 && launch_persistent<2, 4, 512, 2>(h, P, nwarps, nrhs)) return true;
        }
    }
    if (!(ssh ? h->ssq.enabled : h->sq.enabled) || h->sq_disable || h->L < 4) return cg_persistent_generic(h, nrhs, B);
    const int Lx = ssh ? h->ssq.Lx : h->sq.Lx, Ly = ssh ? h->ssq.Ly : h->sq.Ly;
    const int PY = (Lx == 32) ? 8 : 4;
    if (Ly % PY) return false;
    const int nwarps = Ly / PY;
    if (nwarps < 2 || nwarps > 32) return false;
    PcgParams P;
    fill_io(P, B, h->L);
    P.D = h->d_D; P.tab = ssh ? h->ssq.d_tab : nullptr;
    P.L = h->L; P.Ly = Ly;
    P.c0 = h->sq.c[0]; P.s0 = h->sq.s[0]; P.c1 = h->sq.c[1]; P.s1 = h->sq.s[1];
    P.c2 = h->sq.c[2]; P.s2 = h->sq.s[2]; P.c3 = h->sq.c[3]; P.s3 = h->sq.s[3];
    if (ssh) {
        if (Lx == 32 && PY == 8 && nwarps * 32 <= 128) return launch_persistent<1, 8, 128, 1, 3>(h, P, nwarps, nrhs);
        if (Lx == 32 && PY == 8 && nwarps * 32 <= 256) return launch_persistent<1, 8, 256, 1>(h, P, nwarps, nrhs);
        return false;
    }
    if (Lx == 32 && PY == 8 && nwarps * 32 <= 128) return launch_persistent<1, 8, 128, 0, 3>(h, P, nwarps, nrhs);
    if (Lx == 32 && PY == 8 && nwarps * 32 <= 256) return launch_persistent<1, 8, 256, 0>(h, P, nwarps, nrhs);
    if (Lx == 64 && PY == 4 && nwarps * 32 <= 512) return launch_persistent<2, 4, 512, 0>(h, P, nwarps, nrhs);
    return false;
}

// single solve on the handle's own CG buffers (elph_cg_device)
bool elph_cg_persistent(elph_handle* h, double* x_dev) {
    if (h->partial_cap < 2 * h->L) return false;
    CgBatchBufs B;
    B.x = x_dev; B.R = h->d_r; B.P0 = h->d_p[0]; B.P1 = h->d_p[1];
    B.partial = h->d_partial; B.bar = h->d_bar; B.S = h->d_cg;
    B.vstride = 0; B.pstride = 0;
    return elph_cg_persistent_batch(h, 1, B);
}
