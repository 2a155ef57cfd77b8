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

namespace {

using namespace sqt;

struct PcgParams {
    const double* __restrict__ D;    // Holstein: expnV [L][N]; SSH: exp(dtau mu) [N]
    const double2* __restrict__ tab; // SSH: (cosh, sinh) [L][2][N] in the tile layout of ssh_square.cu
    double* __restrict__ x;          // [L][N] in: initial guess, out: solution
    double* R;                       // [L][N] residual (in: r0), updated every iteration
    double* P0;                      // [L][N] p buffers (double-buffered); P1 must hold zeros on entry
    double* P1;
    double* partialA;                // [L]
    double* partialB;                // [L]
    unsigned int* bar;               // monotonically increasing arrival counter (zero on entry)
    CgScalars* S;                    // in: normb, eps0, rdotz (= r0.r0), tol, kappa_max, maxiter ; out: iter, eps, done
    int L, Ly;
    double c0, s0, c1, s1, c2, s2, c3, s3;
};

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// grid barrier fused with a sum over all CTAs; every CTA returns the same bits (fixed order: lane-strided, then tree)
__device__ __forceinline__ double grid_sum(double block_value, double* partial, unsigned int* bar, unsigned int target, int nb,
                                           double* bcast) {
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = block_value;
        __threadfence();
        atomicAdd(bar, 1u);
        while (ld_acquire(bar) < target) {}
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        double s = 0.0;
        for (int k = threadIdx.x; k < nb; k += 32) s += __ldcg(partial + k);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) *bcast = s;
    }
    __syncthreads();
    return *bcast;
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

template <int NSEG, int PY, int MAXT, bool SSH>
__global__ void __launch_bounds__(MAXT) cg_persistent_kernel(PcgParams P) {
    constexpr int LX = 32 * NSEG;
    extern __shared__ __align__(16) double strips[];   // 2 x [nwarps][4][LX]; SSH: + the tables of slices tau and tau+1
    __shared__ double red[32];
    __shared__ double bcast;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int L = P.L, N = LX * P.Ly, nb = gridDim.x;
    const int tau = blockIdx.x;
    const int taum = (tau == 0) ? L - 1 : tau - 1;
    const int taup = (tau == L - 1) ? 0 : tau + 1;
    const size_t tile_off = (size_t)warp * PY * LX;
    auto eidx = [&](int r, int q) -> size_t { return tile_off + r * LX + 32 * q + lane; };

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
    unsigned int target = 0;
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
        // ---- alpha ------------------------------------------------------------------------------------------------
        const double blockA = tile_block_sum<NSEG, PY>(acc, red, lane, warp, nwarps);
        target += nb;
        const double pAp = grid_sum(blockA, P.partialA, P.bar, target, nb, &bcast);
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
        target += nb;
        const double rrn = grid_sum(blockB, P.partialB, P.bar, target, nb, &bcast);
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

template <int NSEG, int PY, int MAXT, bool SSH>
bool launch_persistent(elph_handle* h, PcgParams& P, int nwarps) {
    constexpr int LX = 32 * NSEG;
    const size_t smem = 2ull * nwarps * 4 * LX * sizeof(double) + (SSH ? 2ull * 2 * h->N * sizeof(double2) : 0);
    if (smem > h->smem_optin) return false;
    auto kern = cg_persistent_kernel<NSEG, PY, MAXT, SSH>;
    elph_enable_smem(h, kern);
    int per_sm = 0;
    ELPH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nwarps * 32, smem));
    if ((long long)per_sm * h->sm_count < h->L) return false;   // all time slices must be co-resident
    ELPH_CUDA(cudaMemsetAsync(h->d_bar, 0, sizeof(unsigned int), h->stream));
    void* args[] = {&P};
    ELPH_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(h->L), dim3(nwarps * 32), args, smem, h->stream));
    h->launches++;
    return true;
}

}  // namespace

// r0 (in h->d_r), the scalar block (normb, eps0, rdotz = r0.r0, ...) and zeros in h->d_p[1] must be set up by the caller
// (elph_cg_device does that with the same kernels as the multi-launch path).  Returns false if not applicable.
bool elph_cg_persistent(elph_handle* h, double* x_dev) {
    const bool ssh = (h->model == ELPH_MODEL_SSH);
    if (!(ssh ? h->ssq.enabled : h->sq.enabled) || h->sq_disable || !h->use_persistent || h->sharded) return false;
    if (h->L < 4) return false;
    int dev_coop = 0;
    cudaDeviceGetAttribute(&dev_coop, cudaDevAttrCooperativeLaunch, h->device);
    if (!dev_coop) return false;
    const int Lx = ssh ? h->ssq.Lx : h->sq.Lx, Ly = ssh ? h->ssq.Ly : h->sq.Ly;
    const int PY = (Lx == 32) ? 8 : 4;
    if (Ly % PY) return false;
    const int nwarps = Ly / PY;
    if (nwarps < 2 || nwarps > 32) return false;
    if (h->partial_cap < 2 * h->L) return false;
    PcgParams P;
    P.D = h->d_D; P.tab = ssh ? h->ssq.d_tab : nullptr; P.x = x_dev; P.R = h->d_r; P.P0 = h->d_p[0]; P.P1 = h->d_p[1];
    P.partialA = h->d_partial; P.partialB = h->d_partial + h->L; P.bar = h->d_bar; P.S = h->d_cg;
    P.L = h->L; P.Ly = Ly;
    P.c0 = h->sq.c[0]; P.s0 = h->sq.s[0]; P.c1 = h->sq.c[1]; P.s1 = h->sq.s[1];
    P.c2 = h->sq.c[2]; P.s2 = h->sq.s[2]; P.c3 = h->sq.c[3]; P.s3 = h->sq.s[3];
    if (ssh) {
        if (Lx == 32 && PY == 8 && nwarps * 32 <= 256) return launch_persistent<1, 8, 256, true>(h, P, nwarps);
        return false;
    }
    if (Lx == 32 && PY == 8 && nwarps * 32 <= 256) return launch_persistent<1, 8, 256, false>(h, P, nwarps);
    if (Lx == 64 && PY == 4 && nwarps * 32 <= 512) return launch_persistent<2, 4, 512, false>(h, P, nwarps);
    return false;
}
