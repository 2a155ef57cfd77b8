// KPM-preconditioned conjugate gradient on A = M^T M as ONE persistent cooperative kernel (Holstein model, periodic square
// lattices 32 sites wide: config B).
//
// Replaces the loop of solve!(x, A, b, cg, P) (src/IterativeSolvers.jl:153-234) with ldiv!(z, P, r) of the symmetric KPM
// preconditioner (src/KPMPreconditioners.jl:426-481, 606-679) inlined: the launch-per-phase form (cg.cu) spends 4 kernels
// per iteration -- [p update + M^T M + p.Ap] [x/r update + stop rule + tau-FFT] [Chebyshev chains] [inverse FFT + r.z] --
// and ~70 us per iteration at 32x32xL200, of which the chain of the lowest Matsubara frequency (2 x 68 dependent sweeps on
// two SMs) is ~26 us and the rest launch gaps, pipeline fill and drain.  Here the four phases are stages of one kernel,
// separated by grid barriers that also carry the scalar reductions (every CTA folds the same partial sums in the same
// order, so alpha, beta and the stop rule are evaluated identically everywhere and nothing returns to the host):
//
//   F  column-parallel: x += alpha p, r -= alpha Ap, |r|^2; nu = FFT_tau(theta .* r), only the frequencies w < ceil(L/2) that
//      the chains read are written                                                         -> barrier, stop rule
//   C  one 2-CTA cluster per frequency, longest polynomial first (re / im chains, tanh-form sweeps, kpm_chain.cuh); only
//      nu'(w) is written, the mirror frequency L-1-w = conj is rebuilt by the reader       -> barrier
//   I  column-parallel: z = Re(conj(theta) .* iFFT(nu')) / L, r.z                          -> barrier, beta
//   A  slice-parallel (chunks of consecutive time slices, register tiles as in mtm_square.cu): p = z + beta p, Ap = M^T M p,
//      p.Ap = |M p|^2                                                                      -> barrier, alpha
//
// All vectors stay in L2 (1.6 MB each at config B).  Same arithmetic as the launch-per-phase path up to the order of the
// partial sums; iteration counts are required to agree within +-2 with the reference recurrence (tests/test_gpu_pcg_fused.py).
#include "fft_smem.cuh"
#include "kpm_chain.cuh"
#include "square_tiles.cuh"

#include <algorithm>

namespace {

using namespace sqt;
using namespace fftsm;
using namespace kpmch;

constexpr unsigned int kSpinLimit = 1u << 26;   // a grid barrier that is not completed after ~10 s ends the solve with an error

struct FusedParams {
    double* x;
    double* r;
    double* p[2];
    double* ap;
    double* z;
    cplx* nu_in;
    cplx* nu_out;
    const double* D;        // Holstein: expnV [L][N]; SSH: exp(dtau mu) [N]
    const double2* ssh_tab; // SSH: (cosh, sinh)(dtau t') per time slice in the tile layout [L][2][N] (ssh_square.cu)
    FftPlan plan;
    FftPlan plan_half;      // L even: the transforms run at length L/2 (see phase F)
    const cplx* tw;
    const cplx* theta;
    KsqParams K;            // eVbar, coefficients, orders, schedule, window, tanh-form constants
    int Lo2, max_order;
    double* partial;        // [2][gridDim.x]
    unsigned int* bar;
    CgScalars* S;
    int L, Ly, Cs, nchunks;
    unsigned long long* prof;
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Grid barrier fused with a sum over all CTAs (the scheme of cg_persistent.cu: one monotone arrival counter, partial sums in
// a buffer that alternates between consecutive barriers).  Returns the same bits on every thread of every CTA; `failed` is
// raised when the barrier timed out.
__device__ __forceinline__ double grid_sum(double thread_value, double* partial, unsigned int* bar, unsigned int seq, double* red,
                                           double* bcast, int* failed) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, nb = gridDim.x;
    double v = thread_value;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();   // also orders the CTA's global writes of the phase before the arrival below
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double* slot = partial + (size_t)(seq & 1u) * nb;
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < nwarps; ++k) t += red[k];
        slot[blockIdx.x] = t;
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
        const unsigned int target = seq * (unsigned int)nb;
        unsigned int spins = 0;
        while (ld_acquire_u32(bar) < target) {
            if (++spins > kSpinLimit) { *failed = 1; break; }
        }
    }
    __syncthreads();
    double s = 0.0;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) s += __ldcg(slot + k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) bcast[warp] = s;
    __syncthreads();
    double t = 0.0;
    for (int k = 0; k < nwarps; ++k) t += bcast[k];
    return t;
}

// one copy of the transform code for the four call sites (straight-line code that runs once per iteration is paid for in
// instruction fetches, not in arithmetic)
template <int SB>
__device__ __noinline__ cplx* fft_smem_shared(cplx* x, cplx* y, const FftPlan* plan, const cplx* tw, bool inverse) {
    return fft_smem_auto<SB>(x, y, *plan, tw, inverse);
}

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// (registers capped below the full register file of an SM -- 512 threads x 128 -- so that the cooperative launch also fits when a
//  profiler or debugger reserves resources on the SM)
// SSH: per-bond hoppings.  The chains read the tau-averaged tables of the preconditioner and the product phase the tables of the
// CTA's own time slices (chunk + halo slice); both sets stay in shared memory for the whole solve (the field is fixed during a
// solve and the chunk of a CTA never changes).
template <int NSEG, int PY, int SB, int MAXT, bool SSH = false>
__global__ void __maxnreg__(MAXT == 512 ? 120 : 56) pcg_fused_kernel(FusedParams P) {
    constexpr int LX = 32 * NSEG;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[32];
    __shared__ double bcast[32];
    __shared__ int failed;
    __shared__ FftPlan plans[2];   // [0] full length, [1] half length
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int L = P.L, N = LX * P.Ly, G = gridDim.x, cta = blockIdx.x;
    unsigned int crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));

    // shared memory: [twiddles L][theta L] then a region shared by the phases
    cplx* tw = reinterpret_cast<cplx*>(smem_raw);
    cplx* th = tw + L;
    cplx* tw2 = th + L;                                                           // [L/2] twiddles of the half-length transform
    unsigned char* region = reinterpret_cast<unsigned char*>(tw2 + (L + 1) / 2);
    // FFT phases
    cplx* b0 = reinterpret_cast<cplx*>(region);
    cplx* b1 = b0 + (size_t)L * SB;
    // chain phase
    cplx* c_s = reinterpret_cast<cplx*>(region);                                  // [max_order]
    double* strips = reinterpret_cast<double*>(c_s + P.max_order);                // 2 x [nwarps][2][LX]
    double* xch = strips + 2ull * nwarps * 2 * LX;                                // 2 x [N], written by the partner CTA
    // SSH tables behind the region shared by the phases
    const size_t region_bytes = max(2ull * L * SB * sizeof(cplx),
                                    (size_t)P.max_order * sizeof(cplx) + (2ull * nwarps * 2 * LX + 2ull * N) * sizeof(double));
    double2* tabbar = reinterpret_cast<double2*>(region + ((region_bytes + 15) & ~size_t(15)));   // [2][N] tau-averaged
    double2* tabA = tabbar + 2 * N;                                                               // [Cs + 1][2][N]

    const bool half = (L % 2 == 0) && P.plan_half.L == L / 2;
    const int Lh = L / 2;
    for (int k = threadIdx.x; k < L; k += blockDim.x) {
        tw[k] = P.tw[k];
        th[k] = P.theta[k];
        if (half && k < Lh) tw2[k] = P.tw[2 * k];
    }
    if (threadIdx.x == 0) {
        failed = 0;
        plans[0] = P.plan;
        plans[1] = P.plan_half;
    }
    if constexpr (SSH) {
        for (int k = threadIdx.x; k < 2 * N; k += blockDim.x) tabbar[k] = P.K.tab[k];
        if (cta < P.nchunks) {
            const int a0 = cta * P.Cs, ns = min(P.Cs, L - a0) + 1;
            for (int j = 0; j < ns; ++j) {
                int tau = a0 + j;
                if (tau >= L) tau -= L;
                for (int k = threadIdx.x; k < 2 * N; k += blockDim.x) tabA[(size_t)j * 2 * N + k] = P.ssh_tab[(size_t)tau * 2 * N + k];
            }
        }
    }
    __syncthreads();
    cluster_sync_all();   // the partner CTA is running before its shared memory is addressed
    const uint32_t my_xch = (uint32_t)__cvta_generic_to_shared(xch);
    uint32_t remote_xch;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote_xch) : "r"(my_xch), "r"(crank ^ 1u));

    const double normb = P.S->normb, eps0 = P.S->eps0, tol = P.S->tol, kappa_max = P.S->kappa_max;
    const long long maxiter = P.S->maxiter;
    double alpha = 0.0, beta = 0.0, rdotz = 0.0, kmin = 0.0, eps = eps0, pAp = 0.0;
    long long iter = 0;
    unsigned int seq = 0;
    int par = 0;
    bool first = true;
    unsigned int nswap = 0;
    const int site = threadIdx.x % SB, slot = threadIdx.x / SB, nslots = blockDim.x / SB;
    const int ngroups = (N + SB - 1) / SB;
    const double invL = 1.0 / (double)L;
    const size_t tile_off = (size_t)warp * PY * LX;
    unsigned long long tph[6] = {0, 0, 0, 0, 0, 0};
    unsigned long long tf[2] = {0, 0};
    long long tlast = clock64();
    auto lap = [&](int k) {
        const long long now = clock64();
        tph[k] += (unsigned long long)(now - tlast);
        tlast = now;
    };

    for (;;) {
        // ---- F: x += alpha p, r -= alpha Ap, |r|^2 (src/IterativeSolvers.jl:205-211); nu = FFT(theta .* r) ------------------
        double acc = 0.0;
        {
            const double* pk = P.p[par];
            for (int grp = cta; grp < ngroups; grp += G) {
                const int gsite = grp * SB + site;
                const bool ok = gsite < N;
                if (half) {
                    // L even: theta .* r is an odd-frequency transform of a real sequence,
                    //     nu(w) = sum_t r_t e^{-i pi t (2w+1)/L} = sum_{t<L/2} (r_t - i (-1)^w r_{t+L/2}) e^{-i pi t (2w+1)/L},
                    // so ONE complex transform of length L/2 of c_t = theta_t (r_t - i r_{t+L/2}) gives all of them:
                    //     nu(2k) = C_k ,  nu(2k+1) = conj(C_{L/2-1-k}).
                    // Two (t, t+L/2) pairs per thread and round; all loads of a round precede the first store (x, r, p, Ap may
                    // alias as far as the compiler knows), so a round costs one L2 round trip.
                    for (int t0 = slot; t0 < Lh; t0 += 2 * nslots) {
                        double xv[4], rv[4], pv[4], av[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int t = t0 + (k >> 1) * nslots + (k & 1) * Lh;
                            xv[k] = rv[k] = pv[k] = av[k] = 0.0;
                            if (ok && t0 + (k >> 1) * nslots < Lh) {
                                const size_t e = (size_t)t * N + gsite;
                                rv[k] = P.r[e];
                                if (!first) {
                                    xv[k] = P.x[e];
                                    pv[k] = __ldcg(pk + e);
                                    av[k] = __ldcg(P.ap + e);
                                }
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int t = t0 + (k >> 1) * nslots + (k & 1) * Lh;
                            if (ok && !first && t0 + (k >> 1) * nslots < Lh) {
                                const size_t e = (size_t)t * N + gsite;
                                P.x[e] = fma(alpha, pv[k], xv[k]);
                                const double xr = fma(-alpha, av[k], rv[k]);
                                P.r[e] = xr;
                                acc = fma(xr, xr, acc);
                                rv[k] = xr;
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const int t = t0 + k * nslots;
                            if (t < Lh) {
                                const cplx tt = th[t];
                                const double a = rv[2 * k], b = rv[2 * k + 1];
                                b0[(size_t)t * SB + site] = make_double2(fma(a, tt.x, b * tt.y), fma(a, tt.y, -b * tt.x));
                            }
                        }
                    }
                    __syncthreads();
                    const long long tq0 = clock64();
                    cplx* res = fft_smem_shared<SB>(b0, b1, &plans[1], tw2, false);
                    tf[0] += (unsigned long long)(clock64() - tq0);
                    for (int w = slot; w < P.Lo2; w += nslots)
                        if (ok) {
                            cplx v;
                            if (w & 1) {
                                v = res[(size_t)(Lh - 1 - (w >> 1)) * SB + site];
                                v.y = -v.y;
                            } else {
                                v = res[(size_t)(w >> 1) * SB + site];
                            }
                            P.nu_in[(size_t)w * N + gsite] = v;
                        }
                    __syncthreads();
                    continue;
                }
                // four time slices per thread and round (loads before stores, as above)
                for (int t0 = slot; t0 < L; t0 += 4 * nslots) {
                    double xv[4], rv[4], pv[4], av[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int t = t0 + k * nslots;
                        xv[k] = rv[k] = pv[k] = av[k] = 0.0;
                        if (ok && t < L) {
                            const size_t e = (size_t)t * N + gsite;
                            rv[k] = P.r[e];
                            if (!first) {
                                xv[k] = P.x[e];
                                pv[k] = __ldcg(pk + e);
                                av[k] = __ldcg(P.ap + e);
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int t = t0 + k * nslots;
                        if (t < L) {
                            double xr = rv[k];
                            if (ok && !first) {
                                const size_t e = (size_t)t * N + gsite;
                                P.x[e] = fma(alpha, pv[k], xv[k]);
                                xr = fma(-alpha, av[k], rv[k]);
                                P.r[e] = xr;
                                acc = fma(xr, xr, acc);
                            }
                            const cplx tt = th[t];
                            b0[(size_t)t * SB + site] = make_double2(tt.x * xr, tt.y * xr);
                        }
                    }
                }
                __syncthreads();
                const long long tq0 = clock64();
                cplx* res = fft_smem_shared<SB>(b0, b1, &plans[0], tw, false);
                tf[0] += (unsigned long long)(clock64() - tq0);
                for (int t = slot; t < P.Lo2; t += nslots)
                    if (ok) P.nu_in[(size_t)t * N + gsite] = res[(size_t)t * SB + site];
                __syncthreads();
            }
        }
        lap(0);
        const double rr = grid_sum(acc, P.partial, P.bar, ++seq, red, bcast, &failed);
        if (failed) break;
        if (!first) {
            // stop rule (src/IterativeSolvers.jl:208-219)
            const long long j = iter + 1;
            eps = sqrt(rr) / normb;
            const double lg = log(2.0 * eps0 / eps);
            const double q = 2.0 * (double)j / lg;
            const double kap = q * q;
            if (kap > kmin) kmin = kap;
            iter = j;
            if (eps < tol || kmin > kappa_max || j >= maxiter) break;
        }
        lap(1);

        // ---- C: nu'(w) = p_w(A') conj(p_w)(A'^T) nu(w), one cluster per frequency (src/KPMPreconditioners.jl:606-679) -------
        {
            const int ncl = G >> 1, cl = cta >> 1;
            const double* in_comp = reinterpret_cast<const double*>(P.nu_in) + crank;
            double* out_comp = reinterpret_cast<double*>(P.nu_out) + crank;
            const double sc = 2.0 * P.K.inv_mag * P.K.cprod;
            for (int k = cl; k < P.Lo2; k += ncl) {
                const int w = P.K.schedule[k];
                const int order = P.K.order[w];
                for (int i = threadIdx.x; i < order; i += blockDim.x) c_s[i] = P.K.coeff[P.K.coeff_off[w] + i];
                Tile<NSEG, PY> v, A, B, evs, t1;
#pragma unroll
                for (int r = 0; r < PY; ++r)
#pragma unroll
                    for (int q = 0; q < NSEG; ++q) {
                        const size_t e = tile_off + r * LX + 32 * q + lane;
                        v.a[r][q] = __ldcg(in_comp + 2 * ((size_t)w * N + e));
                        evs.a[r][q] = (SSH ? 1.0 : sc) * P.K.eVbar[e];
                    }
                __syncthreads();
                // the two exchange buffers alternate: the partner writes buffer b again only after it has passed the cluster
                // barrier of the swap in between, which this CTA reaches after reading buffer b -- one barrier per swap
                auto swap_combine = [&](Tile<NSEG, PY>& out) {
                    const uint32_t boff = (uint32_t)((nswap & 1) * N);
#pragma unroll
                    for (int r = 0; r < PY; ++r)
#pragma unroll
                        for (int q = 0; q < NSEG; ++q) {
                            const uint32_t addr = remote_xch + (uint32_t)((boff + tile_off + r * LX + 32 * q + lane) * sizeof(double));
                            asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(B.a[r][q]) : "memory");
                        }
                    cluster_sync_all();
#pragma unroll
                    for (int r = 0; r < PY; ++r)
#pragma unroll
                        for (int q = 0; q < NSEG; ++q) {
                            const double bp = xch[boff + tile_off + r * LX + 32 * q + lane];
                            out.a[r][q] = (crank == 0) ? (A.a[r][q] - bp) : (A.a[r][q] + bp);
                        }
                    ++nswap;
                };
                int xbuf = 0;
                if constexpr (SSH) {
                    poly_real<NSEG, PY, true, true>(A, B, v, evs, c_s, order, P.K, strips, xbuf, warp, nwarps, lane, tabbar);
                    swap_combine(t1);
                    poly_real<NSEG, PY, false, true>(A, B, t1, evs, c_s, order, P.K, strips, xbuf, warp, nwarps, lane, tabbar);
                    swap_combine(v);
                } else {
                    poly_real_fast<NSEG, PY, true>(A, B, v, evs, c_s, order, P.K, strips, xbuf, warp, nwarps, lane);
                    swap_combine(t1);
                    poly_real_fast<NSEG, PY, false>(A, B, t1, evs, c_s, order, P.K, strips, xbuf, warp, nwarps, lane);
                    swap_combine(v);
                }
#pragma unroll
                for (int r = 0; r < PY; ++r)
#pragma unroll
                    for (int q = 0; q < NSEG; ++q) {
                        const size_t e = tile_off + r * LX + 32 * q + lane;
                        out_comp[2 * ((size_t)w * N + e)] = v.a[r][q];
                    }
                __syncthreads();   // c_s and the strips are rewritten by the next frequency
            }
        }
        lap(2);
        grid_sum(0.0, P.partial, P.bar, ++seq, red, bcast, &failed);
        if (failed) break;
        lap(3);

        // ---- I: z = Re(conj(theta) .* iFFT(nu')) / L, r.z (src/TimeFreqFFTs.jl:112-130, IterativeSolvers.jl:221-227) ---------
        acc = 0.0;
        for (int grp = cta; grp < ngroups; grp += G) {
            const int gsite = grp * SB + site;
            const bool ok = gsite < N;
            if (half) {
                // the mirror symmetry nu'(L-1-w) = conj(nu'(w)) folds the sum over all frequencies onto the even ones:
                //     z_t = (2/L) Re[ conj(theta_t) S_t ],  z_{t+L/2} = -(2/L) Im[ conj(theta_t) S_t ],
                //     S = inverse transform of length L/2 of C'_k = nu'(2k)   (= conj(nu'(L-1-2k)) where 2k >= L/2)
                for (int k0 = slot; k0 < Lh; k0 += 4 * nslots) {
                    cplx vv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int kk = k0 + k * nslots;
                        vv[k] = make_double2(0.0, 0.0);
                        if (ok && kk < Lh) {
                            const int w = 2 * kk;
                            if (w < P.Lo2) {
                                vv[k] = __ldcg(P.nu_out + (size_t)w * N + gsite);
                            } else {
                                vv[k] = __ldcg(P.nu_out + (size_t)(L - 1 - w) * N + gsite);
                                vv[k].y = -vv[k].y;
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int kk = k0 + k * nslots;
                        if (kk < Lh) b0[(size_t)kk * SB + site] = vv[k];
                    }
                }
                __syncthreads();
                const long long tq0 = clock64();
                cplx* res = fft_smem_shared<SB>(b0, b1, &plans[1], tw2, true);
                tf[1] += (unsigned long long)(clock64() - tq0);
                const double sc2 = 2.0 * invL;
                for (int t0 = slot; t0 < Lh; t0 += 2 * nslots) {
                    double rv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int t = t0 + (k >> 1) * nslots;
                        rv[k] = (ok && t < Lh) ? P.r[(size_t)(t + (k & 1) * Lh) * N + gsite] : 0.0;
                    }
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int t = t0 + k * nslots;
                        if (ok && t < Lh) {
                            const cplx v = res[(size_t)t * SB + site];
                            const cplx tt = th[t];
                            const double z1 = (tt.x * v.x + tt.y * v.y) * sc2;       // Re(conj(theta) S)
                            const double z2 = -(tt.x * v.y - tt.y * v.x) * sc2;      // -Im(conj(theta) S)
                            P.z[(size_t)t * N + gsite] = z1;
                            P.z[(size_t)(t + Lh) * N + gsite] = z2;
                            acc = fma(rv[2 * k], z1, acc);
                            acc = fma(rv[2 * k + 1], z2, acc);
                        }
                    }
                }
                __syncthreads();
                continue;
            }
            for (int w0 = slot; w0 < P.Lo2; w0 += 4 * nslots) {
                cplx vv[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int w = w0 + k * nslots;
                    vv[k] = make_double2(0.0, 0.0);
                    if (ok && w < P.Lo2) vv[k] = __ldcg(P.nu_out + (size_t)w * N + gsite);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int w = w0 + k * nslots;
                    if (w < P.Lo2) {
                        const int wm = L - 1 - w;
                        // the mirror frequency is the complex conjugate; for odd Ltau the middle frequency is its own mirror and
                        // the reference's loop leaves the conjugate there (kpm_square.cu writes it the same way)
                        if (wm != w) b0[(size_t)w * SB + site] = vv[k];
                        b0[(size_t)wm * SB + site] = make_double2(vv[k].x, -vv[k].y);
                    }
                }
            }
            __syncthreads();
            const long long tq0 = clock64();
            cplx* res = fft_smem_shared<SB>(b0, b1, &plans[0], tw, true);
            tf[1] += (unsigned long long)(clock64() - tq0);
            for (int t0 = slot; t0 < L; t0 += 4 * nslots) {
                double rv[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int t = t0 + k * nslots;
                    rv[k] = (ok && t < L) ? P.r[(size_t)t * N + gsite] : 0.0;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int t = t0 + k * nslots;
                    if (ok && t < L) {
                        const cplx v = res[(size_t)t * SB + site];
                        const cplx tt = th[t];
                        const double zz = (tt.x * v.x + tt.y * v.y) * invL;
                        P.z[(size_t)t * N + gsite] = zz;
                        acc = fma(rv[k], zz, acc);
                    }
                }
            }
            __syncthreads();
        }
        lap(0);
        const double rz = grid_sum(acc, P.partial, P.bar, ++seq, red, bcast, &failed);
        if (failed) break;
        beta = first ? 0.0 : rz / rdotz;
        rdotz = rz;
        if (!first) par ^= 1;
        lap(1);

        // ---- A: p = z + beta p (:228), Ap = M^T M p (src/Models.jl:215-224), p.Ap = |M p|^2 -------------------------------------
        acc = 0.0;
        {
            const double* pold = P.p[par ^ 1];
            double* pnew = P.p[par];
            int xbuf = 0;
            double above[NSEG], below[NSEG];
            for (int c = cta; c < P.nchunks; c += G) {
                const int a = c * P.Cs;
                const int nout = min(P.Cs, L - a);
                Tile<NSEG, PY> vprev, wprev, t, u, vc, dt;
                {
                    const int taum = (a == 0) ? L - 1 : a - 1;
                    const size_t g = (size_t)taum * N + tile_off;
#pragma unroll
                    for (int r = 0; r < PY; ++r)
#pragma unroll
                        for (int q = 0; q < NSEG; ++q) {
                            const size_t e = g + r * LX + 32 * q + lane;
                            vprev.a[r][q] = fma(beta, __ldcg(pold + e), __ldcg(P.z + e));
                            wprev.a[r][q] = 0.0;
                        }
                }
                for (int j = 0; j <= nout; ++j) {
                    int tau = a + j;
                    if (tau >= L) tau -= L;
                    const bool wrap = (tau == 0);   // antiperiodic boundary (src/HolsteinModels.jl:594-601)
                    const size_t g = (size_t)tau * N + tile_off;
#pragma unroll
                    for (int r = 0; r < PY; ++r)
#pragma unroll
                        for (int q = 0; q < NSEG; ++q) {
                            const size_t e = g + r * LX + 32 * q + lane;
                            vc.a[r][q] = fma(beta, __ldcg(pold + e), __ldcg(P.z + e));
                            dt.a[r][q] = SSH ? P.D[tile_off + r * LX + 32 * q + lane] : P.D[e];
                            t.a[r][q] = dt.a[r][q] * vprev.a[r][q];
                        }
                    // SSH: K(tau) from the CTA's resident tables (x bonds, y bonds, the y row above the tile)
                    const int y0 = warp * PY;
                    const double2* txs = tabA + (size_t)j * 2 * N + (size_t)y0 * LX;
                    const double2* tys = tabA + (size_t)j * 2 * N + N + (size_t)y0 * LX;
                    const double2* hys = tabA + (size_t)j * 2 * N + N + (size_t)((y0 + P.Ly - 1) % P.Ly) * LX;
                    if constexpr (SSH) {
                        g0_tab(t, txs, lane);
                        g1_tab(t, txs, lane);
                        g2_tab(t, tys, lane);
                    } else {
                        g0_x_even(t, P.K.c0, P.K.s0);
                        g1_x_odd(t, P.K.c1, P.K.s1, lane);
                        g2_y_even(t, P.K.c2, P.K.s2);
                    }
                    exchange_edges1(t, strips + (size_t)xbuf * nwarps * 2 * LX, warp, nwarps, lane, above, below);
                    xbuf ^= 1;
                    if constexpr (SSH) g3_tab(t, tys, hys, lane, above, below);
                    else g3_y_odd(t, P.K.c3, P.K.s3, above, below);
#pragma unroll
                    for (int r = 0; r < PY; ++r)
#pragma unroll
                        for (int q = 0; q < NSEG; ++q) {
                            const size_t e = g + r * LX + 32 * q + lane;
                            if (j < nout) pnew[e] = vc.a[r][q];
                            const double w = wrap ? (vc.a[r][q] + t.a[r][q]) : (vc.a[r][q] - t.a[r][q]);
                            t.a[r][q] = w;
                            vprev.a[r][q] = vc.a[r][q];
                            if (j < nout) acc = fma(w, w, acc);
                        }
                    if (j >= 1) {
#pragma unroll
                        for (int r = 0; r < PY; ++r)
#pragma unroll
                            for (int q = 0; q < NSEG; ++q) u.a[r][q] = t.a[r][q];
                        exchange_edges1(u, strips + (size_t)xbuf * nwarps * 2 * LX, warp, nwarps, lane, above, below);
                        xbuf ^= 1;
                        if constexpr (SSH) {
                            g3_tab(u, tys, hys, lane, above, below);
                            g2_tab(u, tys, lane);
                            g1_tab(u, txs, lane);
                            g0_tab(u, txs, lane);
                        } else {
                            g3_y_odd(u, P.K.c3, P.K.s3, above, below);
                            g2_y_even(u, P.K.c2, P.K.s2);
                            g1_x_odd(u, P.K.c1, P.K.s1, lane);
                            g0_x_even(u, P.K.c0, P.K.s0);
                        }
                        const size_t gm = (size_t)(a + j - 1) * N + tile_off;
#pragma unroll
                        for (int r = 0; r < PY; ++r)
#pragma unroll
                            for (int q = 0; q < NSEG; ++q) {
                                const int e = r * LX + 32 * q + lane;
                                const double du = dt.a[r][q] * u.a[r][q];
                                P.ap[gm + e] = wrap ? (wprev.a[r][q] + du) : (wprev.a[r][q] - du);
                            }
                    }
#pragma unroll
                    for (int r = 0; r < PY; ++r)
#pragma unroll
                        for (int q = 0; q < NSEG; ++q) wprev.a[r][q] = t.a[r][q];
                }
            }
        }
        lap(4);
        pAp = grid_sum(acc, P.partial, P.bar, ++seq, red, bcast, &failed);
        if (failed) break;
        alpha = rdotz / pAp;
        first = false;
        lap(5);
    }
    if (cta == 0 && threadIdx.x == 0) {
        CgScalars* S = P.S;
        S->iter = iter;
        S->eps = eps;
        S->kappa_min = kmin;
        S->alpha = alpha;
        S->beta = beta;
        S->rdotz = rdotz;
        S->pAp = pAp;
        S->done = failed ? 2 : 1;
    }
    if (P.prof && threadIdx.x == 0 && cta < 4)
    {
        for (int k = 0; k < 6; ++k) P.prof[8 * cta + k] = tph[k];
        P.prof[8 * cta + 6] = tf[0];
        P.prof[8 * cta + 7] = tf[1];
    }
}

template <int NSEG, int PY, int SB>
size_t fused_smem(int L, int Ly, int nwarps, int max_order, int ssh_slices) {
    constexpr int LX = 32 * NSEG;
    const size_t fft = 2ull * L * SB * sizeof(cplx);
    const size_t chain = (size_t)max_order * sizeof(cplx) + (2ull * nwarps * 2 * LX + 2ull * LX * Ly) * sizeof(double);
    const size_t region = (std::max(fft, chain) + 15) & ~size_t(15);
    // SSH: the tau-averaged tables + the tables of the CTA's slices (chunk + halo), 2 N (cosh, sinh) pairs each
    const size_t tabs = ssh_slices ? (size_t)(1 + ssh_slices) * 2 * LX * Ly * sizeof(double2) : 0;
    return (2ull * L + (L + 1) / 2) * sizeof(cplx) + region + tabs;
}

template <int NSEG, int PY, int SB, int MAXT, bool SSH = false>
bool launch_fused(elph_handle* h, FusedParams& P, int nwarps) {
    auto kern = pcg_fused_kernel<NSEG, PY, SB, MAXT, SSH>;
    // SSH keeps the tables of a CTA's slices resident: the chunk length is fixed by the full grid (one CTA per SM)
    int full = (h->spec_running ? h->sm_count - 2 : h->sm_count) & ~1;   // a speculative set-up runs its Arnoldi kernel on two SMs
    if (h->pcg_grid >= 2) full = std::min(full, h->pcg_grid & ~1);
    const int cs_full = (h->L + full - 1) / full;
    const size_t smem = fused_smem<NSEG, PY, SB>(h->L, P.Ly, nwarps, P.max_order, SSH ? cs_full + 1 : 0);
    if (smem > h->smem_optin) return false;
    elph_enable_smem(h, kern);
    const int threads = nwarps * 32;
    // one CTA per SM, an even number of them (2-CTA clusters), all co-resident
    int grid = (h->spec_running ? h->sm_count - 2 : h->sm_count) & ~1;
    if (h->pcg_grid >= 2) grid = std::min(grid, h->pcg_grid & ~1);   // tuning key 20: leave SMs to other chains on the same GPU
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return false; }
    if (nclusters < 1) return false;
    grid = std::min(grid, 2 * nclusters);
    if (SSH && (h->L + grid - 1) / grid != cs_full) return false;   // fewer co-resident CTAs than assumed: tables would not fit
    cfg.gridDim = dim3(grid);
    // slices per chunk of the product phase: every CTA at most one chunk
    P.Cs = (h->L + grid - 1) / grid;
    P.nchunks = (h->L + P.Cs - 1) / P.Cs;
    ELPH_REQUIRE(2 * grid <= h->partial_cap, ELPH_ERR_STATE, "partial-sum buffer too small for the fused PCG");
    ELPH_CUDA(cudaMemsetAsync(h->d_bar, 0, sizeof(unsigned int), h->stream));
    at[1].id = cudaLaunchAttributeCooperative;
    at[1].val.cooperative = 1;
    // ELPH_PCG_NOCOOP=1 (profiling aid): Nsight Compute refuses launches that are both cooperative and clustered; the occupancy
    // query above already guarantees that the whole grid is co-resident
    static const bool nocoop = [] { const char* e = getenv("ELPH_PCG_NOCOOP"); return e && e[0] == '1'; }();
    cfg.numAttrs = nocoop ? 1 : 2;
    void* args[] = {&P};
    cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)kern, args);
    if (e != cudaSuccess) {
        // cooperative + cluster not accepted by this driver: the occupancy query above already guarantees co-residency
        cudaGetLastError();
        cfg.numAttrs = 1;
        e = cudaLaunchKernelExC(&cfg, (const void*)kern, args);
    }
    ELPH_CUDA(e);
    h->launches++;
    return true;
}

FftPlan elph_fft_plan(const elph_handle* h) {
    FftPlan p;
    p.L = h->L;
    p.nrad = (int)h->fft_radices.size();
    for (int i = 0; i < p.nrad; ++i) p.rad[i] = h->fft_radices[i];
    return p;
}

}  // namespace

// The whole preconditioned solve after cg_init_kernel (r = b - A x0, the norms and the stop-rule constants in h->d_cg) in one
// launch.  Returns false when the configuration is not served (the caller runs the launch-per-phase loop).
bool elph_pcg_fused(elph_handle* h, double* x_dev, double* z_dev) {
    const bool ssh = (h->model == ELPH_MODEL_SSH);
    if (!h->pcg_persistent || h->sq_disable || h->sharded) return false;
    if (ssh ? !h->ssq.enabled : !h->sq.enabled) return false;
    const KpmState& K = h->kpm;
    if (!K.active || !K.d_coeff) return false;
    if (ssh && !K.d_csbar_tile) return false;
    const int Lx = ssh ? h->ssq.Lx : h->sq.Lx, Ly = ssh ? h->ssq.Ly : h->sq.Ly;
    if (Lx != 32 || Ly % 2 || Ly / 2 < 2 || Ly / 2 > 32 || h->L < 4) return false;
    const int nwarps = Ly / 2;
    int max_order = 1;
    for (int w = 0; w < K.Lo2; ++w) max_order = std::max(max_order, K.order[w]);
    FusedParams P;
    P.x = x_dev; P.r = h->d_r; P.p[0] = h->d_p[0]; P.p[1] = h->d_p[1]; P.ap = h->d_z; P.z = z_dev;
    P.nu_in = K.d_nu; P.nu_out = h->d_nu2; P.D = h->d_D;
    P.plan = elph_fft_plan(h);
    P.plan_half.L = 0;
    P.plan_half.nrad = 0;
    if (h->L % 2 == 0 && h->pcg_half_fft) {   // same factorisation rule as elph_fft_init: 4 first, then the primes in ascending order
        FftPlan& q = P.plan_half;
        int n = h->L / 2;
        q.L = n;
        while (n % 4 == 0 && q.nrad < kMaxRad) { q.rad[q.nrad++] = 4; n /= 4; }
        for (int f = 2; (long long)f * f <= n; ++f)
            while (n % f == 0 && q.nrad < kMaxRad) { q.rad[q.nrad++] = f; n /= f; }
        if (n > 1 && q.nrad < kMaxRad) { q.rad[q.nrad++] = n; n = 1; }
        if (n != 1 || q.L < 2) q.L = 0;
    }
    P.tw = h->d_twiddle; P.theta = h->d_theta;
    KsqParams& Q = P.K;
    Q.in = nullptr; Q.out = nullptr; Q.eVbar = K.d_eVbar; Q.coeff = K.d_coeff; Q.order = K.d_order; Q.coeff_off = K.d_coeff_off;
    Q.schedule = K.d_schedule; Q.skip = nullptr; Q.L = h->L; Q.Ly = Ly;
    Q.inv_mag = 1.0 / K.lam_mag; Q.avg_over_mag = K.lam_avg / K.lam_mag;
    if (ssh) {
        Q.c0 = Q.c1 = Q.c2 = Q.c3 = 1.0; Q.s0 = Q.s1 = Q.s2 = Q.s3 = 0.0;
    } else {
        Q.c0 = h->sq.c[0]; Q.s0 = h->sq.s[0]; Q.c1 = h->sq.c[1]; Q.s1 = h->sq.s[1];
        Q.c2 = h->sq.c[2]; Q.s2 = h->sq.s[2]; Q.c3 = h->sq.c[3]; Q.s3 = h->sq.s[3];
    }
    Q.t0 = Q.s0 / Q.c0; Q.t1 = Q.s1 / Q.c1; Q.t2 = Q.s2 / Q.c2; Q.t3 = Q.s3 / Q.c3;
    Q.cprod = Q.c0 * Q.c1 * Q.c2 * Q.c3;
    Q.fast = ssh ? 0 : 1; Q.prof = nullptr; Q.tab = ssh ? K.d_csbar_tile : nullptr;
    P.ssh_tab = ssh ? h->ssq.d_tab : nullptr;
    P.Lo2 = K.Lo2; P.max_order = max_order;
    P.partial = h->d_partial; P.bar = h->d_bar; P.S = h->d_cg;
    P.L = h->L; P.Ly = Ly; P.Cs = 1; P.nchunks = h->L;
    P.prof = h->pipe_prof ? h->pipe_prof_buf : nullptr;
    ELPH_CUDA(cudaMemsetAsync(h->d_p[1], 0, h->Ndim * sizeof(double), h->stream));   // p_old of the first product (beta = 0)
    if (ssh) return (nwarps * 32 <= 512) ? launch_fused<1, 2, 8, 512, true>(h, P, nwarps) : false;
    if (nwarps * 32 <= 512) return launch_fused<1, 2, 8, 512>(h, P, nwarps);
    return launch_fused<1, 2, 8, 1024>(h, P, nwarps);
}
