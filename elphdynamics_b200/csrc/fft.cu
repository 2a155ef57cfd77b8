// K4: batched shared-memory FFT along imaginary time, fused with the reference's
// pre/post operations.
//
// Reference: src/TimeFreqFFTs.jl:32-45 (Theta twist + plan), :55-73 tau_to_omega!,
// :112-130 omega_to_tau! (real part); src/FourierAcceleration.jl:91-143
// fourier_accelerate! (plain FFT, multiply by Q^p or M^p, inverse FFT, real part).
// FFTW conventions (FFTW.jl 1.3.2): forward sum_j x_j exp(-2 pi i j k/L) unnormalised,
// inverse scaled by 1/L.
//
// Engine layout is [tau][site]; one CTA owns SB consecutive sites for ALL tau, so
// global accesses are coalesced along the site index and the transform runs entirely
// in shared memory ([tau][SB] complex, site fastest => conflict-free butterflies,
// twiddles warp-uniform and staged in shared memory).  Lengths are arbitrary: a
// mixed-radix Stockham autosort (decimation in frequency).  Radices 2, 3, 4, 5 are
// register butterflies (one thread = one butterfly); any other prime factor uses a
// generic O(r)-per-output stage.  Ltau = 20, 40, 200, 400 factor into {4, 2, 5}.
#include "elph_internal.cuh"
#include "fft_smem.cuh"

#include <cmath>

namespace {

constexpr int kT = 256;
using namespace fftsm;

// mode 0: tau_to_omega  (real in, complex out):  out = FFT(theta .* in)
// mode 1: omega_to_tau  (complex in, real out):  out = Re(conj(theta) .* iFFT(in))
// mode 2: fourier accelerate (real in, real out): out = Re(iFFT(diag^power .* FFT(in)))
// CG fusion (preconditioned iteration, cg.cu): with F.S set, mode 0 first performs x += alpha p, r -= alpha Ap on the
// elements it loads (every element belongs to exactly one thread), transforms the NEW r, and the last CTA applies the
// stop rule; mode 1 accumulates r.z for the vector z it produces and the last CTA publishes beta.
struct CgFuse {
    double* x;
    double* r;
    const double* p;
    const double* ap;
    double* partial;
    CgScalars* S;
    unsigned int* ticket;
};

__device__ __forceinline__ double fft_block_sum(double x, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = x;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
    return t;
}

// per-CTA partial -> last CTA folds all partials in index order; returns true (+ the sum on thread 0) in the last CTA
__device__ __forceinline__ bool fft_last_block_sum(double acc, const CgFuse& F, double* red, bool* flag, double* total) {
    const double t = fft_block_sum(acc, red);
    if (threadIdx.x == 0) {
        F.partial[blockIdx.x] = t;
        __threadfence();
        const unsigned int n = atomicAdd(F.ticket, 1u);
        *flag = (n == gridDim.x - 1);
    }
    __syncthreads();
    if (!*flag) return false;
    __threadfence();
    double s = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) s += ((const volatile double*)F.partial)[k];
    *total = fft_block_sum(s, red);
    return true;
}

template <int SB>
// rin / rout and cin / cout may alias (fourier_accelerate!(eta, fa, eta, ...) works in place: a column is read completely into
// shared memory before it is written back): no __restrict__ on them.
__global__ void __launch_bounds__(kT) fft_kernel(int mode, const double* rin, const cplx* cin, double* rout, cplx* cout, FftPlan plan, int N,
                                                 const cplx* __restrict__ tw_g, const cplx* __restrict__ theta,
                                                 const double* __restrict__ diag, double power, const int* skip, CgFuse F) {
    extern __shared__ __align__(16) double smem_raw[];
    __shared__ double red[32];
    __shared__ bool lastflag;
    if (skip && *skip) return;
    double fuse_acc = 0.0;
    const double alpha = (F.S && mode == 0) ? F.S->alpha : 0.0;
    const int L = plan.L;
    cplx* b0 = reinterpret_cast<cplx*>(smem_raw);
    cplx* b1 = b0 + (size_t)L * SB;
    cplx* tw = b1 + (size_t)L * SB;  // [L] twiddles staged once per CTA
    const int site = threadIdx.x % SB;
    const int slot = threadIdx.x / SB;
    const int nslots = blockDim.x / SB;
    const int gsite = blockIdx.x * SB + site;
    const bool ok = gsite < N;
    const double invL = 1.0 / (double)L;

    for (int k = threadIdx.x; k < L; k += blockDim.x) tw[k] = tw_g[k];
    for (int t = slot; t < L; t += nslots) {
        cplx v = make_double2(0.0, 0.0);
        if (ok) {
            if (mode == 0) {
                double xr;
                if (F.S) {   // fused x += alpha p ; r -= alpha Ap ; |r|^2   (src/IterativeSolvers.jl:205-211)
                    const size_t e = (size_t)t * N + gsite;
                    F.x[e] = fma(alpha, F.p[e], F.x[e]);
                    xr = fma(-alpha, F.ap[e], F.r[e]);
                    F.r[e] = xr;
                    fuse_acc = fma(xr, xr, fuse_acc);
                } else {
                    xr = rin[(size_t)t * N + gsite];
                }
                const cplx th = theta[t];
                v = make_double2(th.x * xr, th.y * xr);
            } else if (mode == 1) {
                v = cin[(size_t)t * N + gsite];
            } else {
                v = make_double2(rin[(size_t)t * N + gsite], 0.0);
            }
        }
        b0[(size_t)t * SB + site] = v;
    }
    __syncthreads();
    cplx* res = fft_smem_auto<SB>(b0, b1, plan, tw, mode == 1);
    if (mode == 0) {
        for (int t = slot; t < L; t += nslots)
            if (ok) cout[(size_t)t * N + gsite] = res[(size_t)t * SB + site];
        if (F.S) {
            double rr;
            if (fft_last_block_sum(fuse_acc, F, red, &lastflag, &rr) && threadIdx.x == 0) {
                // stop rule, preconditioned variant (src/IterativeSolvers.jl:208-219): beta is set later from r.z
                CgScalars* S = F.S;
                const long long j = S->iter + 1;
                const double eps = sqrt(rr) / S->normb;
                const double lg = log(2.0 * S->eps0 / eps);
                const double q = 2.0 * (double)j / lg;
                const double kap = q * q;
                double kmin = S->kappa_min;
                if (kap > kmin) kmin = kap;
                S->kappa_min = kmin;
                S->eps = eps;
                S->iter = j;
                if (eps < S->tol || kmin > S->kappa_max || j >= S->maxiter) S->done = 1;
                *F.ticket = 0u;
            }
        }
        return;
    }
    if (mode == 1) {
        for (int t = slot; t < L; t += nslots) {
            if (ok) {
                const cplx v = res[(size_t)t * SB + site];
                const cplx th = theta[t];  // conj(theta) * v, real part
                const double z = (th.x * v.x + th.y * v.y) * invL;
                rout[(size_t)t * N + gsite] = z;
                if (F.S) fuse_acc = fma(F.r[(size_t)t * N + gsite], z, fuse_acc);
            }
        }
        if (F.S) {
            double rz;
            if (fft_last_block_sum(fuse_acc, F, red, &lastflag, &rz) && threadIdx.x == 0) {
                // beta = (r.z)_new / (r.z)_old   (src/IterativeSolvers.jl:221-227)
                CgScalars* S = F.S;
                S->beta = (S->iter == 0) ? 0.0 : rz / S->rdotz;
                S->rdotz = rz;
                *F.ticket = 0u;
            }
        }
        return;
    }
    // mode 2: scale in frequency space, inverse transform, real part
    cplx* other = (res == b0) ? b1 : b0;
    for (int t = slot; t < L; t += nslots) {
        double d = ok ? diag[(size_t)t * N + gsite] : 1.0;
        double f;
        if (power == 1.0) f = d;
        else if (power == 0.5) f = sqrt(d);
        else if (power == -1.0) f = 1.0 / d;
        else if (power == -0.5) f = 1.0 / sqrt(d);
        else f = pow(d, power);
        cplx v = res[(size_t)t * SB + site];
        res[(size_t)t * SB + site] = make_double2(v.x * f, v.y * f);
    }
    __syncthreads();
    cplx* res2 = fft_smem_auto<SB>(res, other, plan, tw, true);
    for (int t = slot; t < L; t += nslots)
        if (ok) rout[(size_t)t * N + gsite] = res2[(size_t)t * SB + site].x * invL;
}

FftPlan make_plan(const elph_handle* h) {
    FftPlan p;
    p.L = h->L;
    p.nrad = (int)h->fft_radices.size();
    for (int i = 0; i < p.nrad; ++i) p.rad[i] = h->fft_radices[i];
    return p;
}

// CG fusion request for the next launch (set and cleared by the *_cg entry points below; one caller thread per handle)
thread_local CgFuse g_fuse = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

template <int SB>
size_t fft_smem_bytes(int L) { return (2ull * L * SB + L) * sizeof(cplx); }

template <int SB>
void launch_fft(elph_handle* h, int mode, int ncols, const double* rin, const cplx* cin, double* rout, cplx* cout,
                const double* diag, double power, const int* skip) {
    const size_t smem = fft_smem_bytes<SB>(h->L);
    ELPH_REQUIRE(smem <= h->smem_optin, ELPH_ERR_UNSUPPORTED, "Ltau too large for the shared-memory FFT");
    elph_enable_smem(h, fft_kernel<SB>);
    const int blocks = (ncols + SB - 1) / SB;
    fft_kernel<SB><<<blocks, kT, smem, h->stream>>>(mode, rin, cin, rout, cout, make_plan(h), ncols, h->d_twiddle, h->d_theta,
                                                     diag, power, skip, g_fuse);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}


// Inner loop of the multi-timestep HMC integrator (src/HMC.jl:556-600) for SB phonon columns per CTA, entirely in shared
// memory: the bosonic force dSb/dx couples only the time slices of ONE phonon column and the mass-matrix acceleration
// is a tau-FFT of that column, so a column never needs another one.
//     y = M^-1 dSb/dx ;  repeat Nb times:  v -= dt'/2 y ;  x += dt' v ;  y = M^-1 dSb/dx ;  v -= dt'/2 y
// One launch instead of 6 Nb + 3 (memset, dSb, FFT pair and lincombs per inner step); the arithmetic is that of
// dSb_kernel / fft_kernel mode 2 / lincomb_kernel operation for operation (trajectories agree to the last bit or two:
// only FMA contraction differs between the two compilations).  Config D: 18.7 -> 16.1 ms per trajectory, 793 -> 283 launches.
template <int SB>
__global__ void __launch_bounds__(kT) hmc_inner_kernel(double* __restrict__ xg, double* __restrict__ vg, FftPlan plan, int N,
                                                       const cplx* __restrict__ tw_g, const double* __restrict__ diag,
                                                       const double* __restrict__ omega, const double* __restrict__ omega4,
                                                       double dtau, double dtp, int Nb) {
    extern __shared__ __align__(16) double smem_raw[];
    const int L = plan.L;
    cplx* b0 = reinterpret_cast<cplx*>(smem_raw);
    cplx* b1 = b0 + (size_t)L * SB;
    cplx* tw = b1 + (size_t)L * SB;
    double* xs = reinterpret_cast<double*>(tw + L);   // [L][SB]
    double* vs = xs + (size_t)L * SB;
    double* ys = vs + (size_t)L * SB;
    double* fs = ys + (size_t)L * SB;                 // 1 / M(k) of the column
    const int site = threadIdx.x % SB;
    const int slot = threadIdx.x / SB;
    const int nslots = blockDim.x / SB;
    const int gsite = blockIdx.x * SB + site;
    const bool ok = gsite < N;
    const double invL = 1.0 / (double)L;
    const double w = ok ? omega[gsite] : 0.0, w4 = ok ? omega4[gsite] : 0.0;
    for (int k = threadIdx.x; k < L; k += blockDim.x) tw[k] = tw_g[k];
    for (int t = slot; t < L; t += nslots) {
        const size_t e = (size_t)t * SB + site;
        xs[e] = ok ? xg[(size_t)t * N + gsite] : 0.0;
        vs[e] = ok ? vg[(size_t)t * N + gsite] : 0.0;
        fs[e] = ok ? 1.0 / diag[(size_t)t * N + gsite] : 1.0;
    }
    __syncthreads();
    auto boson_force = [&]() {      // ys = Re(iFFT(M^-1 .* FFT(dSb/dx(xs)))) / L ; xs must be complete (synchronised)
        for (int t = slot; t < L; t += nslots) {
            const double d = dSb_term(xs, t, site, SB, L, dtau, w, w4, 0.0);
            b0[(size_t)t * SB + site] = make_double2(0.0 + d, 0.0);
        }
        __syncthreads();
        cplx* res = fft_smem_auto<SB>(b0, b1, plan, tw, false);
        cplx* other = (res == b0) ? b1 : b0;
        for (int t = slot; t < L; t += nslots) {
            const double f = fs[(size_t)t * SB + site];
            const cplx v = res[(size_t)t * SB + site];
            res[(size_t)t * SB + site] = make_double2(v.x * f, v.y * f);
        }
        __syncthreads();
        cplx* res2 = fft_smem_auto<SB>(res, other, plan, tw, true);
        for (int t = slot; t < L; t += nslots) ys[(size_t)t * SB + site] = res2[(size_t)t * SB + site].x * invL;
        // every thread reads back only the ys it wrote; b0 / b1 are rewritten after the next __syncthreads
    };
    boson_force();
    for (int it = 0; it < Nb; ++it) {
        for (int t = slot; t < L; t += nslots) {
            const size_t e = (size_t)t * SB + site;
            const double vn = fma(-dtp / 2, ys[e], 1.0 * vs[e]);     // lincomb(v, 1, v, -dt'/2, y)
            vs[e] = vn;
            xs[e] = fma(dtp, vn, 1.0 * xs[e]);                       // lincomb(x, 1, x, dt', v)
        }
        __syncthreads();
        boson_force();
        for (int t = slot; t < L; t += nslots) {
            const size_t e = (size_t)t * SB + site;
            vs[e] = fma(-dtp / 2, ys[e], 1.0 * vs[e]);
        }
    }
    for (int t = slot; t < L; t += nslots) {
        if (ok) {
            xg[(size_t)t * N + gsite] = xs[(size_t)t * SB + site];
            vg[(size_t)t * N + gsite] = vs[(size_t)t * SB + site];
        }
    }
}

template <int SB>
bool launch_hmc_inner(elph_handle* h, double* x, double* v, double dtp, int Nb) {
    const size_t smem = fft_smem_bytes<SB>(h->L) + 4ull * h->L * SB * sizeof(double);
    if (smem > h->smem_optin) return false;
    elph_enable_smem(h, hmc_inner_kernel<SB>);
    const int blocks = (h->Nph + SB - 1) / SB;
    hmc_inner_kernel<SB><<<blocks, kT, smem, h->stream>>>(x, v, make_plan(h), h->Nph, h->d_twiddle, h->d_Mass, h->d_omega, h->d_omega4,
                                                           h->dtau, dtp, Nb);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
    return true;
}

void dispatch_fft(elph_handle* h, int mode, int ncols, const double* rin, const cplx* cin, double* rout, cplx* cout,
                  const double* diag, double power, const int* skip = nullptr) {
    // fewer sites per CTA -> more CTAs: aim at ~2 CTAs per SM, bounded below by 4 sites and above by shared memory
    int sb = 32;
    while (sb > 4 && (ncols + sb - 1) / sb < 2 * h->sm_count) sb >>= 1;
    while (sb > 4 && (2ull * h->L * sb + h->L) * sizeof(cplx) > h->smem_optin) sb >>= 1;
    switch (sb) {
        case 32: launch_fft<32>(h, mode, ncols, rin, cin, rout, cout, diag, power, skip); break;
        case 16: launch_fft<16>(h, mode, ncols, rin, cin, rout, cout, diag, power, skip); break;
        case 8: launch_fft<8>(h, mode, ncols, rin, cin, rout, cout, diag, power, skip); break;
        default: launch_fft<4>(h, mode, ncols, rin, cin, rout, cout, diag, power, skip); break;
    }
}

}  // namespace

void elph_fft_init(elph_handle* h) {
    const int L = h->L;
    // factor L: 4 first, then 2, 3, 5, then remaining primes
    std::vector<int> rad;
    int n = L;
    while (n % 4 == 0) { rad.push_back(4); n /= 4; }
    for (int f = 2; (long long)f * f <= n; ++f)
        while (n % f == 0) { rad.push_back(f); n /= f; }
    if (n > 1) rad.push_back(n);
    ELPH_REQUIRE((int)rad.size() <= kMaxRad, ELPH_ERR_UNSUPPORTED, "Ltau has too many prime factors");
    h->fft_radices = rad;
    std::vector<cplx> tw(L), th(L);
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int k = 0; k < L; ++k) {
        const long double a = -2.0L * pi * (long double)k / (long double)L;
        tw[k] = make_double2((double)cosl(a), (double)sinl(a));
        const long double b = -pi * (long double)k / (long double)L;
        th[k] = make_double2((double)cosl(b), (double)sinl(b));
    }
    h->d_twiddle = elph_dalloc<cplx>(L);
    h->d_theta = elph_dalloc<cplx>(L);
    ELPH_CUDA(cudaMemcpy(h->d_twiddle, tw.data(), L * sizeof(cplx), cudaMemcpyHostToDevice));
    ELPH_CUDA(cudaMemcpy(h->d_theta, th.data(), L * sizeof(cplx), cudaMemcpyHostToDevice));
}

void elph_tau_to_omega_dev_skip(elph_handle* h, const double* vin, cplx* vout, const int* skip) {
    dispatch_fft(h, 0, h->N, vin, nullptr, nullptr, vout, nullptr, 0.0, skip);
}
void elph_omega_to_tau_dev_skip(elph_handle* h, const cplx* vin, double* vout, const int* skip) {
    dispatch_fft(h, 1, h->N, nullptr, vin, vout, nullptr, nullptr, 0.0, skip);
}

// preconditioned CG, fused variants (see CgFuse): forward = x/r update + stop rule + FFT(r); inverse = z + r.z + beta
void elph_tau_to_omega_dev_cg(elph_handle* h, double* x, double* r, const double* p, const double* ap, cplx* vout) {
    const int blocks_max = (h->N + 3) / 4;
    ELPH_REQUIRE(blocks_max <= h->partial_cap, ELPH_ERR_INVALID, "partial buffer too small for the fused FFT");
    g_fuse = CgFuse{x, r, p, ap, h->d_partial, h->d_cg, h->d_ticket};
    try {
        dispatch_fft(h, 0, h->N, r, nullptr, nullptr, vout, nullptr, 0.0, &h->d_cg->done);
    } catch (...) {
        g_fuse = CgFuse{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        throw;
    }
    g_fuse = CgFuse{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
}
void elph_omega_to_tau_dev_cg(elph_handle* h, const cplx* vin, double* z, double* r) {
    g_fuse = CgFuse{nullptr, r, nullptr, nullptr, h->d_partial, h->d_cg, h->d_ticket};
    try {
        dispatch_fft(h, 1, h->N, nullptr, vin, z, nullptr, nullptr, 0.0, &h->d_cg->done);
    } catch (...) {
        g_fuse = CgFuse{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        throw;
    }
    g_fuse = CgFuse{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
}

void elph_tau_to_omega_dev(elph_handle* h, const double* vin, cplx* vout) {
    dispatch_fft(h, 0, h->N, vin, nullptr, nullptr, vout, nullptr, 0.0);
}

void elph_omega_to_tau_dev(elph_handle* h, const cplx* vin, double* vout) {
    dispatch_fft(h, 1, h->N, nullptr, vin, vout, nullptr, nullptr, 0.0);
}

// tau_to_omega! / omega_to_tau! of `ncols` columns ([tau][col] <-> [omega][col]): the site-sharded stage of the tau-sharded
// preconditioner (after the first all-to-all a rank holds ALL time slices of a subset of the sites)
void elph_tau_to_omega_cols_dev(elph_handle* h, const double* vin, cplx* vout, int ncols) {
    ELPH_REQUIRE(vin && vout && ncols >= 1, ELPH_ERR_INVALID, "bad arguments");
    dispatch_fft(h, 0, ncols, vin, nullptr, nullptr, vout, nullptr, 0.0);
}
void elph_omega_to_tau_cols_dev(elph_handle* h, const cplx* vin, double* vout, int ncols) {
    ELPH_REQUIRE(vin && vout && ncols >= 1, ELPH_ERR_INVALID, "bad arguments");
    dispatch_fft(h, 1, ncols, nullptr, vin, vout, nullptr, nullptr, 0.0);
}

// Fourier acceleration of `ncols` columns ([k][col] layout) with an explicit diagonal: used by the tau-sharded driver
// after the all-to-all transpose, when a rank holds ALL time slices of a subset of the sites.
void elph_fourier_accelerate_cols_dev(elph_handle* h, const double* vin, double* vout, int ncols, const double* diag, double power) {
    ELPH_REQUIRE(vin && vout && diag && ncols >= 1, ELPH_ERR_INVALID, "bad arguments");
    dispatch_fft(h, 2, ncols, vin, nullptr, vout, nullptr, diag, power);
}

void elph_fourier_accelerate_dev(elph_handle* h, const double* vin, double* vout, double power, bool use_mass) {
    const double* diag = use_mass ? h->d_Mass : h->d_Q;
    ELPH_REQUIRE(use_mass ? h->have_M : h->have_Q, ELPH_ERR_STATE, "fourier acceleration diagonal (fa_Q / fa_M) was not provided");
    dispatch_fft(h, 2, h->Nph, vin, nullptr, vout, nullptr, diag, power);
}

// The Nb inner steps of one outer step of the multi-timestep integrator in one launch (hmc_inner_kernel); false if the
// column does not fit in shared memory (the caller runs the step-by-step kernels).
bool elph_hmc_inner_dev(elph_handle* h, double* x, double* v, double dtp, int Nb) {
    ELPH_REQUIRE(h->have_M, ELPH_ERR_STATE, "fourier acceleration diagonal (fa_M) was not provided");
    int sb = 32;
    while (sb > 4 && (h->Nph + sb - 1) / sb < 2 * h->sm_count) sb >>= 1;
    while (sb > 4 && (2ull * h->L * sb + h->L) * sizeof(cplx) + 4ull * h->L * sb * sizeof(double) > h->smem_optin) sb >>= 1;
    switch (sb) {
        case 32: return launch_hmc_inner<32>(h, x, v, dtp, Nb);
        case 16: return launch_hmc_inner<16>(h, x, v, dtp, Nb);
        case 8: return launch_hmc_inner<8>(h, x, v, dtp, Nb);
        default: return launch_hmc_inner<4>(h, x, v, dtp, Nb);
    }
}
