// K5: KPM (Chebyshev) preconditioner for M^T M.
//
// Reference: src/KPMPreconditioners.jl
//   KPMExpansion :101-146, setup! :269-321, update_A! :332-381, expansion mul!/ldiv! :387-420,
//   apply ldiv! :426-481, SymmetricKPMPreconditioner mul! :606-679, mulA'! :685-693, mulA! :758-778,
//   kpm_coefficients! :789-839, arnoldi_eigenvalue_bounds! :845-942, scalar_invM :948-951.
//
// Split of work:
//  * tau-mean of the operator tables: device kernel (sequential-in-tau sum per site, like the reference loop).
//  * Arnoldi bounds (2 x <=20 single-slice products on an N-vector, Gram-Schmidt, eigenvalues of a <=20x20
//    Hessenberg matrix) and the Chebyshev coefficient tables: host, once per setup!.  The start vectors are
//    injected by the caller (the reference draws them from model.rng).
//  * apply: twisted FFT (fft.cu) -> one CTA per Matsubara-like frequency runs the two three-term recurrences
//    on an N-vector (thread-owned sites in registers, the checkerboard sweep in shared memory), longest
//    polynomial first -> inverse FFT.  In the engine layout the frequency-space vector is [omega][site], so
//    the reference's two (L,N)<->(N,L) transposes disappear.
#include "elph_internal.cuh"

#include <future>
#include <algorithm>
#include <cmath>
#include <numeric>

typedef std::complex<double> zc;

// ------------------------------------------------------------------------------------------------
// host: eigenvalues of a small real upper-Hessenberg matrix (replaces LAPACK eigvals!, :891,:935)
// Complex single-shift (Wilkinson) QR with Givens rotations and deflation.
// ------------------------------------------------------------------------------------------------
static std::vector<zc> hessenberg_eigvals(const std::vector<double>& hreal, int n) {
    std::vector<zc> H((size_t)n * n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) H[(size_t)i * n + j] = (i <= j + 1) ? zc(hreal[(size_t)i * n + j], 0.0) : zc(0.0, 0.0);
    auto at = [&](int i, int j) -> zc& { return H[(size_t)i * n + j]; };
    std::vector<zc> ev(n);
    double hnorm = 0.0;
    for (auto& z : H) hnorm += std::norm(z);
    hnorm = std::sqrt(hnorm);
    if (hnorm == 0.0) hnorm = 1.0;
    const double eps = 2.220446049250313e-16;
    int hi = n - 1;
    int iter = 0;
    while (hi >= 0) {
        if (hi == 0) {
            ev[0] = at(0, 0);
            break;
        }
        // look for a negligible sub-diagonal entry
        int lo = hi;
        while (lo > 0) {
            double s = std::abs(at(lo - 1, lo - 1)) + std::abs(at(lo, lo));
            if (s == 0.0) s = hnorm;
            if (std::abs(at(lo, lo - 1)) <= eps * s) {
                at(lo, lo - 1) = 0.0;
                break;
            }
            --lo;
        }
        if (lo == hi) {
            ev[hi] = at(hi, hi);
            --hi;
            iter = 0;
            continue;
        }
        // Wilkinson shift from the trailing 2x2 block
        zc a = at(hi - 1, hi - 1), b = at(hi - 1, hi), c = at(hi, hi - 1), d = at(hi, hi);
        zc tr = a + d, det = a * d - b * c;
        zc disc = std::sqrt(tr * tr - 4.0 * det);
        zc l1 = 0.5 * (tr + disc), l2 = 0.5 * (tr - disc);
        zc mu = (std::abs(l1 - d) < std::abs(l2 - d)) ? l1 : l2;
        ++iter;
        if (iter % 11 == 10) mu = zc(std::abs(at(hi, hi - 1)) + std::abs(at(hi - 1, hi - 2 >= lo ? hi - 2 : lo)), 0.0);  // exceptional shift
        if (iter > 30 * n + 300) break;  // give up (values so far are returned; caller treats NaN as inactive)
        // QR step on the active block [lo, hi]
        std::vector<zc> cs(hi - lo), sn(hi - lo);
        for (int k = lo; k <= hi; ++k) at(k, k) -= mu;
        for (int k = lo; k < hi; ++k) {
            zc x = at(k, k), y = at(k + 1, k);
            double r = std::sqrt(std::norm(x) + std::norm(y));
            zc cc, ss;
            if (r == 0.0) {
                cc = 1.0;
                ss = 0.0;
            } else {
                cc = x / r;
                ss = y / r;
            }
            cs[k - lo] = cc;
            sn[k - lo] = ss;
            // rows k, k+1:  [ conj(c) conj(s); -s c ]
            for (int j = k; j < n; ++j) {
                zc t1 = at(k, j), t2 = at(k + 1, j);
                at(k, j) = std::conj(cc) * t1 + std::conj(ss) * t2;
                at(k + 1, j) = -ss * t1 + cc * t2;
            }
        }
        for (int k = lo; k < hi; ++k) {
            zc cc = cs[k - lo], ss = sn[k - lo];
            const int top = std::min(k + 2, hi);
            for (int i = 0; i <= top; ++i) {
                zc t1 = at(i, k), t2 = at(i, k + 1);
                at(i, k) = t1 * cc + t2 * ss;
                at(i, k + 1) = -t1 * std::conj(ss) + t2 * std::conj(cc);
            }
        }
        for (int k = lo; k <= hi; ++k) at(k, k) += mu;
    }
    return ev;
}

std::vector<zc> elph_debug_hess_eig(const std::vector<double>& h, int n) { return hessenberg_eigvals(h, n); }

// ------------------------------------------------------------------------------------------------
// host: single-slice products with the tau-averaged operator (for Arnoldi only)
// ------------------------------------------------------------------------------------------------
static void host_mulA(const elph_handle* h, const KpmState& K, const std::vector<double>& v, std::vector<double>& out) {
    const int N = h->N;
    for (int i = 0; i < N; ++i) out[i] = K.eVbar[i] * v[i];
    for (int n = 0; n < h->Nb; ++n) {  // checkerboard_mul!, bond 1 first (src/Checkerboard.jl:123-141)
        const int i = h->bonds_host[n].x, j = h->bonds_host[n].y;
        const double c = K.cbar[n], s = K.sbar[n];
        const double t1 = out[i], t2 = out[j];
        out[i] = c * t1 + s * t2;
        out[j] = c * t2 + s * t1;
    }
}

static void host_ldivA(const elph_handle* h, const KpmState& K, const std::vector<double>& v, std::vector<double>& out) {
    const int N = h->N;
    out = v;
    for (int n = h->Nb - 1; n >= 0; --n) {  // checkerboard_inverse_mul! (src/Checkerboard.jl:298-316)
        const int i = h->bonds_host[n].x, j = h->bonds_host[n].y;
        const double c = K.cbar[n], s = K.sbar[n];
        const double t1 = out[i], t2 = out[j];
        out[i] = c * t1 - s * t2;
        out[j] = c * t2 - s * t1;
    }
    for (int i = 0; i < N; ++i) out[i] /= K.eVbar[i];
}

// dot product with four independent accumulators (the strict left-to-right sum does not vectorise; the Arnoldi bounds
// only enter through isapprox(rtol = buf) and floor() of the polynomial orders, far above this rounding difference)
static double host_dot(const double* a, const double* b, int n) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int i = 0;
    for (; i + 3 < n; i += 4) {
        s0 += a[i] * b[i];
        s1 += a[i + 1] * b[i + 1];
        s2 += a[i + 2] * b[i + 2];
        s3 += a[i + 3] * b[i + 3];
    }
    for (; i < n; ++i) s0 += a[i] * b[i];
    return (s0 + s1) + (s2 + s3);
}

// one half of arnoldi_eigenvalue_bounds! (:845-942): returns max real eigenvalue of the projected operator
static double host_arnoldi(const elph_handle* h, const KpmState& K, const double* start, bool inverse) {
    const int N = h->N, n = K.n;
    std::vector<double> Q((size_t)(n + 1) * N, 0.0), hm((size_t)(n + 1) * n, 0.0), b(N), v(N);
    double nb = 0.0;
    for (int i = 0; i < N; ++i) nb += start[i] * start[i];
    nb = std::sqrt(nb);
    for (int i = 0; i < N; ++i) b[i] = start[i] / nb;
    std::copy(b.begin(), b.end(), Q.begin());
    int l = n;
    for (int k = 0; k < n; ++k) {
        if (inverse) host_ldivA(h, K, b, v); else host_mulA(h, K, b, v);
        for (int j = 0; j <= k; ++j) {
            const double* Qj = &Q[(size_t)j * N];
            const double d = host_dot(Qj, v.data(), N);
            hm[(size_t)j * n + k] = d;
            for (int i = 0; i < N; ++i) v[i] -= d * Qj[i];
        }
        const double nv = std::sqrt(host_dot(v.data(), v.data(), N));
        hm[(size_t)(k + 1) * n + k] = nv;
        if (nv > 1e-12) {
            for (int i = 0; i < N; ++i) b[i] = v[i] / nv;
            std::copy(b.begin(), b.end(), Q.begin() + (size_t)(k + 1) * N);
        } else {
            l = k + 1;
            break;
        }
    }
    std::vector<double> hh((size_t)l * l);
    bool finite = true;
    for (int i = 0; i < l; ++i)
        for (int j = 0; j < l; ++j) {
            hh[(size_t)i * l + j] = hm[(size_t)i * n + j];
            if (!std::isfinite(hh[(size_t)i * l + j])) finite = false;
        }
    if (!finite) return INFINITY;
    std::vector<zc> ev = hessenberg_eigvals(hh, l);
    double mx = -INFINITY;
    for (auto& e : ev) mx = std::max(mx, e.real());
    return mx;
}

// kpm_coefficients! (:789-839): c_0 = S_0/(2M), c_m = 2 S_m/(2M), S_m = sum_n f(x_n) cos(pi m (n+1/2)/(2M))
static void host_kpm_coefficients(zc* c, int order, double lam_lo, double lam_hi, double phi) {
    const int M = order, NM = 2 * M;
    const double lam_avg = (lam_hi + lam_lo) / 2, lam_mag = (lam_hi - lam_lo) / 2;
    const double pi = 3.14159265358979323846;
    std::vector<zc> f(NM);
    const zc eph = std::exp(zc(0.0, -phi));
    for (int n = 0; n < NM; ++n) {
        const double xn = lam_mag * std::cos(pi * (n + 0.5) / NM) + lam_avg;
        f[n] = 1.0 / (1.0 - eph * xn);
    }
    for (int m = 0; m < M; ++m) {
        zc S(0.0, 0.0);
        for (int n = 0; n < NM; ++n) S += f[n] * std::cos(pi * m * (n + 0.5) / NM);
        c[m] = (m == 0) ? S / (double)NM : 2.0 * S / (double)NM;
    }
}

static bool isapprox_rtol(double x, double y, double rtol) { return x == y || std::fabs(x - y) <= rtol * std::max(std::fabs(x), std::fabs(y)); }

// ------------------------------------------------------------------------------------------------
// device kernels
// ------------------------------------------------------------------------------------------------
namespace {

// Holstein: eVbar[i] = (sum_tau expnV[tau][i]) / L, summed in tau order (:340-347)
__global__ void taumean_kernel(const double* __restrict__ tab, double* __restrict__ out, int ncols, int L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncols) return;
    double s = 0.0;
    for (int t = 0; t < L; ++t) s += tab[(size_t)t * ncols + i];
    out[i] = s / (double)L;
}
// SSH: (cbar,sbar)[b] = mean_tau (cosh,sinh)[tau][b] (:367-376)
__global__ void taumean2_kernel(const double2* __restrict__ tab, double2* __restrict__ out, int ncols, int L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncols) return;
    double sc = 0.0, ss = 0.0;
    for (int t = 0; t < L; ++t) {
        const double2 v = tab[(size_t)t * ncols + i];
        sc += v.x;
        ss += v.y;
    }
    out[i] = make_double2(sc / (double)L, ss / (double)L);
}

struct KpmParams {
    const cplx* __restrict__ in;   // [L][N] frequency-space input
    cplx* __restrict__ out;        // [L][N]
    const double* __restrict__ eVbar;
    const double2* __restrict__ csbar;
    const int2* __restrict__ bonds;
    const int* __restrict__ goff;
    const cplx* __restrict__ coeff;
    const int* __restrict__ order;
    const int* __restrict__ coeff_off;
    const int* __restrict__ schedule;
    const int* skip;
    int ngroups, N, L;
    double inv_mag, avg_over_mag;
};

__device__ __forceinline__ cplx cmulc(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

template <bool REVERSE>
__device__ __forceinline__ void sweep_cplx(cplx* __restrict__ sw, const KpmParams& P) {
    for (int gg = 0; gg < P.ngroups; ++gg) {
        const int g = REVERSE ? (P.ngroups - 1 - gg) : gg;
        const int lo = P.goff[g], hi = P.goff[g + 1];
        for (int b = lo + threadIdx.x; b < hi; b += blockDim.x) {
            const int2 ij = P.bonds[b];
            const double2 cs = P.csbar[b];
            const cplx t1 = sw[ij.x], t2 = sw[ij.y];
            sw[ij.x] = make_double2(cs.x * t1.x + cs.y * t2.x, cs.x * t1.y + cs.y * t2.y);
            sw[ij.y] = make_double2(cs.x * t2.x + cs.y * t1.x, cs.x * t2.y + cs.y * t1.y);
        }
        __syncthreads();
    }
}

// sum_m c_m T_m(A') v  (TRANSPOSED: A'^T and conjugated coefficients)  -- :625-676
template <int SPT, bool TRANSPOSED>
__device__ __forceinline__ void poly(cplx (&acc)[SPT], const cplx (&vin)[SPT], cplx* __restrict__ sw, const cplx* __restrict__ c,
                                     int order, const KpmParams& P, const double (&ev)[SPT]) {
    cplx uprev[SPT], un[SPT];
    cplx c0 = c[0];
    if (TRANSPOSED) c0.y = -c0.y;
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
        acc[k] = cmulc(c0, vin[k]);
        un[k] = vin[k];
        uprev[k] = make_double2(0.0, 0.0);
    }
    for (int n = 1; n < order; ++n) {
        // sw = A u_n  or  A^T u_n
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
            const int i = threadIdx.x + k * blockDim.x;
            if (i < P.N) sw[i] = TRANSPOSED ? un[k] : make_double2(ev[k] * un[k].x, ev[k] * un[k].y);
        }
        __syncthreads();
        sweep_cplx<TRANSPOSED>(sw, P);
        cplx cn = c[n];
        if (TRANSPOSED) cn.y = -cn.y;
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
            const int i = threadIdx.x + k * blockDim.x;
            cplx t = (i < P.N) ? sw[i] : make_double2(0.0, 0.0);
            if (TRANSPOSED) t = make_double2(ev[k] * t.x, ev[k] * t.y);
            // A' u = (1/mag) A u - (avg/mag) u
            cplx a = make_double2(P.inv_mag * t.x - P.avg_over_mag * un[k].x, P.inv_mag * t.y - P.avg_over_mag * un[k].y);
            if (n > 1) a = make_double2(2.0 * a.x - uprev[k].x, 2.0 * a.y - uprev[k].y);
            uprev[k] = un[k];
            un[k] = a;
            const cplx ca = cmulc(cn, a);
            acc[k].x += ca.x;
            acc[k].y += ca.y;
        }
        // the next iteration's write to sw[i] is by the owning thread; the barrier after it orders it against the sweep
    }
}

template <int SPT>
__global__ void __launch_bounds__(512) kpm_apply_kernel(KpmParams P) {
    extern __shared__ double smem_raw[];
    cplx* sw = reinterpret_cast<cplx*>(smem_raw);
    if (P.skip && *P.skip) return;
    const int w = P.schedule[blockIdx.x];
    const int order = P.order[w];
    const cplx* c = P.coeff + P.coeff_off[w];
    cplx v[SPT], t1[SPT], t2[SPT];
    double ev[SPT];
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
        const int i = threadIdx.x + k * blockDim.x;
        v[k] = (i < P.N) ? P.in[(size_t)w * P.N + i] : make_double2(0.0, 0.0);
        ev[k] = (i < P.N) ? P.eVbar[i] : 0.0;
    }
    poly<SPT, true>(t1, v, sw, c, order, P, ev);   // M^-T[w,w]
    poly<SPT, false>(t2, t1, sw, c, order, P, ev); // M^-1[w,w]
    const int wm = P.L - 1 - w;  // mirror frequency, conj (:464-466; for odd L the middle one overwrites itself)
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
        const int i = threadIdx.x + k * blockDim.x;
        if (i < P.N) {
            if (wm != w) P.out[(size_t)w * P.N + i] = t2[k];
            P.out[(size_t)wm * P.N + i] = make_double2(t2[k].x, -t2[k].y);
        }
    }
}

template <int SPT>
void launch_apply(elph_handle* h, const KpmParams& P, int threads) {
    const size_t smem = (size_t)h->N * sizeof(cplx);
    ELPH_REQUIRE(smem <= h->smem_optin, ELPH_ERR_UNSUPPORTED, "Nsites too large for the KPM shared-memory kernel");
    elph_enable_smem(h, kpm_apply_kernel<SPT>);
    kpm_apply_kernel<SPT><<<h->kpm.Lo2, threads, smem, h->stream>>>(P);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}

}  // namespace

void elph_kpm_init(elph_handle* h, int n, double buf, double c1, double c2) {
    KpmState& K = h->kpm;
    K.configured = true;
    K.n = std::min(n, h->N);
    K.buf = buf;
    K.c1 = c1;
    K.c2 = c2;
    K.Lo2 = (h->L + 1) / 2;
    K.phis.resize(K.Lo2);
    const double pi = 3.14159265358979323846;
    for (int w = 0; w < K.Lo2; ++w) K.phis[w] = 2 * pi / h->L * (w + 0.5);
    K.order.assign(K.Lo2, 1);
    K.coeff_off.resize(K.Lo2);
    std::iota(K.coeff_off.begin(), K.coeff_off.end(), 0);
    K.coeff.assign(K.Lo2, zc(0.0, 0.0));
    K.schedule.resize(K.Lo2);
    std::iota(K.schedule.begin(), K.schedule.end(), 0);
    K.eVbar.assign(h->N, 0.0);
    K.cbar.assign(h->Nb, 0.0);
    K.sbar.assign(h->Nb, 0.0);
    K.d_eVbar = elph_dalloc<double>(h->N);
    K.d_csbar = elph_dalloc<double2>(h->Nb);
    K.d_order = elph_dalloc<int>(K.Lo2);
    K.d_coeff_off = elph_dalloc<int>(K.Lo2);
    K.d_schedule = elph_dalloc<int>(K.Lo2);
    K.d_nu = elph_dalloc<cplx>((size_t)h->L * h->N);
    ELPH_CUDA(cudaMemset(K.d_nu, 0, (size_t)h->L * h->N * sizeof(cplx)));
    ELPH_CUDA(cudaDeviceSynchronize());
    h->kpm_version++;
}

void elph_kpm_free(elph_handle* h) {
    KpmState& K = h->kpm;
    cudaFree(K.d_eVbar);
    cudaFree(K.d_csbar);
    cudaFree(K.d_coeff);
    cudaFree(K.d_order);
    cudaFree(K.d_coeff_off);
    cudaFree(K.d_schedule);
    cudaFree(K.d_nu);
    K = KpmState();
}

void elph_kpm_setup_impl(elph_handle* h, const double* noise, elph_kpm_info* info) {
    KpmState& K = h->kpm;
    ELPH_REQUIRE(K.configured, ELPH_ERR_STATE, "KPM preconditioner not configured (kpm_n == 0 at elph_create)");
    ELPH_REQUIRE(noise != nullptr, ELPH_ERR_INVALID, "arnoldi_noise must provide 2*Nsites values");
    const int N = h->N, L = h->L, Nb = h->Nb, T = 256;
    // update_A!
    if (h->model == ELPH_MODEL_HOLSTEIN) {
        taumean_kernel<<<(N + T - 1) / T, T, 0, h->stream>>>(h->d_D, K.d_eVbar, N, L);
        ELPH_CUDA(cudaGetLastError());
        h->launches++;
        if (!K.ever_setup) {  // static hoppings: cbar = cosht, sbar = sinht (:128-130)
            std::vector<double2> cs(Nb);
            ELPH_CUDA(cudaMemcpyAsync(cs.data(), h->d_cs, Nb * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
            ELPH_CUDA(cudaStreamSynchronize(h->stream));
            for (int b = 0; b < Nb; ++b) { K.cbar[b] = cs[b].x; K.sbar[b] = cs[b].y; }
            ELPH_CUDA(cudaMemcpyAsync(K.d_csbar, h->d_cs, Nb * sizeof(double2), cudaMemcpyDeviceToDevice, h->stream));
        }
        ELPH_CUDA(cudaMemcpyAsync(K.eVbar.data(), K.d_eVbar, N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        ELPH_CUDA(cudaStreamSynchronize(h->stream));
    } else {
        if (Nb > 0) {
            taumean2_kernel<<<(Nb + T - 1) / T, T, 0, h->stream>>>(h->d_cs, K.d_csbar, Nb, L);
            ELPH_CUDA(cudaGetLastError());
            h->launches++;
        }
        ELPH_CUDA(cudaMemcpyAsync(K.d_eVbar, h->d_D, N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        std::vector<double2> cs(Nb);
        ELPH_CUDA(cudaMemcpyAsync(cs.data(), K.d_csbar, Nb * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
        ELPH_CUDA(cudaMemcpyAsync(K.eVbar.data(), K.d_eVbar, N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        ELPH_CUDA(cudaStreamSynchronize(h->stream));
        for (int b = 0; b < Nb; ++b) { K.cbar[b] = cs[b].x; K.sbar[b] = cs[b].y; }
    }
    // Arnoldi bounds
    // the two Krylov runs (on A for e_max, on A^-1 for e_min) are independent: second host thread for the inverse one
    auto inv_run = std::async(std::launch::async, [&]() { return host_arnoldi(h, K, noise + N, true); });
    const double e_max = host_arnoldi(h, K, noise, false);
    const double inv_max = inv_run.get();
    const double e_min = std::isfinite(inv_max) ? 1.0 / inv_max : -INFINITY;
    K.e_min = e_min;
    K.e_max = e_max;
    bool recomputed = false;
    if ((0.0 < e_min && e_min < 1.0) && (1.0 < e_max) && (e_max - e_min) < 2.0) {
        const double lam_lo = std::max(0.0, (1 - 2 * K.buf) * e_min);
        const double lam_hi = (1 + 2 * K.buf) * e_max;
        if (!isapprox_rtol(lam_lo, K.lam_lo, K.buf) || !isapprox_rtol(lam_hi, K.lam_hi, K.buf)) {
            K.lam_lo = lam_lo;
            K.lam_hi = lam_hi;
            K.lam_avg = (lam_hi + lam_lo) / 2;
            K.lam_mag = (lam_hi - lam_lo) / 2;
            int total = 0;
            for (int w = 0; w < K.Lo2; ++w) {
                int order = (int)std::floor((lam_hi - lam_lo) * (K.c1 / K.phis[w] + K.c2));
                order = std::max(1, order);
                K.order[w] = order;
                K.coeff_off[w] = total;
                total += order;
            }
            K.coeff.assign(total, zc(0.0, 0.0));
            for (int w = 0; w < K.Lo2; ++w) host_kpm_coefficients(&K.coeff[K.coeff_off[w]], K.order[w], lam_lo, lam_hi, K.phis[w]);
            std::iota(K.schedule.begin(), K.schedule.end(), 0);
            std::stable_sort(K.schedule.begin(), K.schedule.end(), [&](int a, int b) { return K.order[a] > K.order[b]; });
            if ((size_t)total > K.d_coeff_cap) {
                cudaFree(K.d_coeff);
                K.d_coeff_cap = (size_t)total * 2;
                K.d_coeff = elph_dalloc<cplx>(K.d_coeff_cap);
            }
            static_assert(sizeof(zc) == sizeof(cplx), "complex layout");
            ELPH_CUDA(cudaMemcpyAsync(K.d_coeff, K.coeff.data(), total * sizeof(cplx), cudaMemcpyHostToDevice, h->stream));
            ELPH_CUDA(cudaMemcpyAsync(K.d_order, K.order.data(), K.Lo2 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
            ELPH_CUDA(cudaMemcpyAsync(K.d_coeff_off, K.coeff_off.data(), K.Lo2 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
            ELPH_CUDA(cudaMemcpyAsync(K.d_schedule, K.schedule.data(), K.Lo2 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
            ELPH_CUDA(cudaStreamSynchronize(h->stream));  // host vectors may be reallocated by the next setup
            recomputed = true;
            h->kpm_version++;   // captured CG graphs carry the old polynomial orders / window
        }
        K.active = true;
    } else {
        K.active = false;
    }
    K.ever_setup = true;
    if (info) {
        info->active = K.active ? 1 : 0;
        info->recomputed = recomputed ? 1 : 0;
        info->e_min = e_min;
        info->e_max = e_max;
        info->lambda_lo = K.lam_lo;
        info->lambda_hi = K.lam_hi;
        int64_t tot = 0, mx = 0;
        for (int w = 0; w < K.Lo2; ++w) { tot += K.order[w]; mx = std::max<int64_t>(mx, K.order[w]); }
        info->total_order = tot;
        info->max_order = mx;
    }
}

void elph_kpm_apply_dev(elph_handle* h, const double* vin, double* vout) { elph_kpm_apply_dev_cg(h, vin, vout, nullptr); }

// cgf != nullptr (inside the preconditioned CG loop, preconditioner active): the forward FFT kernel first applies
// x += alpha p, r -= alpha Ap and the stop rule, and the inverse FFT kernel accumulates r.z -> beta (fft.cu);
// vin must then be the residual vector r.
void elph_kpm_apply_dev_cg(elph_handle* h, const double* vin, double* vout, const KpmCgFuse* cgf) {
    KpmState& K = h->kpm;
    ELPH_REQUIRE(K.configured && K.ever_setup, ELPH_ERR_STATE, "elph_kpm_apply before elph_kpm_setup");
    ELPH_REQUIRE(!cgf || K.active, ELPH_ERR_STATE, "fused KPM apply needs an active preconditioner");
    if (!K.active) {  // identity (:475-478)
        if (vout != vin) ELPH_CUDA(cudaMemcpyAsync(vout, vin, h->Ndim * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        return;
    }
    ELPH_REQUIRE(K.d_coeff != nullptr, ELPH_ERR_STATE, "KPM coefficients missing");
    const int* skip = &h->d_cg->done;
    if (!h->kpm_skip_enabled) skip = nullptr;
    // the recurrence reads nu_in and writes nu_out (the reference's v1 / v2)
    cplx* nu_in = K.d_nu;
    cplx* nu_out = h->d_nu2;
    auto inverse_fft = [&]() {
        if (cgf) elph_omega_to_tau_dev_cg(h, nu_out, vout, cgf->r);
        else elph_omega_to_tau_dev_skip(h, nu_out, vout, skip);
    };
    if (cgf) elph_tau_to_omega_dev_cg(h, cgf->x, cgf->r, cgf->p, cgf->ap, nu_in);
    else elph_tau_to_omega_dev_skip(h, vin, nu_in, skip);
    KpmParams P;
    P.in = nu_in;
    P.out = nu_out;
    P.eVbar = K.d_eVbar;
    P.csbar = K.d_csbar;
    P.bonds = h->d_bonds;
    P.goff = h->d_goff;
    P.coeff = K.d_coeff;
    P.order = K.d_order;
    P.coeff_off = K.d_coeff_off;
    P.schedule = K.d_schedule;
    P.skip = skip;
    P.ngroups = h->ngroups;
    P.N = h->N;
    P.L = h->L;
    P.inv_mag = 1.0 / K.lam_mag;
    P.avg_over_mag = K.lam_avg / K.lam_mag;
    if (elph_launch_kpm_square(h, nu_in, nu_out, skip)) {   // register/shuffle kernel (kpm_square.cu)
        inverse_fft();
        return;
    }
    int threads = 256;
    while (threads < 512 && threads * 4 < h->N) threads *= 2;   // <= 512 threads: __launch_bounds__(512) on the kernel
    const int spt = (h->N + threads - 1) / threads;
    if (spt <= 1) launch_apply<1>(h, P, threads);
    else if (spt <= 2) launch_apply<2>(h, P, threads);
    else if (spt <= 4) launch_apply<4>(h, P, threads);
    else if (spt <= 8) launch_apply<8>(h, P, threads);
    else if (spt <= 16) launch_apply<16>(h, P, threads);
    else ELPH_REQUIRE(false, ELPH_ERR_UNSUPPORTED, "Nsites too large for the KPM kernel");
    inverse_fft();
}
