// Hybrid Monte Carlo pieces and the whole-trajectory update, device resident.
//
// Reference: src/HMC.jl  update! :310-335, standard_update! :343-473, multitimestep_update! :479-638,
// refresh_v! :648-660, refresh_phi! :666-692, calc_H/K/S/Sf :698-783, calc_dSdx! :749-759, calc_dSfdx! :790-814,
// calc_O^-1 Lambda phi! :820-915, Lambda operators (Holstein; no-ops for SSH) :921-1030.
// Noise (R_v, R+-, Arnoldi start values) and the Metropolis uniform are injected; the accept/reject decision is
// taken on the host exactly like the reference.
#include "elph_internal.cuh"

#include <cmath>

void elph_lincomb(elph_handle* h, double* out, double a, const double* X, double b, const double* Y, double c, const double* Z,
                  int64_t n);
void elph_gather_primary(elph_handle* h, double* out, const double* in);

namespace {

constexpr int kT = 256;

// Lambda[tau][i] = exp(-dtau (lam x + lam2 x^2)/2)        :921-938
__global__ void lambda_kernel(const double* __restrict__ x, const double* __restrict__ lam, const double* __restrict__ lam2,
                              double* __restrict__ Lam, int N, long long n, double dtau) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % N);
        const double xv = x[idx];
        Lam[idx] = exp(-dtau * (lam[i] * xv + lam2[i] * xv * xv) / 2);
    }
}

// mode 0: out(tau) = -Lam(tau+1) v(tau+1), out(L-1) = +Lam(0) v(0)          mulLambda!      :948-962
// mode 1: out(tau) = -v(tau-1)/Lam(tau),   out(0)   = +v(L-1)/Lam(0)        mulLambda^-1!   :975-989
__global__ void lam_apply_kernel(double* __restrict__ out, const double* __restrict__ v, const double* __restrict__ Lam, int N,
                                 int L, int mode) {
    const long long n = (long long)N * L;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % N);
        const int tau = (int)(idx / N);
        if (mode == 0) {
            if (tau < L - 1) out[idx] = -Lam[idx + N] * v[idx + N];
            else out[idx] = Lam[i] * v[i];
        } else {
            if (tau >= 1) out[idx] = -(1.0 / Lam[idx]) * v[idx - N];
            else out[idx] = (1.0 / Lam[i]) * v[(size_t)(L - 1) * N + i];
        }
    }
}

// dS[tau][i] += vl * (+-dtau (lam/2 + lam2 x)) Lam * vr(tau-1)        muldLambdadx!   :1005-1025
__global__ void dlam_kernel(double* __restrict__ dS, const double* __restrict__ vl, const double* __restrict__ vr,
                            const double* __restrict__ x, const double* __restrict__ Lam, const double* __restrict__ lam,
                            const double* __restrict__ lam2, int N, int L, double dtau) {
    const long long n = (long long)N * L;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % N);
        const int tau = (int)(idx / N);
        const double f = dtau * (lam[i] / 2 + lam2[i] * x[idx]);
        const double vrm = (tau == 0) ? vr[(size_t)(L - 1) * N + i] : vr[idx - N];
        dS[idx] += vl[idx] * ((tau == 0) ? -f : f) * Lam[idx] * vrm;
    }
}

// partial sums of a.b restricted to primary fields (SSH kinetic energy, :720-735)
__global__ void __launch_bounds__(kT) primary_dot_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                         const int* __restrict__ primary_ph, int Nph, long long n,
                                                         double* __restrict__ partial) {
    __shared__ double red[32];
    double s = 0.0;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int ph = (int)(idx % Nph);
        if (primary_ph[ph] == ph) s += a[idx] * b[idx];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
        partial[blockIdx.x] = t;
    }
}

int nblocks(elph_handle* h, int64_t n) { return (int)std::min<int64_t>((n + kT - 1) / kT, 8LL * h->sm_count); }

double host_dot(elph_handle* h, const double* a, const double* b, int64_t n) {
    elph_dot_async(h, a, b, n, h->d_scal);
    ELPH_CUDA(cudaMemcpyAsync(h->h_scal, h->d_scal, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    ELPH_CUDA(cudaStreamSynchronize(h->stream));
    return h->h_scal[0];
}

}  // namespace

void elph_hmc_ensure(elph_handle* h) {
    HmcState& S = h->hmc;
    if (S.init) return;
    auto z = [&](int64_t n) {
        double* p = elph_dalloc<double>(n);
        ELPH_CUDA(cudaMemset(p, 0, (n ? n : 1) * sizeof(double)));
        return p;
    };
    S.v = z(h->Ndof); S.v0 = z(h->Ndof); S.x0 = z(h->Ndof); S.dS = z(h->Ndof); S.y = z(h->Ndof); S.Q = z(h->Ndof);
    S.Lam = z(h->Ndim); S.Rp = z(h->Ndim); S.Rm = z(h->Ndim); S.phip = z(h->Ndim); S.phim = z(h->Ndim);
    S.Lphip = z(h->Ndim); S.Lphim = z(h->Ndim); S.Op = z(h->Ndim); S.Om = z(h->Ndim); S.u = z(h->Ndim);
    ELPH_CUDA(cudaDeviceSynchronize());
    S.init = true;
}

void elph_hmc_free(elph_handle* h) {
    HmcState& S = h->hmc;
    if (!S.init) return;
    double* ptrs[] = {S.v, S.v0, S.x0, S.dS, S.y, S.Q, S.Lam, S.Rp, S.Rm, S.phip, S.phim, S.Lphip, S.Lphim, S.Op, S.Om, S.u};
    for (double* p : ptrs) cudaFree(p);
    S = HmcState();
}

static void update_Lam(elph_handle* h) {
    if (h->model != ELPH_MODEL_HOLSTEIN) return;
    lambda_kernel<<<nblocks(h, h->Ndim), kT, 0, h->stream>>>(h->d_x, h->d_lam, h->d_lam2, h->hmc.Lam, h->N, h->Ndim, h->dtau);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}
static void lam_apply(elph_handle* h, double* out, const double* v, int mode) {
    if (h->model != ELPH_MODEL_HOLSTEIN) return;  // no-ops for SSH: `out` keeps its content (:964-966,991-993)
    lam_apply_kernel<<<nblocks(h, h->Ndim), kT, 0, h->stream>>>(out, v, h->hmc.Lam, h->N, h->L, mode);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}

// refresh_v!: v = alpha v + sqrt(1-alpha^2) sqrt(M^-1) R          :648-660
void elph_hmc_refresh_v_dev(elph_handle* h, double alpha, const double* R_dev) {
    HmcState& S = h->hmc;
    if (h->model == ELPH_MODEL_SSH) elph_gather_primary(h, S.y, R_dev);
    else ELPH_CUDA(cudaMemcpyAsync(S.y, R_dev, h->Ndof * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    elph_fourier_accelerate_dev(h, S.y, S.y, -0.5, true);
    elph_lincomb(h, S.v, alpha, S.v, std::sqrt(1.0 - alpha * alpha), S.y, 0.0, nullptr, h->Ndof);
}

// refresh_phi!: Lphi = M^T R ; phi = Lambda^-1 Lphi ; S = (R+^2 + R-^2)/2 + Sb      :666-692
double elph_hmc_refresh_phi_dev(elph_handle* h) {
    HmcState& S = h->hmc;
    update_Lam(h);
    MatvecArgs m;
    m.v = S.Rp; m.y = S.Lphip;
    elph_launch_matvec(h, MODE_MT, m);
    lam_apply(h, S.phip, S.Lphip, 1);
    m.v = S.Rm; m.y = S.Lphim;
    elph_launch_matvec(h, MODE_MT, m);
    lam_apply(h, S.phim, S.Lphim, 1);
    double act = host_dot(h, S.Rp, S.Rp, h->Ndim) / 2 + host_dot(h, S.Rm, S.Rm, h->Ndim) / 2;
    double sb = 0.0;
    elph_Sb_dev(h, false, &sb);
    return act + sb;
}

// calc_O^-1 Lambda phi!  (:820-915)
void elph_hmc_calc_Oinv_dev(elph_handle* h, bool use_precond, const double* arnoldi_host, double power, int64_t* iters, int* flag) {
    HmcState& S = h->hmc;
    // setup!(P) :829; speculative as in the Langevin force (dynamics.cu): the two solves start with the previous polynomials while
    // the Arnoldi bounds are computed beside them, and are repeated if the set-up changes the polynomials
    bool speculative = false;
    if (use_precond && h->kpm.configured) {
        if (elph_kpm_can_speculate(h)) {
            elph_kpm_setup_begin(h, arnoldi_host);
            speculative = true;
            h->spec_running = true;
        } else {
            elph_kpm_setup_impl(h, arnoldi_host, nullptr);
        }
    }
    update_Lam(h);
    lam_apply(h, S.Lphip, S.phip, 0);
    lam_apply(h, S.Lphim, S.phim, 0);
    if (!(use_precond && h->kpm.configured) && h->use_persistent && !h->sharded) {
        // both flavours in one batched launch of the persistent CG (same matrix, independent right-hand sides).  The
        // reference solves "-" only if "+" succeeded and otherwise leaves O^-1 Lambda phi_- untouched: "-" goes to scratch.
        const double* bl[2] = {S.Lphip, S.Lphim};
        double* xl[2] = {S.Op, h->d_vc};
        elph_solve_info inf[2] = {};
        elph_solve_batch_device(h, 2, bl, xl, false, power, inf);
        int64_t tot2 = inf[0].iters;
        int fl2 = inf[0].flag;
        if (fl2 == 0) {
            ELPH_CUDA(cudaMemcpyAsync(S.Om, h->d_vc, h->Ndim * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
            tot2 += inf[1].iters;
            fl2 = inf[1].flag;
        }
        if (fl2 == 0) tot2 = (tot2 + 1) / 2;  // cld(iters, 2)
        *iters = tot2;
        *flag = fl2;
        return;
    }
    // a rejected speculation must leave O^-1 Lambda phi_- as the reference would: keep what it held before the first attempt
    if (speculative)
        ELPH_CUDA(cudaMemcpyAsync(h->d_vc, S.Om, h->Ndim * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    auto solve_pair = [&]() {
        elph_solve_info info = {};
        int64_t tot = 0;
        ELPH_CUDA(cudaMemsetAsync(S.Op, 0, h->Ndim * sizeof(double), h->stream));
        elph_solve_device(h, S.Lphip, S.Op, use_precond, power, &info);
        tot += info.iters;
        int fl = info.flag;
        if (fl == 0) {
            ELPH_CUDA(cudaMemsetAsync(S.Om, 0, h->Ndim * sizeof(double), h->stream));
            elph_solve_device(h, S.Lphim, S.Om, use_precond, power, &info);
            tot += info.iters;
            fl = info.flag;
        }
        if (fl == 0) tot = (tot + 1) / 2;  // cld(iters, 2)
        *iters = tot;
        *flag = fl;
    };
    if (speculative) {
        try {
            solve_pair();
        } catch (...) {
            h->spec_running = false;
            try { elph_kpm_setup_finish(h, nullptr); } catch (...) {}
            throw;
        }
        h->spec_running = false;
        if (elph_kpm_setup_finish(h, nullptr)) {
            ELPH_CUDA(cudaMemcpyAsync(S.Om, h->d_vc, h->Ndim * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
            solve_pair();
        }
    } else {
        solve_pair();
    }
}

double elph_hmc_calc_Sf_dev(elph_handle* h);

// field move of the special updates on the device copy of x ([tau][phonon]): kind 0 negates column i (reflection,
// src/SpecialUpdates.jl:129), kind 1 exchanges columns i and j (swap, :269 / :335)
__global__ void field_move_kernel(double* __restrict__ x, int L, int Nph, int kind, int i, int j) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < L; t += gridDim.x * blockDim.x) {
        double* row = x + (size_t)t * Nph;
        if (kind == 0) {
            row[i] = -row[i];
        } else {
            const double a = row[i];
            row[i] = row[j];
            row[j] = a;
        }
    }
}

// One proposal of special_update! (src/SpecialUpdates.jl:97-160 reflection, :233-290 and :296-366 swap): S0 from a
// phi refresh with the injected R+- (already in S.Rp, S.Rm), the field move, update_model!, the two solves at tol^2,
// S1 = Sf + Sb, Metropolis test with the injected uniform; the move is undone on rejection (or solver failure).
void elph_hmc_special_update_dev(elph_handle* h, int kind, int i, int j, bool use_precond, const double* arnoldi_host, double uniform,
                                 int* accepted, double* S0_out, double* S1_out, int64_t* iters, int* flag) {
    auto move = [&]() {
        field_move_kernel<<<(h->L + 127) / 128, 128, 0, h->stream>>>(h->d_x, h->L, h->Nph, kind, i, j);
        ELPH_CUDA(cudaGetLastError());
        h->launches++;
        elph_launch_update_model(h);
    };
    elph_launch_update_model(h);
    const double S0 = elph_hmc_refresh_phi_dev(h);
    move();
    int64_t it = 0;
    int fl = 0;
    elph_hmc_calc_Oinv_dev(h, use_precond, arnoldi_host, 2.0, &it, &fl);
    double sb = 0.0;
    elph_Sb_dev(h, false, &sb);
    const double S1 = elph_hmc_calc_Sf_dev(h) + sb;
    const double P = std::min(1.0, std::exp(-(S1 - S0)));
    const bool ok = (uniform < P) && fl == 0;
    if (!ok) move();
    *accepted = ok ? 1 : 0;
    *S0_out = S0;
    *S1_out = S1;
    *iters = it;
    *flag = fl;
}

double elph_hmc_calc_Sf_dev(elph_handle* h) {
    HmcState& S = h->hmc;
    return host_dot(h, S.Lphip, S.Op, h->Ndim) / 2 + host_dot(h, S.Lphim, S.Om, h->Ndim) / 2;
}

double elph_hmc_calc_K_dev(elph_handle* h) {
    HmcState& S = h->hmc;
    elph_fourier_accelerate_dev(h, S.v, S.y, 1.0, true);
    if (h->model == ELPH_MODEL_HOLSTEIN) return host_dot(h, S.v, S.y, h->Ndof) / 2;
    const int blocks = std::min(h->partial_cap, 2 * h->sm_count);
    primary_dot_kernel<<<blocks, kT, 0, h->stream>>>(S.v, S.y, h->d_primary_ph, h->Nph, h->Ndof, h->d_partial);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
    std::vector<double> part(blocks);
    ELPH_CUDA(cudaMemcpyAsync(part.data(), h->d_partial, blocks * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    ELPH_CUDA(cudaStreamSynchronize(h->stream));
    double s = 0.0;
    for (double p : part) s += p;
    return s / 2;
}

void elph_hmc_calc_H_dev(elph_handle* h, double* H, double* Sout, double* K) {
    double sb = 0.0;
    elph_Sb_dev(h, false, &sb);
    const double s = elph_hmc_calc_Sf_dev(h) + sb;
    const double k = elph_hmc_calc_K_dev(h);
    *H = s + k;
    *Sout = s;
    *K = k;
}

// dS += fermionic force (:790-814); dS must be zeroed by the caller like the reference's fill!(dSdx, 0)
void elph_hmc_calc_dSfdx_dev(elph_handle* h, double* dS) {
    HmcState& S = h->hmc;
    const double* O[2] = {S.Op, S.Om};
    const double* phi[2] = {S.phip, S.phim};
    for (int s = 0; s < 2; ++s) {
        MatvecArgs m;
        m.v = O[s]; m.y = S.u;
        elph_launch_matvec(h, MODE_M, m);                                   // u = M O^-1 Lambda phi
        elph_muldMdx_dev(h, S.u, O[s], h->d_dSdx2, 1.0, false, false);      // <dM/dx>
        elph_lincomb(h, dS, 1.0, dS, -1.0, h->d_dSdx2, 0.0, nullptr, h->Ndof);
    }
    if (h->model == ELPH_MODEL_HOLSTEIN) {
        for (int s = 0; s < 2; ++s) {
            dlam_kernel<<<nblocks(h, h->Ndim), kT, 0, h->stream>>>(dS, phi[s], O[s], h->d_x, S.Lam, h->d_lam, h->d_lam2, h->N, h->L,
                                                                  h->dtau);
            ELPH_CUDA(cudaGetLastError());
            h->launches++;
        }
    }
}

// update! (:310-335): whole trajectory.  arnoldi: (Nt+2) blocks of 2N values in call order, or NULL.
void elph_hmc_update_dev(elph_handle* h, double dt, int Nt, int Nb, double alpha, const double* Rv_dev, bool use_precond,
                         const double* arnoldi_host, double uniform, int32_t* accepted, double* iters_out, double* H0out,
                         double* H1out, int32_t* flag_out) {
    HmcState& S = h->hmc;
    const int64_t nd = h->Ndof;
    const double dtp = dt / Nb;
    int noise_idx = 0;
    auto next_noise = [&]() -> const double* {
        const double* p = arnoldi_host ? arnoldi_host + (size_t)noise_idx * 2 * h->N : nullptr;
        ++noise_idx;
        return p;
    };
    int64_t iters = 0, it = 0;
    int flag = 0;
    elph_launch_update_model(h);
    elph_hmc_refresh_v_dev(h, alpha, Rv_dev);
    ELPH_CUDA(cudaMemcpyAsync(S.x0, h->d_x, nd * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    ELPH_CUDA(cudaMemcpyAsync(S.v0, S.v, nd * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    elph_hmc_refresh_phi_dev(h);
    elph_hmc_calc_Oinv_dev(h, use_precond, next_noise(), 2.0, &it, &flag);
    if (Nb == 1) iters = it;  // multitimestep drops the first count: `iters += iters` (:515)
    double H0 = NAN, H1 = NAN, Sx, Kx;
    auto force = [&](bool with_boson) {
        ELPH_CUDA(cudaMemsetAsync(S.dS, 0, nd * sizeof(double), h->stream));
        elph_hmc_calc_dSfdx_dev(h, S.dS);
        if (with_boson) elph_dSbdx_dev(h, S.dS, false);
        elph_fourier_accelerate_dev(h, S.dS, S.Q, -1.0, true);
    };
    auto boson_force = [&]() {
        ELPH_CUDA(cudaMemsetAsync(S.dS, 0, nd * sizeof(double), h->stream));
        elph_dSbdx_dev(h, S.dS, false);
        elph_fourier_accelerate_dev(h, S.dS, S.y, -1.0, true);
    };
    elph_trace_mark(h, "hmc: refresh + first solves");
    if (flag == 0) {
        elph_hmc_calc_H_dev(h, &H0, &Sx, &Kx);
        force(Nb == 1);
        elph_trace_mark(h, "hmc: H0 + first force");
        for (int t = 0; t < Nt; ++t) {
            elph_lincomb(h, S.v, 1.0, S.v, -dt / 2, S.Q, 0.0, nullptr, nd);
            if (Nb == 1) {
                elph_lincomb(h, h->d_x, 1.0, h->d_x, dt, S.v, 0.0, nullptr, nd);
            } else if (h->hmc_fused_inner && elph_hmc_inner_dev(h, h->d_x, S.v, dtp, Nb)) {
                // the whole inner loop ran in one launch (same operations as the step-by-step kernels below)
            } else {
                boson_force();
                for (int tp = 0; tp < Nb; ++tp) {
                    elph_lincomb(h, S.v, 1.0, S.v, -dtp / 2, S.y, 0.0, nullptr, nd);
                    elph_lincomb(h, h->d_x, 1.0, h->d_x, dtp, S.v, 0.0, nullptr, nd);
                    boson_force();
                    elph_lincomb(h, S.v, 1.0, S.v, -dtp / 2, S.y, 0.0, nullptr, nd);
                }
            }
            elph_trace_mark(h, "hmc: kick + inner steps");
            elph_launch_update_model(h);
            elph_hmc_calc_Oinv_dev(h, use_precond, next_noise(), 1.0, &it, &flag);
            iters += it;
            elph_trace_mark(h, "hmc: update_model + solves");
            if (flag > 0) break;
            force(Nb == 1);
            elph_lincomb(h, S.v, 1.0, S.v, -dt / 2, S.Q, 0.0, nullptr, nd);
            elph_trace_mark(h, "hmc: force + kick");
        }
    }
    double P = 0.0;
    if (flag == 0) {
        elph_hmc_calc_Oinv_dev(h, use_precond, next_noise(), 2.0, &it, &flag);
        iters += it;
        if (flag == 0) {
            elph_hmc_calc_H_dev(h, &H1, &Sx, &Kx);
            P = std::min(1.0, std::exp(-(H1 - H0)));
        }
    }
    const bool acc = (uniform < P) && flag == 0;
    if (!acc) {
        ELPH_CUDA(cudaMemcpyAsync(h->d_x, S.x0, nd * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        elph_lincomb(h, S.v, -1.0, S.v0, 0.0, nullptr, 0.0, nullptr, nd);
        elph_launch_update_model(h);
    }
    ELPH_CUDA(cudaStreamSynchronize(h->stream));
    if (accepted) *accepted = acc ? 1 : 0;
    if (iters_out) *iters_out = (double)((iters + (Nt + 2) - 1) / (Nt + 2));  // cld(iters, Nt+2)
    if (H0out) *H0out = H0;
    if (H1out) *H1out = H1;
    if (flag_out) *flag_out = flag;
}
