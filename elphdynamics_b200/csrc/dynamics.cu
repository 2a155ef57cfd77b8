// K7: Langevin dynamics updates built from the device-resident pieces.
//
// Reference: src/LangevinDynamics.jl  evolve! Euler :81-119, Runge-Kutta :162-225, Heun :272-324,
// calc_dSdx! :334-345, calc_dSfdx! :350-384.  Noise (eta, g) and the Arnoldi start values are injected.
#include "elph_internal.cuh"

#include <chrono>
#include <cmath>
#include <cstdio>

namespace {

constexpr int kT = 256;

// out = a*X + b*Y + c*Z   (Y, Z optional).  out may alias an input (x = x + dx, v = v - dt/2 Q, ...): no __restrict__, every
// thread reads its own index before it writes it.
__global__ void lincomb_kernel(double* out, double a, const double* X, double b, const double* Y, double c, const double* Z,
                               long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double r = a * X[i];
        if (Y) r += b * Y[i];
        if (Z) r += c * Z[i];
        out[i] = r;
    }
}

// eta[tau][ph] = eta_in[tau][primary(ph)]   (randn!(eta, ssh): v = v[primary_field], src/SSHModels.jl:567-576)
__global__ void gather_primary_kernel(double* __restrict__ out, const double* __restrict__ in, const int* __restrict__ primary_ph,
                                      int Nph, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int ph = (int)(i % Nph);
        out[i] = in[i - ph + primary_ph[ph]];
    }
}

}  // namespace

void elph_trace_mark(elph_handle* h, const char* label) {
    if (!h->trace) return;
    cudaStreamSynchronize(h->stream);
    const double now = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    if (label) fprintf(stderr, "[elph trace] %-28s %9.1f us\n", label, (now - h->trace_t0) * 1e6);
    h->trace_t0 = now;
}

void elph_lincomb(elph_handle* h, double* out, double a, const double* X, double b, const double* Y, double c, const double* Z,
                  int64_t n) {
    const int blocks = (int)std::min<int64_t>((n + kT - 1) / kT, 8LL * h->sm_count);
    lincomb_kernel<<<blocks, kT, 0, h->stream>>>(out, a, X, b, Y, c, Z, n);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}

void elph_gather_primary(elph_handle* h, double* out, const double* in) {
    const int64_t n = h->Ndof;
    const int blocks = (int)std::min<int64_t>((n + kT - 1) / kT, 8LL * h->sm_count);
    gather_primary_kernel<<<blocks, kT, 0, h->stream>>>(out, in, h->d_primary_ph, h->Nph, n);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}

// calc_dSdx!(dSdx, g, M^-1 g, model, P): src/LangevinDynamics.jl:334-384 with g injected
void elph_calc_dSdx_dev(elph_handle* h, const double* g_dev, const double* arnoldi_host, bool use_precond, double* dSdx_dev,
                        double* Minv_dev, elph_solve_info* info) {
    elph_trace_mark(h, "(before calc_dSdx)");
    // setup!(P) :364.  The polynomials survive a set-up unless the spectral window moved by more than the hysteresis buffer, so the
    // solve starts right behind update_A! with the polynomials it would almost always end up with, while the Arnoldi kernel runs
    // beside it (two SMs are left free) and a host thread reduces its result; if the set-up does change them, the solve is repeated.
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
    elph_trace_mark(h, "kpm setup");
    ELPH_CUDA(cudaMemsetAsync(Minv_dev, 0, h->Ndim * sizeof(double), h->stream));       // fill!(M^-1 g, 0) :365
    MatvecArgs m;
    m.v = g_dev;
    m.y = h->d_b;
    elph_launch_matvec(h, MODE_MT, m);                                                    // b = M^T g :373
    elph_trace_mark(h, "b = M^T g");
    if (speculative) {
        bool stale = true;
        try {
            elph_solve_device(h, h->d_b, Minv_dev, use_precond, 1.0, info);
        } catch (...) {
            h->spec_running = false;
            try { elph_kpm_setup_finish(h, nullptr); } catch (...) {}
            throw;
        }
        h->spec_running = false;
        stale = elph_kpm_setup_finish(h, nullptr);
        if (stale) {   // the window moved: repeat with the new polynomials (or without, if the preconditioner switched itself off)
            ELPH_CUDA(cudaMemsetAsync(Minv_dev, 0, h->Ndim * sizeof(double), h->stream));
            elph_solve_device(h, h->d_b, Minv_dev, use_precond, 1.0, info);
        }
    } else {
        elph_solve_device(h, h->d_b, Minv_dev, use_precond, 1.0, info);                   // ldiv! :374
    }
    elph_trace_mark(h, "solve");
    // dSdx = -2 <dM/dx> + dSb/dx (shifted = true)   :378-381, :341
    elph_muldMdx_dev(h, g_dev, Minv_dev, dSdx_dev, -2.0, true, true);
    elph_trace_mark(h, "force + dSb/dx");
}

void elph_langevin_step_dev(elph_handle* h, int method, double dt, const double* eta_dev, const double* g1_dev,
                            const double* g2_dev, const double* arn1, const double* arn2, bool use_precond, int64_t* iters,
                            elph_solve_info* info1, elph_solve_info* info2) {
    const int64_t nd = h->Ndof;
    const double s2dt = std::sqrt(2.0 * dt);
    double* x = h->d_x;
    double* eta = h->d_eta;
    // eta (and g2) may still be on their way on the upload stream (elph_langevin_step): wait right before the first use
    auto fetch_eta = [&]() {
        if (h->upload_pending) {
            ELPH_CUDA(cudaStreamWaitEvent(h->stream, h->upload_event, 0));
            h->upload_pending = false;
        }
        if (h->model == ELPH_MODEL_SSH) {
            elph_gather_primary(h, eta, eta_dev);
        } else {
            ELPH_CUDA(cudaMemcpyAsync(eta, eta_dev, nd * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        }
    };
    if (method == ELPH_LANGEVIN_HEUN) fetch_eta();
    elph_solve_info i1 = {}, i2 = {};
    if (method == ELPH_LANGEVIN_EULER) {
        elph_launch_update_model(h);                                                       // :91
        elph_calc_dSdx_dev(h, g1_dev, arn1, use_precond, h->d_dSdx, h->d_Minv, &i1);       // :101
        fetch_eta();
        elph_fourier_accelerate_dev(h, h->d_dSdx, h->d_dSdx, 1.0, false);                  // :104
        elph_fourier_accelerate_dev(h, eta, eta, 0.5, false);                              // :107
        elph_lincomb(h, h->d_dx, s2dt, eta, -dt, h->d_dSdx, 0.0, nullptr, nd);             // :110
        elph_lincomb(h, x, 1.0, x, 1.0, h->d_dx, 0.0, nullptr, nd);                        // :113
        elph_launch_update_model(h);                                                       // :116
        if (iters) *iters = i1.iters;
    } else if (method == ELPH_LANGEVIN_RK) {
        elph_launch_update_model(h);                                                       // :178
        elph_calc_dSdx_dev(h, g1_dev, arn1, use_precond, h->d_dSdx, h->d_Minv, &i1);       // :185
        fetch_eta();
        elph_lincomb(h, h->d_dx, s2dt, eta, -dt, h->d_dSdx, 0.0, nullptr, nd);             // :188
        elph_lincomb(h, x, 1.0, x, 1.0, h->d_dx, 0.0, nullptr, nd);                        // :191
        elph_launch_update_model(h);                                                       // :194
        elph_calc_dSdx_dev(h, g2_dev, arn2, use_precond, h->d_dSdx2, h->d_Minv, &i2);      // :198
        elph_lincomb(h, x, 1.0, x, -1.0, h->d_dx, 0.0, nullptr, nd);                       // :201
        elph_launch_update_model(h);                                                       // :204
        elph_lincomb(h, h->d_dSdx, 0.5, h->d_dSdx2, 0.5, h->d_dSdx, 0.0, nullptr, nd);     // :207
        elph_fourier_accelerate_dev(h, h->d_dSdx, h->d_dSdx, 1.0, false);                  // :210
        elph_fourier_accelerate_dev(h, eta, eta, 0.5, false);                              // :213
        elph_lincomb(h, h->d_dx, s2dt, eta, -dt, h->d_dSdx, 0.0, nullptr, nd);             // :216
        elph_lincomb(h, x, 1.0, x, 1.0, h->d_dx, 0.0, nullptr, nd);                        // :219
        elph_launch_update_model(h);                                                       // :222
        if (iters) *iters = i2.iters;  // the reference returns the second solve's count (:198)
    } else if (method == ELPH_LANGEVIN_HEUN) {
        elph_fourier_accelerate_dev(h, eta, eta, 0.5, false);                              // :293 xi
        elph_launch_update_model(h);                                                       // :296
        elph_calc_dSdx_dev(h, g1_dev, arn1, use_precond, h->d_dSdx, h->d_Minv, &i1);       // :298
        elph_fourier_accelerate_dev(h, h->d_dSdx, h->d_dSdx, 1.0, false);                  // :301
        elph_lincomb(h, h->d_dx, s2dt, eta, -dt, h->d_dSdx, 0.0, nullptr, nd);             // :304
        elph_lincomb(h, x, 1.0, x, 1.0, h->d_dx, 0.0, nullptr, nd);                        // :307
        elph_launch_update_model(h);                                                       // :308
        elph_calc_dSdx_dev(h, g2_dev, arn2, use_precond, h->d_dSdx2, h->d_Minv, &i2);      // :312
        elph_fourier_accelerate_dev(h, h->d_dSdx2, h->d_dSdx2, 1.0, false);                // :315
        elph_lincomb(h, x, 1.0, x, -1.0, h->d_dx, 0.0, nullptr, nd);                       // :318
        // x'' = x + sqrt(2dt) xi - dt (dG + dG')/2                                         :321
        elph_lincomb(h, h->d_dx, s2dt, eta, -0.5 * dt, h->d_dSdx, -0.5 * dt, h->d_dSdx2, nd);
        elph_lincomb(h, x, 1.0, x, 1.0, h->d_dx, 0.0, nullptr, nd);
        elph_launch_update_model(h);                                                       // :322
        if (iters) *iters = (i1.iters + i2.iters) / 2;                                     // div(iters1+iters2,2) :324
    } else {
        ELPH_REQUIRE(false, ELPH_ERR_INVALID, "unknown Langevin update method");
    }
    elph_trace_mark(h, "rest of the step");
    if (info1) *info1 = i1;
    if (info2) *info2 = i2;
}
