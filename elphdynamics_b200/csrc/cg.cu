// K3: conjugate gradient on A = M^T M, resident on the device.
//
// Reference: src/IterativeSolvers.jl:153-234 (preconditioned), :239-314 (plain);
// wrappers with the true-residual check, flags and fallback: src/Models.jl:74-186.
//
// One iteration is two kernels (plus the KPM apply when preconditioned):
//   K_A  matvec_kernel<MTM, FUSEP>:  p_new = z_or_r + beta*p_old (on the fly, double-buffered p),
//        Ap = M^T M p_new, per-CTA partials of p.Ap; last CTA: alpha = (r.z)/(p.Ap)
//   K_B  cg_xr_kernel:  x += alpha p, r -= alpha Ap, per-CTA partials of |r|^2; last CTA applies the
//        reference stop rule  eps < tol || kappa_min > kappa_max  and latches `done`.
// All scalars live in a CgScalars block in device memory; the host only polls the
// latch every few iterations, so there is no host round trip per iteration and the
// iteration count is exact (launches after the latch are no-ops).  Reductions use a
// fixed order (per-CTA partial, then index-ordered fold), so runs are bit-reproducible.
#include "elph_internal.cuh"

namespace {

constexpr int kT = 256;

__device__ __forceinline__ double block_sum_cg(double x, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();  // protect red[] against a previous use
    if (l == 0) red[w] = x;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int k = 0; k < nw; ++k) t += red[k];
    }
    return t;
}

// returns true on every thread of the CTA that finished last
__device__ __forceinline__ bool last_block(unsigned int* ticket, bool* flag) {
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int n = atomicAdd(ticket, 1u);
        *flag = (n == gridDim.x - 1);
    }
    __syncthreads();
    const bool r = *flag;
    if (r) __threadfence();
    return r;
}

__device__ __forceinline__ double fold(const double* partial, int n, double* red) {
    double s = 0.0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) s += ((const volatile double*)partial)[k];
    return block_sum_cg(s, red);
}

// r = b - Ax (Ax in `ax`), partials of |b|^2 and |r|^2; last CTA initialises the scalar block.
__global__ void __launch_bounds__(kT) cg_init_kernel(const double* __restrict__ b, const double* __restrict__ ax,
                                                     double* __restrict__ r, long long n, double* __restrict__ partial,
                                                     CgScalars* S, unsigned int* ticket, double tol, double kappa_max,
                                                     long long maxiter, int precond) {
    __shared__ double red[32];
    __shared__ bool flag;
    double sb = 0.0, sr = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double bv = b[i];
        const double rv = bv - ax[i];
        r[i] = rv;
        sb += bv * bv;
        sr += rv * rv;
    }
    const double tb = block_sum_cg(sb, red);
    const double tr = block_sum_cg(sr, red);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = tb;
        partial[gridDim.x + blockIdx.x] = tr;
    }
    if (last_block(ticket, &flag)) {
        const double bb = fold(partial, gridDim.x, red);
        const double rr = fold(partial + gridDim.x, gridDim.x, red);
        if (threadIdx.x == 0) {
            S->normb = sqrt(bb);
            S->eps0 = sqrt(rr) / sqrt(bb);
            S->eps = S->eps0;
            S->kappa_min = 0.0;
            S->rdotz = precond ? 0.0 : rr;
            S->beta = 0.0;
            S->alpha = 0.0;
            S->pAp = 0.0;
            S->tol = tol;
            S->kappa_max = kappa_max;
            S->iter = 0;
            S->maxiter = maxiter;
            S->done = 0;
            *ticket = 0u;
        }
    }
}

// x += alpha p ; r -= alpha Ap ; |r|^2 ; stop rule (src/IterativeSolvers.jl:205-219 / :287-301)
__global__ void __launch_bounds__(kT) cg_xr_kernel(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
                                                   const double* __restrict__ ap, long long n, double* __restrict__ partial,
                                                   CgScalars* S, unsigned int* ticket, int precond) {
    __shared__ double red[32];
    __shared__ bool flag;
    if (S->done) return;
    const double alpha = S->alpha;
    double sr = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        x[i] = fma(alpha, p[i], x[i]);
        const double rv = fma(-alpha, ap[i], r[i]);
        r[i] = rv;
        sr += rv * rv;
    }
    const double t = block_sum_cg(sr, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
    if (last_block(ticket, &flag)) {
        const double rr = fold(partial, gridDim.x, red);
        if (threadIdx.x == 0) {
            const long long j = S->iter + 1;
            const double eps = sqrt(rr) / S->normb;
            const double lg = log(2.0 * S->eps0 / eps);
            const double q = 2.0 * (double)j / lg;
            const double kap = q * q;
            double kmin = S->kappa_min;
            if (kap > kmin) kmin = kap;  // NaN never wins, like max() on the accumulated bound
            S->kappa_min = kmin;
            S->eps = eps;
            S->iter = j;
            if (eps < S->tol || kmin > S->kappa_max || j >= S->maxiter) {
                S->done = 1;
            } else if (!precond) {
                S->beta = rr / S->rdotz;
                S->rdotz = rr;
            }
            *ticket = 0u;
        }
    }
}

// preconditioned: partials of r.z ; last CTA: beta = (r.z)_new/(r.z)_old  (src/IterativeSolvers.jl:221-227)
__global__ void __launch_bounds__(kT) cg_rz_kernel(const double* __restrict__ r, const double* __restrict__ z, long long n,
                                                   double* __restrict__ partial, CgScalars* S, unsigned int* ticket) {
    __shared__ double red[32];
    __shared__ bool flag;
    if (S->done) return;
    double s = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        s += r[i] * z[i];
    const double t = block_sum_cg(s, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
    if (last_block(ticket, &flag)) {
        const double rz = fold(partial, gridDim.x, red);
        if (threadIdx.x == 0) {
            S->beta = (S->iter == 0) ? 0.0 : rz / S->rdotz;
            S->rdotz = rz;
            *ticket = 0u;
        }
    }
}

// out[0] = sum a.b   (generic deterministic dot; out[1] = sum b.b when want_bb)
__global__ void __launch_bounds__(kT) dot_kernel(const double* __restrict__ a, const double* __restrict__ b, long long n,
                                                 double* __restrict__ partial, double* __restrict__ out, unsigned int* ticket,
                                                 int mode) {
    // mode 0: out[0] = a.b ; mode 1: out[0] = |a-b|^2, out[1] = |b|^2
    __shared__ double red[32];
    __shared__ bool flag;
    double s0 = 0.0, s1 = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double av = a[i], bv = b[i];
        if (mode == 0) {
            s0 += av * bv;
        } else {
            const double d = av - bv;
            s0 += d * d;
            s1 += bv * bv;
        }
    }
    const double t0 = block_sum_cg(s0, red);
    const double t1 = block_sum_cg(s1, red);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = t0;
        partial[gridDim.x + blockIdx.x] = t1;
    }
    if (last_block(ticket, &flag)) {
        const double r0 = fold(partial, gridDim.x, red);
        const double r1 = fold(partial + gridDim.x, gridDim.x, red);
        if (threadIdx.x == 0) {
            out[0] = r0;
            out[1] = r1;
            *ticket = 0u;
        }
    }
}

int vec_blocks(elph_handle* h, int64_t n) {
    int64_t b = (n + kT - 1) / kT;
    const int64_t cap = std::min<int64_t>(2LL * h->sm_count, h->partial_cap / 2);
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace

void elph_dot_async(elph_handle* h, const double* a, const double* b, int64_t n, double* d_out) {
    const int blocks = vec_blocks(h, n);
    dot_kernel<<<blocks, kT, 0, h->stream>>>(a, b, n, h->d_partial, d_out, h->d_ticket, 0);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}

void elph_diffnorm2_async(elph_handle* h, const double* a, const double* b, int64_t n, double* d_out2) {
    const int blocks = vec_blocks(h, n);
    dot_kernel<<<blocks, kT, 0, h->stream>>>(a, b, n, h->d_partial, d_out2, h->d_ticket, 1);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
}

// solve!(x, A, b, cg[, P]) on device pointers.  x: in = initial guess, out = solution.
void elph_cg_device(elph_handle* h, const double* b_dev, double* x_dev, bool use_precond, double tol, int64_t maxiter,
                    int64_t* iters, double* eps) {
    if (tol == 0.0) tol = h->cg_tol;
    if (maxiter == 0) maxiter = h->cg_maxiter;
    if (use_precond) ELPH_REQUIRE(h->kpm.configured && h->kpm.ever_setup, ELPH_ERR_STATE,
                                  "preconditioned solve requested before elph_kpm_setup");
    const bool precond = use_precond && h->kpm.active;  // inactive KPM behaves as the identity (:475-478)
    const int64_t n = h->Ndim;
    const int vb = vec_blocks(h, n);
    cudaStream_t st = h->stream;

    // r0 = b - A x0
    MatvecArgs m0;
    m0.v = x_dev;
    m0.y = h->d_z;
    elph_launch_matvec(h, MODE_MTM, m0);
    cg_init_kernel<<<vb, kT, 0, st>>>(b_dev, h->d_z, h->d_r, n, h->d_partial, h->d_cg, h->d_ticket, tol, h->cg_kappa_max,
                                      (long long)maxiter, precond ? 1 : 0);
    ELPH_CUDA(cudaGetLastError());
    h->launches++;
    double* zprec = h->d_res;  // z = P^-1 r lives in the residual scratch during a preconditioned solve
    // Preconditioned solve on a 32-wide square lattice: the whole loop, preconditioner included, is ONE persistent kernel
    // (pcg_fused.cu) -- four grid barriers per iteration instead of four launches.
    if (precond && elph_pcg_fused(h, x_dev, zprec)) {
        ELPH_CUDA(cudaMemcpyAsync(h->h_cg, h->d_cg, sizeof(CgScalars), cudaMemcpyDeviceToHost, st));
        ELPH_CUDA(cudaStreamSynchronize(st));
        ELPH_REQUIRE(h->h_cg->done == 1, ELPH_ERR_STATE, "fused preconditioned CG: a grid barrier timed out");
        if (iters) *iters = h->h_cg->iter;
        if (eps) *eps = h->h_cg->eps;
        return;
    }
    if (precond) {
        elph_kpm_apply_dev(h, h->d_r, zprec);
        cg_rz_kernel<<<vb, kT, 0, st>>>(h->d_r, zprec, n, h->d_partial, h->d_cg, h->d_ticket);
        ELPH_CUDA(cudaGetLastError());
        h->launches++;
    }

    ELPH_CUDA(cudaMemsetAsync(h->d_p[1], 0, n * sizeof(double), st));  // p_old of the first iteration (beta = 0)
    // Unpreconditioned solve on a square lattice whose time slices are all co-resident: the whole loop is ONE cooperative
    // persistent kernel (cg_persistent.cu) -- no launches, two grid barriers per iteration.
    if (!precond && ((h->use_persistent && elph_cg_single_reduction(h, x_dev)) || elph_cg_persistent(h, x_dev))) {
        ELPH_CUDA(cudaMemcpyAsync(h->h_cg, h->d_cg, sizeof(CgScalars), cudaMemcpyDeviceToHost, st));
        ELPH_CUDA(cudaStreamSynchronize(st));
        if (iters) *iters = h->h_cg->iter;
        if (eps) *eps = h->h_cg->eps;
        return;
    }
    // The host polls the convergence latch every `check_every` iterations (even, so the p double buffer is back at
    // parity 0).  A block of check_every iterations is captured ONCE into a CUDA graph (per solution vector /
    // preconditioner state) and replayed: one graph launch instead of 2-6 kernel launches per iteration.
    const int check_every = precond ? 4 : 8;
    h->kpm_skip_enabled = true;
    int64_t launched = 0;
    auto body = [&](int64_t niter) {
        int parity = 0;
        for (int64_t k = 0; k < niter; ++k) {
            MatvecArgs m;
            m.v = nullptr;
            m.y = h->d_z;
            m.partial_dot = h->d_partial;
            m.cg_pr = precond ? zprec : h->d_r;
            m.cg_pold = h->d_p[parity ^ 1];
            m.cg_pnew = h->d_p[parity];
            m.cg_S = h->d_cg;
            m.cg_ticket = h->d_ticket;
            elph_launch_matvec(h, MODE_MTM, m);
            if (precond && h->pcg_fuse) {
                // 3 more kernels: [x/r update + stop rule + FFT(r)] [Chebyshev recurrences] [iFFT -> z, r.z -> beta]
                KpmCgFuse f{x_dev, h->d_r, h->d_p[parity], h->d_z};
                elph_kpm_apply_dev_cg(h, h->d_r, zprec, &f);
                parity ^= 1;
                continue;
            }
            cg_xr_kernel<<<vb, kT, 0, st>>>(x_dev, h->d_r, h->d_p[parity], h->d_z, n, h->d_partial, h->d_cg, h->d_ticket,
                                            precond ? 1 : 0);
            ELPH_CUDA(cudaGetLastError());
            h->launches++;
            if (precond) {
                elph_kpm_apply_dev(h, h->d_r, zprec);  // kernels are no-ops once S->done is latched
                cg_rz_kernel<<<vb, kT, 0, st>>>(h->d_r, zprec, n, h->d_partial, h->d_cg, h->d_ticket);
                ELPH_CUDA(cudaGetLastError());
                h->launches++;
            }
            parity ^= 1;
        }
    };
    auto graph_for = [&]() -> CgGraph* {
        for (auto& g : h->cg_graphs)
            if (g.x == x_dev && g.precond == precond && g.kpm_version == h->kpm_version && g.chunk == h->chunk_override &&
                g.sq_disable == h->sq_disable && g.stream == st)
                return &g;
        CgGraph g;
        g.x = x_dev; g.precond = precond; g.kpm_version = h->kpm_version; g.chunk = h->chunk_override;
        g.sq_disable = h->sq_disable; g.stream = st;
        const int64_t before = h->launches;
        cudaGraph_t graph = nullptr;
        ELPH_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
        try {
            body(check_every);
        } catch (...) {
            cudaStreamEndCapture(st, &graph);
            if (graph) cudaGraphDestroy(graph);
            throw;
        }
        ELPH_CUDA(cudaStreamEndCapture(st, &graph));
        g.nlaunch = h->launches - before;
        h->launches = before;
        ELPH_CUDA(cudaGraphInstantiate(&g.exec, graph, 0));
        ELPH_CUDA(cudaGraphDestroy(graph));
        if (h->cg_graphs.size() >= 16) {   // bounded cache
            cudaGraphExecDestroy(h->cg_graphs.front().exec);
            h->cg_graphs.erase(h->cg_graphs.begin());
        }
        h->cg_graphs.push_back(g);
        return &h->cg_graphs.back();
    };
    while (true) {
        const int64_t todo = std::min<int64_t>(check_every, maxiter - launched);
        // the first block runs as plain launches (it also performs the one-time kernel attribute set-up)
        if (h->use_graphs && todo == check_every && launched > 0 && st != nullptr) {
            CgGraph* g = graph_for();
            ELPH_CUDA(cudaGraphLaunch(g->exec, st));
            h->launches += g->nlaunch;
        } else {
            body(todo);
        }
        launched += todo;
        ELPH_CUDA(cudaMemcpyAsync(h->h_cg, h->d_cg, sizeof(CgScalars), cudaMemcpyDeviceToHost, st));
        ELPH_CUDA(cudaStreamSynchronize(st));
        if (h->h_cg->done || launched >= maxiter) break;
    }
    h->kpm_skip_enabled = false;
    if (iters) *iters = h->h_cg->iter;
    if (eps) *eps = h->h_cg->eps;
}

// ldiv!(x, model, b[, P]; maxiter): src/Models.jl:74-186
static void residual_check(elph_handle* h, const double* b_dev, double* x_dev, double* resid) {
    MatvecArgs m;
    m.v = x_dev;
    m.y = h->d_z;
    elph_launch_matvec(h, MODE_MTM, m);
    elph_diffnorm2_async(h, h->d_z, b_dev, h->Ndim, h->d_scal);
    ELPH_CUDA(cudaMemcpyAsync(h->h_scal, h->d_scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    ELPH_CUDA(cudaStreamSynchronize(h->stream));
    *resid = sqrt(h->h_scal[0]) / sqrt(h->h_scal[1]);
}

void elph_solve_device(elph_handle* h, const double* b_dev, double* x_dev, bool use_precond, double tol_power,
                       elph_solve_info* info) {
    // HMC runs its solves with tol^power and restores tol afterwards (src/HMC.jl:838-842,909-911)
    const double tol = (tol_power == 1.0) ? h->cg_tol : pow(h->cg_tol, tol_power);
    const int64_t maxiter = h->cg_maxiter;
    elph_solve_info out = {};
    int64_t it = 0;
    double eps = 0.0, resid = 0.0;
    const bool precond = use_precond && h->kpm.configured;
    if (precond) {
        elph_cg_device(h, b_dev, x_dev, true, tol, maxiter, &it, &eps);
        out.pcg_iters = it;
        residual_check(h, b_dev, x_dev, &resid);
        int flag = 0;
        if (resid > sqrt(tol)) {
            flag = (it == maxiter) ? 1 : 2;
            ELPH_CUDA(cudaMemsetAsync(x_dev, 0, h->Ndim * sizeof(double), h->stream));
        }
        out.iters = it;
        out.residual = resid;
        out.flag = flag;
        if (flag > 0) {
            // retry without preconditioner at 10*maxiter (src/Models.jl:129-133)
            out.used_fallback = 1;
            elph_cg_device(h, b_dev, x_dev, false, tol, 10 * maxiter, &it, &eps);
            residual_check(h, b_dev, x_dev, &resid);
            flag = 0;
            if (resid > sqrt(tol)) {
                flag = (it == h->cg_maxiter) ? 1 : 2;  // compares with solver.maxiter (src/Models.jl:160)
                ELPH_CUDA(cudaMemsetAsync(x_dev, 0, h->Ndim * sizeof(double), h->stream));
            }
            out.iters = it;
            out.residual = resid;
            out.flag = flag;
        }
    } else {
        elph_cg_device(h, b_dev, x_dev, false, tol, maxiter, &it, &eps);
        out.pcg_iters = it;
        residual_check(h, b_dev, x_dev, &resid);
        int flag = 0;
        if (resid > sqrt(tol)) {
            flag = (it == h->cg_maxiter) ? 1 : 2;
            ELPH_CUDA(cudaMemsetAsync(x_dev, 0, h->Ndim * sizeof(double), h->stream));
        }
        out.iters = it;
        out.residual = resid;
        out.flag = flag;
    }
    if (info) *info = out;
}

// ---- batched solves: nrhs right-hand sides on the same field (the n_v measurement vectors of
// update!(Gr, ...), src/GreensFunctions.jl:201-234, and the two pseudofermion flavours of HMC, src/HMC.jl:820-915) ------
static void batch_reserve(elph_handle* h, int nrhs) {
    auto& B = h->batch;
    if (B.cap >= nrhs) return;
    for (void* p : {(void*)B.x, (void*)B.r, (void*)B.p0, (void*)B.p1, (void*)B.partial, (void*)B.bar, (void*)B.S})
        if (p) ELPH_CUDA(cudaFree(p));
    const size_t n = (size_t)h->Ndim;
    B.x = elph_dalloc<double>(n * nrhs);
    B.r = elph_dalloc<double>(n * nrhs);
    B.p0 = elph_dalloc<double>(n * nrhs);
    B.p1 = elph_dalloc<double>(n * nrhs);
    B.partial = elph_dalloc<double>((size_t)2 * h->L * nrhs);
    B.bar = elph_dalloc<unsigned int>(nrhs);
    B.S = elph_dalloc<CgScalars>(nrhs);
    B.hS.resize(nrhs);
    B.cap = nrhs;
}

void elph_solve_batch_device(elph_handle* h, int nrhs, const double* const* b_dev, double* const* x_dev, bool use_precond,
                             double tol_power, elph_solve_info* infos) {
    ELPH_REQUIRE(nrhs >= 1 && nrhs <= 4096, ELPH_ERR_INVALID, "number of right-hand sides out of range");
    const double tol = (tol_power == 1.0) ? h->cg_tol : pow(h->cg_tol, tol_power);
    const int64_t n = h->Ndim;
    cudaStream_t st = h->stream;
    bool done = false;
    if (!(use_precond && h->kpm.configured) && nrhs > 1 && h->use_persistent && !h->sharded) {
        batch_reserve(h, nrhs);
        auto& B = h->batch;
        const int vb = vec_blocks(h, n);
        ELPH_CUDA(cudaMemsetAsync(B.x, 0, (size_t)n * nrhs * sizeof(double), st));   // x0 = 0 (every caller of ldiv! zeroes x)
        ELPH_CUDA(cudaMemsetAsync(B.p1, 0, (size_t)n * nrhs * sizeof(double), st));
        for (int k = 0; k < nrhs; ++k) {
            // r0 = b - A 0 = b and the scalar block, with the same kernel as the single solve
            cg_init_kernel<<<vb, kT, 0, st>>>(b_dev[k], B.x + (size_t)k * n, B.r + (size_t)k * n, n, h->d_partial, B.S + k,
                                              h->d_ticket, tol, h->cg_kappa_max, (long long)h->cg_maxiter, 0);
            ELPH_CUDA(cudaGetLastError());
            h->launches++;
        }
        CgBatchBufs io;
        io.x = B.x; io.R = B.r; io.P0 = B.p0; io.P1 = B.p1; io.partial = B.partial; io.bar = B.bar; io.S = B.S;
        io.vstride = n; io.pstride = 2 * h->L;
        if (elph_cg_persistent_batch(h, nrhs, io)) {
            ELPH_CUDA(cudaMemcpyAsync(B.hS.data(), B.S, nrhs * sizeof(CgScalars), cudaMemcpyDeviceToHost, st));
            for (int k = 0; k < nrhs; ++k)
                ELPH_CUDA(cudaMemcpyAsync(x_dev[k], B.x + (size_t)k * n, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
            ELPH_CUDA(cudaStreamSynchronize(st));
            for (int k = 0; k < nrhs; ++k) {
                elph_solve_info out = {};
                double resid = 0.0;
                residual_check(h, b_dev[k], x_dev[k], &resid);
                out.iters = out.pcg_iters = B.hS[k].iter;
                out.residual = resid;
                if (resid > sqrt(tol)) {
                    out.flag = (out.iters == h->cg_maxiter) ? 1 : 2;
                    ELPH_CUDA(cudaMemsetAsync(x_dev[k], 0, n * sizeof(double), st));
                }
                if (infos) infos[k] = out;
            }
            done = true;
        }
    }
    if (!done) {
        for (int k = 0; k < nrhs; ++k) {
            ELPH_CUDA(cudaMemsetAsync(x_dev[k], 0, n * sizeof(double), st));
            elph_solve_device(h, b_dev[k], x_dev[k], use_precond, tol_power, infos ? infos + k : nullptr);
        }
    }
}
