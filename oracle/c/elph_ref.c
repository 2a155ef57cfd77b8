/*
 * CPU restatement (plain C) of the reference's hot loops -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.
 *
 * Mirrors, loop for loop, the Julia code of cohensbw/ElPhDynamics v1.1.3 (which cannot run here: no
 * Julia in the image).  Host layout: index = site*Ltau + tau (src/Utilities.jl:12-15), 0-based.
 *   checkerboard_mul!            src/Checkerboard.jl:57-83     (bond-major, inner @simd loop over tau)
 *   checkerboard_transpose_mul!  src/Checkerboard.jl:149-175
 *   mulM!, mulMT!                src/HolsteinModels.jl:569-626, 631-684
 *   mulMTM!                      src/Models.jl:215-224
 *   solve! (plain CG)            src/IterativeSolvers.jl:239-314
 * Compiled with -O3 -ffast-math to match the reference's @fastmath @inbounds @simd loops.
 * Only tests/, bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke() may use it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int64_t N, L, Nb;
    const int64_t* nt;   /* 2*Nb, (i,j) pairs, 0-based, checkerboard order */
    const double* cosht; /* Nb */
    const double* sinht; /* Nb */
} ref_model;

static void checkerboard_mul(double* y, const ref_model* m) {
    const int64_t L = m->L;
    for (int64_t n = 0; n < m->Nb; ++n) {
        const double c = m->cosht[n], s = m->sinht[n];
        double* yi = y + m->nt[2 * n] * L;
        double* yj = y + m->nt[2 * n + 1] * L;
        for (int64_t t = 0; t < L; ++t) {
            const double t1 = yi[t], t2 = yj[t];
            yi[t] = c * t1 + s * t2;
            yj[t] = c * t2 + s * t1;
        }
    }
}

static void checkerboard_transpose_mul(double* y, const ref_model* m) {
    const int64_t L = m->L;
    for (int64_t n = m->Nb - 1; n >= 0; --n) {
        const double c = m->cosht[n], s = m->sinht[n];
        double* yi = y + m->nt[2 * n] * L;
        double* yj = y + m->nt[2 * n + 1] * L;
        for (int64_t t = 0; t < L; ++t) {
            const double t1 = yi[t], t2 = yj[t];
            yi[t] = c * t1 + s * t2;
            yj[t] = c * t2 + s * t1;
        }
    }
}

void ref_mulM(double* y, const ref_model* m, const double* expnV, const double* v) {
    const int64_t N = m->N, L = m->L;
    for (int64_t i = 0; i < N; ++i)
        for (int64_t t = 0; t < L; ++t) {
            const int64_t tm1 = (t == 0) ? L - 1 : t - 1;
            y[i * L + t] = expnV[i * L + t] * v[i * L + tm1];
        }
    checkerboard_mul(y, m);
    for (int64_t i = 0; i < N; ++i) {
        y[i * L] = v[i * L] + y[i * L];
        for (int64_t t = 1; t < L; ++t) y[i * L + t] = v[i * L + t] - y[i * L + t];
    }
}

void ref_mulMT(double* y, const ref_model* m, const double* expnV, const double* v) {
    const int64_t N = m->N, L = m->L;
    memcpy(y, v, (size_t)(N * L) * sizeof(double));
    checkerboard_transpose_mul(y, m);
    for (int64_t i = 0; i < N; ++i) {
        const double yL = v[i * L + L - 1] + expnV[i * L] * y[i * L];
        for (int64_t t = 0; t < L - 1; ++t) y[i * L + t] = v[i * L + t] - expnV[i * L + t + 1] * y[i * L + t + 1];
        y[i * L + L - 1] = yL;
    }
}

void ref_mulMTM(double* y, const ref_model* m, const double* expnV, const double* v, double* scratch) {
    ref_mulM(scratch, m, expnV, v);
    ref_mulMT(y, m, expnV, scratch);
}

static double dot(const double* a, const double* b, int64_t n) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

/* plain CG on A = M^T M; x in/out; work = 4*N*L doubles.  Returns the iteration count. */
int64_t ref_cg(double* x, const ref_model* m, const double* expnV, const double* b, double tol, int64_t maxiter,
               double kappa_max, double* work, double* eps_out) {
    const int64_t n = m->N * m->L;
    double *r = work, *p = work + n, *z = work + 2 * n, *scr = work + 3 * n;
    const double normb = sqrt(dot(b, b, n));
    ref_mulMTM(r, m, expnV, x, scr);
    for (int64_t i = 0; i < n; ++i) r[i] = b[i] - r[i];
    memcpy(p, r, (size_t)n * sizeof(double));
    double rdotr = dot(r, r, n);
    const double eps0 = sqrt(rdotr) / normb;
    double eps = eps0, kmin = 0.0;
    for (int64_t j = 1; j <= maxiter; ++j) {
        ref_mulMTM(z, m, expnV, p, scr);
        const double alpha = rdotr / dot(p, z, n);
        for (int64_t i = 0; i < n; ++i) x[i] += alpha * p[i];
        for (int64_t i = 0; i < n; ++i) r[i] -= alpha * z[i];
        const double nr = dot(r, r, n);
        eps = sqrt(nr) / normb;
        const double q = 2.0 * (double)j / log(2.0 * eps0 / eps);
        if (q * q > kmin) kmin = q * q;
        if (eps < tol || kmin > kappa_max) {
            if (eps_out) *eps_out = eps;
            return j;
        }
        const double beta = nr / rdotr;
        rdotr = nr;
        for (int64_t i = 0; i < n; ++i) p[i] = r[i] + beta * p[i];
    }
    if (eps_out) *eps_out = eps;
    return maxiter;
}

/* The same CG with the loops of one product split over OpenMP threads along tau (each (bond, tau) update is independent of
 * the other time slices, so every element sees exactly the operations of the serial code; only the summation order of the
 * dot products differs).  For the full-size parity tests (64x64xL400: 2000 iterations of 1.6 M points). */
static void mulMTM_mt(double* y, const ref_model* m, const double* expnV, const double* v, double* scr) {
    const int64_t N = m->N, L = m->L;
#pragma omp parallel
    {
        int nt = 1, id = 0;
#ifdef _OPENMP
        nt = omp_get_num_threads();
        id = omp_get_thread_num();
#endif
        const int64_t t0 = L * id / nt, t1 = L * (id + 1) / nt;
        /* scr = M v on [t0, t1) */
        for (int64_t i = 0; i < N; ++i)
            for (int64_t t = t0; t < t1; ++t) {
                const int64_t tm1 = (t == 0) ? L - 1 : t - 1;
                scr[i * L + t] = expnV[i * L + t] * v[i * L + tm1];
            }
        for (int64_t n = 0; n < m->Nb; ++n) {
            const double c = m->cosht[n], s = m->sinht[n];
            double* yi = scr + m->nt[2 * n] * L;
            double* yj = scr + m->nt[2 * n + 1] * L;
            for (int64_t t = t0; t < t1; ++t) {
                const double a = yi[t], b = yj[t];
                yi[t] = c * a + s * b;
                yj[t] = c * b + s * a;
            }
        }
        for (int64_t i = 0; i < N; ++i)
            for (int64_t t = t0; t < t1; ++t) scr[i * L + t] = (t == 0) ? v[i * L] + scr[i * L] : v[i * L + t] - scr[i * L + t];
        /* y = K^T scr on [t0, t1), then the shift by one slice needs the neighbours' ranges */
        for (int64_t i = 0; i < N; ++i)
            for (int64_t t = t0; t < t1; ++t) y[i * L + t] = scr[i * L + t];
        for (int64_t n = m->Nb - 1; n >= 0; --n) {
            const double c = m->cosht[n], s = m->sinht[n];
            double* yi = y + m->nt[2 * n] * L;
            double* yj = y + m->nt[2 * n + 1] * L;
            for (int64_t t = t0; t < t1; ++t) {
                const double a = yi[t], b = yj[t];
                yi[t] = c * a + s * b;
                yj[t] = c * b + s * a;
            }
        }
#pragma omp barrier
        /* out(t) = scr(t) -+ expnV(t+1) y(t+1): written into scr's place is not possible (y(t+1) is read by the neighbour), so
         * the result goes to a second pass over a private copy of the boundary */
        for (int64_t i = 0; i < N; ++i) {
            for (int64_t t = t0; t < t1; ++t) {
                const int64_t tp = (t == L - 1) ? 0 : t + 1;
                const double u = expnV[i * L + tp] * y[i * L + tp];
                scr[i * L + t] = (t == L - 1) ? scr[i * L + t] + u : scr[i * L + t] - u;
            }
        }
#pragma omp barrier
        for (int64_t i = 0; i < N; ++i)
            for (int64_t t = t0; t < t1; ++t) y[i * L + t] = scr[i * L + t];
    }
}

static double dot_mt(const double* a, const double* b, int64_t n) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

int64_t ref_cg_mt(double* x, const ref_model* m, const double* expnV, const double* b, double tol, int64_t maxiter,
                  double kappa_max, double* work, double* eps_out, int nthreads) {
    const int64_t n = m->N * m->L;
    double *r = work, *p = work + n, *z = work + 2 * n, *scr = work + 3 * n;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    const double normb = sqrt(dot_mt(b, b, n));
    mulMTM_mt(r, m, expnV, x, scr);
    for (int64_t i = 0; i < n; ++i) r[i] = b[i] - r[i];
    memcpy(p, r, (size_t)n * sizeof(double));
    double rdotr = dot_mt(r, r, n);
    const double eps0 = sqrt(rdotr) / normb;
    double eps = eps0, kmin = 0.0;
    for (int64_t j = 1; j <= maxiter; ++j) {
        mulMTM_mt(z, m, expnV, p, scr);
        const double alpha = rdotr / dot_mt(p, z, n);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            x[i] += alpha * p[i];
            r[i] -= alpha * z[i];
        }
        const double nr = dot_mt(r, r, n);
        eps = sqrt(nr) / normb;
        const double q = 2.0 * (double)j / log(2.0 * eps0 / eps);
        if (q * q > kmin) kmin = q * q;
        if (eps < tol || kmin > kappa_max) {
            if (eps_out) *eps_out = eps;
            return j;
        }
        const double beta = nr / rdotr;
        rdotr = nr;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) p[i] = r[i] + beta * p[i];
    }
    if (eps_out) *eps_out = eps;
    return maxiter;
}

/* ---- the second headline quantity on the host: the loops of a KPM-preconditioned Langevin step ----------------------------
 *   muldMdx!                         src/HolsteinModels.jl:691-755
 *   A v, A^T v, A^-1 v (tau-averaged B)  src/KPMPreconditioners.jl:387-420, 758-778
 *   mul!(::SymmetricKPMPreconditioner) frequency blocks, three-term Chebyshev recurrences  :606-693
 * The FFTs (FFTW in the reference), the <= 20 x 20 eigenvalue problem (LAPACK) and the coefficient DCT stay with NumPy /
 * SciPy in oracle/cfast.py, as the reference delegates them to libraries too. */
void ref_muldMdx(double* dMdx, const ref_model* m, const double* expnV, const double* x, const double* lam, const double* lam2,
                 double dtau, const double* u, const double* v, double* scratch) {
    const int64_t N = m->N, L = m->L;
    memcpy(scratch, u, (size_t)(N * L) * sizeof(double));
    checkerboard_transpose_mul(scratch, m);
    for (int64_t i = 0; i < N; ++i) {
        const double l1 = lam[i], l2 = lam2[i];
        dMdx[i * L] = scratch[i * L] * (-dtau * (l1 + 2 * l2 * x[i * L]) * expnV[i * L] * v[i * L + L - 1]);
        for (int64_t t = 1; t < L; ++t)
            dMdx[i * L + t] = scratch[i * L + t] * (dtau * (l1 + 2 * l2 * x[i * L + t]) * expnV[i * L + t] * v[i * L + t - 1]);
    }
}

/* one N-vector (real): y = cb(cbar, sbar) y in bond order / reverse order with +-s */
static void cb_vec(double* y, const ref_model* m, const double* cbar, const double* sbar, int reverse, double sign) {
    for (int64_t k = 0; k < m->Nb; ++k) {
        const int64_t n = reverse ? m->Nb - 1 - k : k;
        const double c = cbar[n], s = sign * sbar[n];
        const int64_t i = m->nt[2 * n], j = m->nt[2 * n + 1];
        const double a = y[i], b = y[j];
        y[i] = c * a + s * b;
        y[j] = c * b + s * a;
    }
}
/* mode 0: out = A v = cb diag(eVbar) v; 1: out = A^T v = diag(eVbar) cb^T v; 2: out = A^-1 v = (cb^-1 v) ./ eVbar */
void ref_mulA(double* out, const ref_model* m, const double* eVbar, const double* cbar, const double* sbar, const double* v, int mode) {
    const int64_t N = m->N;
    if (mode == 0) {
        for (int64_t i = 0; i < N; ++i) out[i] = eVbar[i] * v[i];
        cb_vec(out, m, cbar, sbar, 0, 1.0);
    } else if (mode == 1) {
        memcpy(out, v, (size_t)N * sizeof(double));
        cb_vec(out, m, cbar, sbar, 1, 1.0);
        for (int64_t i = 0; i < N; ++i) out[i] *= eVbar[i];
    } else {
        memcpy(out, v, (size_t)N * sizeof(double));
        cb_vec(out, m, cbar, sbar, 1, -1.0);
        for (int64_t i = 0; i < N; ++i) out[i] /= eVbar[i];
    }
}

/* complex N-vector as interleaved (re, im): A' v = (A v) / lam_mag - (lam_avg / lam_mag) v, A real */
static void mulAprime_c(double* out, const double* v, const ref_model* m, const double* eVbar, const double* cbar, const double* sbar,
                        double lam_avg, double lam_mag, int transposed) {
    const int64_t N = m->N;
    if (!transposed)
        for (int64_t i = 0; i < N; ++i) { out[2 * i] = eVbar[i] * v[2 * i]; out[2 * i + 1] = eVbar[i] * v[2 * i + 1]; }
    else
        memcpy(out, v, (size_t)(2 * N) * sizeof(double));
    for (int64_t k = 0; k < m->Nb; ++k) {
        const int64_t n = transposed ? m->Nb - 1 - k : k;
        const double c = cbar[n], s = sbar[n];
        const int64_t i = m->nt[2 * n], j = m->nt[2 * n + 1];
        const double ar = out[2 * i], ai = out[2 * i + 1], br = out[2 * j], bi = out[2 * j + 1];
        out[2 * i] = c * ar + s * br;
        out[2 * i + 1] = c * ai + s * bi;
        out[2 * j] = c * br + s * ar;
        out[2 * j + 1] = c * bi + s * ai;
    }
    const double a = 1.0 / lam_mag, b = lam_avg / lam_mag;
    if (transposed)
        for (int64_t i = 0; i < N; ++i) {
            out[2 * i] = a * (eVbar[i] * out[2 * i]) - b * v[2 * i];
            out[2 * i + 1] = a * (eVbar[i] * out[2 * i + 1]) - b * v[2 * i + 1];
        }
    else
        for (int64_t i = 0; i < N; ++i) {
            out[2 * i] = a * out[2 * i] - b * v[2 * i];
            out[2 * i + 1] = a * out[2 * i + 1] - b * v[2 * i + 1];
        }
}

/* out = sum_m c_m T_m(A') v (conj: conjugated coefficients); work: 3 complex N-vectors */
static void kpm_poly(double* out, const double* v, int64_t order, const double* coeff, int conj, int transposed, const ref_model* m,
                     const double* eVbar, const double* cbar, const double* sbar, double lam_avg, double lam_mag, double* work) {
    const int64_t N = m->N;
    const double sg = conj ? -1.0 : 1.0;
    double cr = coeff[0], ci = sg * coeff[1];
    for (int64_t i = 0; i < N; ++i) {
        out[2 * i] = cr * v[2 * i] - ci * v[2 * i + 1];
        out[2 * i + 1] = cr * v[2 * i + 1] + ci * v[2 * i];
    }
    if (order <= 1) return;
    double *up = work, *un = work + 2 * N, *ux = work + 4 * N;
    memcpy(un, v, (size_t)(2 * N) * sizeof(double));
    mulAprime_c(ux, un, m, eVbar, cbar, sbar, lam_avg, lam_mag, transposed);
    for (int64_t n = 2;; ++n) {
        double* t = up; up = un; un = ux; ux = t;              /* (u_prev, u_n) = (u_n, u_next) */
        cr = coeff[2 * (n - 1)]; ci = sg * coeff[2 * (n - 1) + 1];
        for (int64_t i = 0; i < N; ++i) {
            out[2 * i] += cr * un[2 * i] - ci * un[2 * i + 1];
            out[2 * i + 1] += cr * un[2 * i + 1] + ci * un[2 * i];
        }
        if (n == order) break;
        mulAprime_c(ux, un, m, eVbar, cbar, sbar, lam_avg, lam_mag, transposed);
        for (int64_t i = 0; i < 2 * N; ++i) ux[i] = 2.0 * ux[i] - up[i];
    }
}

/* all frequency blocks of the preconditioner: a2T[w] = M^-1[w,w] M^-T[w,w] a1T[w], w < Lo2; a1T, a2T: [omega][site] complex.
 * Frequencies are independent: one OpenMP thread per block when nthreads > 1.  Returns the threads used. */
int ref_kpm_blocks(const ref_model* m, const double* eVbar, const double* cbar, const double* sbar, double lam_avg, double lam_mag,
                   int64_t Lo2, const int64_t* order, const int64_t* coeff_off, const double* coeff, const double* a1T, double* a2T,
                   int nthreads) {
    const int64_t N = m->N;
    int used = 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
    {
#pragma omp single
        used = omp_get_num_threads();
        double* work = (double*)malloc((size_t)(8 * N) * sizeof(double));
#pragma omp for schedule(dynamic, 1)
        for (int64_t w = 0; w < Lo2; ++w) {
            double* tmp = work + 6 * N;
            kpm_poly(tmp, a1T + 2 * w * N, order[w], coeff + 2 * coeff_off[w], 1, 1, m, eVbar, cbar, sbar, lam_avg, lam_mag, work);
            kpm_poly(a2T + 2 * w * N, tmp, order[w], coeff + 2 * coeff_off[w], 0, 0, m, eVbar, cbar, sbar, lam_avg, lam_mag, work);
        }
        free(work);
    }
#else
    (void)nthreads;
    double* work = (double*)malloc((size_t)(8 * N) * sizeof(double));
    for (int64_t w = 0; w < Lo2; ++w) {
        double* tmp = work + 6 * N;
        kpm_poly(tmp, a1T + 2 * w * N, order[w], coeff + 2 * coeff_off[w], 1, 1, m, eVbar, cbar, sbar, lam_avg, lam_mag, work);
        kpm_poly(a2T + 2 * w * N, tmp, order[w], coeff + 2 * coeff_off[w], 0, 0, m, eVbar, cbar, sbar, lam_avg, lam_mag, work);
    }
    free(work);
#endif
    return used;
}

void ref_mulMTM_mt(double* y, const ref_model* m, const double* expnV, const double* v, double* scratch, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    mulMTM_mt(y, m, expnV, v, scratch);
}

/* Throughput driver: `nrep` independent replicas (the reference's own scale-out: independent runs, one
 * thread each -- BLAS/FFTW are pinned to 1 thread, src/ElPhDynamics.jl:74-75), `reps` M^T M products each.
 * v, y, expnV, scratch: nrep contiguous blocks of N*L doubles.  Returns the number of threads used. */
int ref_mulMTM_replicas(const ref_model* m, int64_t nrep, int64_t reps, const double* expnV, double* v, double* y,
                        double* scratch, int nthreads) {
    const int64_t n = m->N * m->L;
    int used = 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
    {
#pragma omp single
        used = omp_get_num_threads();
#pragma omp for schedule(static)
        for (int64_t r = 0; r < nrep; ++r)
            for (int64_t k = 0; k < reps; ++k) ref_mulMTM(y + r * n, m, expnV + r * n, v + r * n, scratch + r * n);
    }
#else
    (void)nthreads;
    for (int64_t r = 0; r < nrep; ++r)
        for (int64_t k = 0; k < reps; ++k) ref_mulMTM(y + r * n, m, expnV + r * n, v + r * n, scratch + r * n);
#endif
    return used;
}
