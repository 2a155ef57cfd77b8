/* A plain C99 caller of libelph_b200.so: what a compiled host program (or Julia's ccall) sees.
 *
 *   abi_check            links every symbol of include/elph_b200.h (SYMBOL_TABLE is generated from the header by the test),
 *                        checks elph_version and the error path of a null handle -- no GPU needed;
 *   abi_check gpu        additionally builds a small Holstein model on a periodic chain (the table assembly of
 *                        initialize_model!, src/HolsteinModels.jl:484-517, done here by hand for 6 sites), runs mulM!,
 *                        mulMT!, mulMTM! and a CG solve through the ABI and checks the operator identities the reference
 *                        relies on: <u, M v> = <M^T u, v>, M^T M v = M^T (M v), |A x - b| <= tol |b|.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "elph_b200.h"

#ifdef SYMBOL_TABLE
static void* const all_symbols[] = {SYMBOL_TABLE};
#else
static void* const all_symbols[] = {(void*)elph_version};
#endif

#define CHECK(cond, msg)                                         \
    do {                                                         \
        if (!(cond)) {                                           \
            fprintf(stderr, "abi_check FAILED: %s\n", msg);      \
            return 1;                                            \
        }                                                        \
    } while (0)

static double dot(const double* a, const double* b, int n) {
    double s = 0.0;
    int i;
    for (i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

static int gpu_part(void) {
    enum { N = 6, L = 5, NB = 6, NDIM = N * L };
    /* periodic chain of 6 sites, bonds (i, i+1): colour 1 = (1,2) (3,4) (5,6), colour 2 = (2,3) (4,5) (1,6); 1-based,
     * first row < second row, checkerboard order -- what model.neighbor_table holds in the reference */
    const int64_t nt[2 * NB] = {1, 2, 3, 4, 5, 6, 2, 3, 4, 5, 1, 6};
    const double dtau = 0.1, t = 1.0;
    double cosht[NB], sinht[NB], lam[N], lam2[N], mu[N], omega[N], omega4[N], x[NDIM], u[NDIM], v[NDIM];
    double Mv[NDIM], Mtu[NDIM], MtMv[NDIM], MtMv2[NDIM], b[NDIM], sol[NDIM], chk[NDIM];
    elph_config cfg;
    elph_handle* h = NULL;
    elph_solve_info info;
    int i;
    unsigned s = 12345u;
    for (i = 0; i < NB; ++i) { cosht[i] = cosh(dtau * t); sinht[i] = sinh(dtau * t); }
    for (i = 0; i < N; ++i) { lam[i] = 1.0; lam2[i] = 0.0; mu[i] = -0.3; omega[i] = 1.0; omega4[i] = 0.0; }
    for (i = 0; i < NDIM; ++i) {   /* a fixed pseudo-random field and vectors (LCG: no RNG on the parity path) */
        s = s * 1664525u + 1013904223u; x[i] = (double)(s >> 8) / 16777216.0 - 0.5;
        s = s * 1664525u + 1013904223u; u[i] = (double)(s >> 8) / 16777216.0 - 0.5;
        s = s * 1664525u + 1013904223u; v[i] = (double)(s >> 8) / 16777216.0 - 0.5;
    }
    memset(&cfg, 0, sizeof(cfg));
    cfg.model = ELPH_MODEL_HOLSTEIN; cfg.index_base = 1; cfg.device = -1;
    cfg.Ltau = L; cfg.Nsites = N; cfg.Nbonds = NB; cfg.Nph = N; cfg.dtau = dtau;
    cfg.neighbor_table = nt; cfg.cosht = cosht; cfg.sinht = sinht; cfg.lambda = lam; cfg.lambda2 = lam2; cfg.mu = mu;
    cfg.omega = omega; cfg.omega4 = omega4; cfg.cg_tol = 1e-10; cfg.cg_maxiter = 1000;
    CHECK(elph_create(&cfg, &h) == ELPH_OK && h, "elph_create");
    CHECK(elph_set_x(h, x) == ELPH_OK && elph_update_model(h) == ELPH_OK, "set_x / update_model");
    CHECK(elph_mulM(h, v, Mv) == ELPH_OK && elph_mulMT(h, u, Mtu) == ELPH_OK, "mulM / mulMT");
    CHECK(fabs(dot(u, Mv, NDIM) - dot(Mtu, v, NDIM)) <= 1e-13 * sqrt(dot(u, u, NDIM) * dot(Mv, Mv, NDIM)), "<u, M v> = <M^T u, v>");
    CHECK(elph_mulMTM(h, v, MtMv) == ELPH_OK && elph_mulMT(h, Mv, MtMv2) == ELPH_OK, "mulMTM");
    for (i = 0; i < NDIM; ++i) CHECK(fabs(MtMv[i] - MtMv2[i]) <= 1e-13 * (1.0 + fabs(MtMv2[i])), "M^T M v = M^T (M v)");
    CHECK(elph_mulMT(h, u, b) == ELPH_OK, "b = M^T u");
    memset(sol, 0, sizeof(sol));
    CHECK(elph_solve(h, b, sol, 0, 1.0, &info) == ELPH_OK && info.flag == 0, "elph_solve");
    CHECK(elph_mulMTM(h, sol, chk) == ELPH_OK, "check product");
    for (i = 0; i < NDIM; ++i) chk[i] -= b[i];
    CHECK(sqrt(dot(chk, chk, NDIM) / dot(b, b, NDIM)) <= 1e-9, "true residual of the solve");
    /* error path: solver misuse is a status + message, never an abort */
    CHECK(elph_mulM(h, NULL, Mv) != ELPH_OK && strlen(elph_last_error(h)) > 0, "null input must be rejected with a message");
    CHECK(elph_destroy(h) == ELPH_OK, "elph_destroy");
    printf("abi_check gpu ok: %lld CG iterations, residual %.2e\n", (long long)info.iters, info.residual);
    return 0;
}

int main(int argc, char** argv) {
    size_t k, n = sizeof(all_symbols) / sizeof(all_symbols[0]);
    for (k = 0; k < n; ++k) CHECK(all_symbols[k] != NULL, "unresolved symbol");
    CHECK(elph_version() && strlen(elph_version()) > 0, "elph_version");
    CHECK(elph_update_model(NULL) != ELPH_OK, "null handle must be rejected");
    CHECK(strlen(elph_last_error(NULL)) > 0, "null-handle error message");
    printf("abi_check ok: %zu symbols linked, version %s\n", n, elph_version());
    if (argc > 1 && strcmp(argv[1], "gpu") == 0) return gpu_part();
    return 0;
}
