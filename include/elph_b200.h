/*
 * elph_b200.h -- C ABI of libelph_b200.so, the B200 (sm_100a) engine for the
 * ElPhDynamics hot path (fermion-matrix solve + force evaluation).
 *
 * The reference (cohensbw/ElPhDynamics v1.1.3) has NO foreign-function
 * interface: its "operator API" is Julia multiple dispatch on AbstractModel
 * (src/Models.jl:65).  Each entry point below therefore cites the Julia method
 * it replaces; a Julia shim type whose methods `ccall` these symbols is shown
 * in INTEGRATION.md and shipped (unexecuted here: no Julia in this image) as
 * julia/ElPhB200.jl.
 *
 * Conventions
 *  - Every function returns an int32 status: 0 = ok, nonzero = error; the
 *    message is available from elph_last_error().  No C++ exception and no
 *    exit() crosses this boundary.  Solver non-convergence is DATA (flag 1/2,
 *    src/Models.jl:100-126), not an error.
 *  - Host vectors use the reference layout, tau-fastest:
 *        index = (site-1)*Ltau + tau          (src/Utilities.jl:12-15)
 *    i.e. a column-major (Ltau, N) Julia array.  Phonon fields the same with
 *    the phonon index in place of the site.  Device-resident state uses the
 *    engine's tau-slice-major layout [tau][site]; the `_dev` entry points take
 *    and return DEVICE pointers in that layout and are asynchronous on the
 *    handle's stream (elph_set_stream).  All other entry points take HOST
 *    pointers, copy in/out, and return after the stream has drained.
 *  - The caller owns every buffer it passes; the library never retains a host
 *    pointer past the call.  One handle = one caller thread at a time.
 *  - All randomness is injected by the caller (eta, g, R+-, Arnoldi start
 *    values); the library has no RNG on the parity path.
 *  - Index tables are int64 and may be 1-based (Julia) or 0-based; see
 *    elph_config.index_base.
 */
#ifndef ELPH_B200_H
#define ELPH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ELPH_OK 0
#define ELPH_ERR_INVALID 1   /* bad argument / configuration            */
#define ELPH_ERR_CUDA 2      /* CUDA runtime failure                    */
#define ELPH_ERR_STATE 3     /* call sequence error (e.g. KPM not set)  */
#define ELPH_ERR_UNSUPPORTED 4

#define ELPH_MODEL_HOLSTEIN 0
#define ELPH_MODEL_SSH 1

#define ELPH_LANGEVIN_EULER 1 /* update_method = 1, src/LangevinDynamics.jl:81  */
#define ELPH_LANGEVIN_RK 2    /* update_method = 2, src/LangevinDynamics.jl:162 */
#define ELPH_LANGEVIN_HEUN 3  /* update_method = 3, src/LangevinDynamics.jl:272 */

typedef struct elph_handle elph_handle;

/*
 * Flat description of a model.  Replaces the fields of HolsteinModel
 * (src/HolsteinModels.jl:22-314) / SSHModel (src/SSHModels.jl:79-314) that the
 * hot path reads, plus the solver and preconditioner parameters of
 * ConjugateGradient (src/IterativeSolvers.jl:36-57) and KPMExpansion
 * (src/KPMPreconditioners.jl:21-146).
 */
typedef struct elph_config {
    int32_t model;        /* ELPH_MODEL_HOLSTEIN | ELPH_MODEL_SSH                          */
    int32_t index_base;   /* 1 if the int64 tables below are 1-based (Julia), else 0       */
    int32_t device;       /* CUDA device ordinal, -1 = current device                      */
    int32_t reserved0;
    int64_t Ltau;         /* model.Ltau                                                     */
    int64_t Nsites;       /* model.Nsites                                                   */
    int64_t Nbonds;       /* model.Nbonds                                                   */
    int64_t Nph;          /* model.Nph (Holstein: == Nsites)                                */
    double dtau;          /* model.dtau                                                     */

    /* (2, Nbonds) column-major = Nbonds (i,j) pairs, ALREADY in checkerboard
     * order (model.neighbor_table; src/HolsteinModels.jl:505-509).  The colour
     * groups are recovered from the order. */
    const int64_t* neighbor_table;

    /* Holstein: cosht/sinht (Nbonds, checkerboard order), lambda, lambda2, mu
     * (Nsites), omega, omega4 (Nph).  src/HolsteinModels.jl:100-136. */
    const double* cosht;
    const double* sinht;
    const double* lambda;
    const double* lambda2;
    const double* mu;
    const double* omega;
    const double* omega4;

    /* SSH only (NULL for Holstein).  t (Nbonds, original bond order), alpha,
     * alpha2 (Nph), maps as in src/SSHModels.jl:146-173:
     *   checkerboard_perm[bond]      -> column in neighbor_table
     *   inv_checkerboard_perm[col]   -> bond
     *   phonon_to_bond[phonon]       -> bond
     *   bond_to_phonon[bond]         -> phonon, or (index_base-1) if none
     *   primary_field[field]         -> field, Ndof = Nph*Ltau entries, host layout */
    const double* t;
    const double* alpha;
    const double* alpha2;
    const int64_t* checkerboard_perm;
    const int64_t* inv_checkerboard_perm;
    const int64_t* phonon_to_bond;
    const int64_t* bond_to_phonon;
    const int64_t* primary_field;

    /* ConjugateGradient(tol, maxiter, kappa_max)  src/IterativeSolvers.jl:36-57 */
    double cg_tol;
    int64_t cg_maxiter;
    double cg_kappa_max;   /* 0 -> 1e12 */

    /* KPMExpansion(model, n, buf, c1, c2)  src/KPMPreconditioners.jl:101; kpm_n = 0 -> no preconditioner */
    int64_t kpm_n;
    double kpm_buf;
    double kpm_c1;
    double kpm_c2;

    /* FourierAccelerator Q and M diagonals (Nph*Ltau, host layout), already
     * filled by update_Q!/update_M! (src/FourierAcceleration.jl:149-193); may be NULL. */
    const double* fa_Q;
    const double* fa_M;
} elph_config;

/* Result of ldiv!(x, model, b[, P]) -> (iters, residual_error, flag), src/Models.jl:74-186 */
typedef struct elph_solve_info {
    int64_t iters;
    double residual;
    int32_t flag;          /* 0 ok, 1 hit maxiter, 2 false convergence */
    int32_t used_fallback; /* 1 if the unpreconditioned retry ran (src/Models.jl:129-133) */
    int64_t pcg_iters;     /* iterations of the preconditioned attempt (== iters if no fallback) */
} elph_solve_info;

/* Result of setup!(P), src/KPMPreconditioners.jl:269-321 */
typedef struct elph_kpm_info {
    int32_t active;
    int32_t recomputed;    /* coefficients were regenerated on this call (hysteresis :288) */
    double e_min, e_max;   /* Arnoldi bounds */
    double lambda_lo, lambda_hi;
    int64_t total_order;   /* sum over omega of the Chebyshev orders */
    int64_t max_order;
} elph_kpm_info;

/* ------------------------------------------------------------------ lifecycle */
const char* elph_version(void);
/* last error message of `h`, or of the failed elph_create when h == NULL */
const char* elph_last_error(const elph_handle* h);

/* HolsteinModel(...) + initialize_model! (src/HolsteinModels.jl:196,484) / SSHModel
 * (src/SSHModels.jl:216,348): uploads tables, allocates all scratch that the
 * reference keeps in model.v', v'', v''' / cg.r,p,z / KPM v1..v5. */
int32_t elph_create(const elph_config* cfg, elph_handle** out);
int32_t elph_destroy(elph_handle* h);
/* launch all work of this handle on `cuda_stream` (a cudaStream_t); NULL = legacy default */
int32_t elph_set_stream(elph_handle* h, void* cuda_stream);
int32_t elph_synchronize(elph_handle* h);
/* Page-lock a long-lived host array of the caller (a preallocated Julia Vector{Float64} that is passed to the host-buffer
 * entry points again and again: noise vectors, right-hand sides) so that its copies run at full PCIe rate instead of
 * through the driver's pageable staging path.  The reference has no counterpart (its arrays never leave the host).
 * Unregister before the array is freed.  Registering the same range twice is an error of the CUDA runtime (status != 0). */
int32_t elph_host_register(elph_handle* h, void* host_ptr, int64_t bytes);
int32_t elph_host_unregister(elph_handle* h, void* host_ptr);

/* Objects the reference constructs AFTER the model may also be configured after elph_create:
 *  - ConjugateGradient(tol, maxiter, kappa_max)            src/IterativeSolvers.jl:36-57  (0 keeps the current value)
 *  - SymmetricKPMPreconditioner(model, n, buf, c1, c2)     src/KPMPreconditioners.jl:219-235 (resets lambda_lo/hi = 0/2)
 *  - FourierAccelerator Q / M after update_Q!/update_M!    src/FourierAcceleration.jl:149-193 (either may be NULL) */
int32_t elph_set_solver(elph_handle* h, double tol, int64_t maxiter, double kappa_max);
int32_t elph_kpm_configure(elph_handle* h, int64_t n, double buf, double c1, double c2);
int32_t elph_set_fourier_acceleration(elph_handle* h, const double* Q, const double* M);

/* -------------------------------------------------------------- field / tables */
/* model.x .= x ; x = model.x (Ndof, host layout).  Does NOT call update_model!. */
int32_t elph_set_x(elph_handle* h, const double* x);
int32_t elph_get_x(elph_handle* h, double* x);
/* model.mu .= mu (Nsites) -- MuFinder writes mu between updates (src/MuFinder.jl) */
int32_t elph_set_mu(elph_handle* h, const double* mu);
/* update_model!(model): src/HolsteinModels.jl:526-549 / src/SSHModels.jl:510-562.
 * SSH: returns ELPH_ERR_STATE if equivalent fields differ (the reference error(), :549-559). */
int32_t elph_update_model(elph_handle* h);
/* copies of derived tables, host layout: Holstein expnDtauV (Ndim); SSH cosht,sinht (Ltau,Nbonds) col-major */
int32_t elph_get_expnV(elph_handle* h, double* out);
int32_t elph_get_cosh_sinh(elph_handle* h, double* cosht, double* sinht);

/* ------------------------------------------------------------------- operators */
/* mulM!(y,model,v) src/HolsteinModels.jl:569, src/SSHModels.jl:581 (y must not alias v) */
int32_t elph_mulM(elph_handle* h, const double* v, double* y);
/* mulMT!(y,model,v) src/HolsteinModels.jl:631, src/SSHModels.jl:646 */
int32_t elph_mulMT(elph_handle* h, const double* v, double* y);
/* mulMTM!(y,model,v) src/Models.jl:215 -- one fused kernel, no v' round trip */
int32_t elph_mulMTM(elph_handle* h, const double* v, double* y);
/* nrhs independent right-hand sides, each Ndim long, contiguous (GreensFunctions.jl:201-234 caller) */
int32_t elph_mulMTM_batch(elph_handle* h, int64_t nrhs, const double* v, double* y);
/* muldMdx!(dMdx,u,model,v) src/HolsteinModels.jl:691, src/SSHModels.jl:707 ; dMdx has Ndof entries */
int32_t elph_muldMdx(elph_handle* h, const double* u, const double* v, double* dMdx);

/* -------------------------------------------------------------------- solvers */
/* setup!(P) src/KPMPreconditioners.jl:269.  arnoldi_noise: 2*Nsites N(0,1) values in the order the
 * reference draws them (:859-861 then :902-904). */
int32_t elph_kpm_setup(elph_handle* h, const double* arnoldi_noise, elph_kpm_info* info);
/* ldiv!(vout,P,vin) src/KPMPreconditioners.jl:426 */
int32_t elph_kpm_apply(elph_handle* h, const double* vin, double* vout);
/* Chebyshev orders (ceil(Ltau/2) entries) and coefficients for frequency w (0-based) -- debug/parity aid */
int32_t elph_kpm_get_orders(elph_handle* h, int64_t* orders);
int32_t elph_kpm_get_coeff(elph_handle* h, int64_t w, double* re_im_interleaved);
/* solve!(x,A,b,cg[,P]) src/IterativeSolvers.jl:153,239: raw CG, x is in/out (initial guess), returns iteration count.
 * use_precond != 0 requires a prior elph_kpm_setup.  tol = 0 / maxiter = 0 -> configured defaults. */
int32_t elph_cg_solve(elph_handle* h, const double* b, double* x, int32_t use_precond, double tol,
                      int64_t maxiter, int64_t* iters, double* eps);
/* ldiv!(x,model,b,P;maxiter) src/Models.jl:74-186 incl. true residual, flags, unpreconditioned fallback.
 * x is in/out (callers zero it).  tol_scale_power: the solve runs with tol^power (HMC calc_O^-1Lambda-phi,
 * src/HMC.jl:838-842); pass 1.0 otherwise. */
int32_t elph_solve(elph_handle* h, const double* b, double* x, int32_t use_precond, double tol_power,
                   elph_solve_info* info);
/* nrhs solves on the same field in one call: the n_v measurement vectors of update!(Gr,model,P)
 * (src/GreensFunctions.jl:201-234: fill!(M^-1 R,0); ldiv! per random vector) and the two pseudofermion flavours of
 * calc_O^-1Lambda-phi! (src/HMC.jl:855-885).  B, X: nrhs consecutive N*Ltau vectors (host layout).  X is output only --
 * the initial guesses are zero, as at those call sites.  infos[k] as elph_solve's info for right-hand side k.
 * Unpreconditioned solves run simultaneously (one persistent cooperative CG kernel, a barrier counter per right-hand
 * side); with use_precond != 0 the call is a loop of elph_solve.  Iteration counts and results per right-hand side are
 * those of elph_solve on a zero initial guess. */
int32_t elph_solve_batch(elph_handle* h, int64_t nrhs, const double* B, double* X, int32_t use_precond, double tol_power,
                         elph_solve_info* infos);
/* update!(Gr,model,P) src/GreensFunctions.jl:201-234 for all n_v random vectors in one call:
 * MinvR[:,k] = (M^T M)^-1 M^T R[:,k]  (mulMT! into scratch, fill!(M^-1 r,0), ldiv!).  R is drawn by the caller
 * (randn!(model.rng, r1), in column order); setup!(P) stays a separate call (elph_kpm_setup) made before this one. */
int32_t elph_Minv_batch(elph_handle* h, int64_t nrhs, const double* R, double* MinvR, int32_t use_precond,
                        elph_solve_info* infos);

/* ------------------------------------------------------------------ transforms */
/* tau_to_omega!(vout,fft,vin) src/TimeFreqFFTs.jl:55; vout complex (re,im interleaved), Ndim entries */
int32_t elph_tau_to_omega(elph_handle* h, const double* vin, double* vout_complex);
/* omega_to_tau!(vout::real,fft,vin::complex) src/TimeFreqFFTs.jl:112 */
int32_t elph_omega_to_tau(elph_handle* h, const double* vin_complex, double* vout);
/* fourier_accelerate!(v',fa,v,power;use_mass) real->real, src/FourierAcceleration.jl:131 */
int32_t elph_fourier_accelerate(elph_handle* h, const double* v, double* vout, double power, int32_t use_mass);

/* ------------------------------------------------------------ action and force */
/* calc_Sb(model,shifted) src/PhononAction.jl:11,68 (uses the device-resident x) */
int32_t elph_Sb(elph_handle* h, int32_t shifted, double* Sb);
/* calc_dSbdx!(dSbdx,model,shifted) src/PhononAction.jl:114,189 -- ACCUMULATES into dSbdx (in/out) */
int32_t elph_dSbdx(elph_handle* h, int32_t shifted, double* dSbdx);
/* calc_dSdx!(dSdx,g,M^-1 g,model,P) src/LangevinDynamics.jl:334 with g injected.  arnoldi_noise may be NULL
 * when use_precond == 0.  Minv_g (Ndim) may be NULL. */
int32_t elph_calc_dSdx(elph_handle* h, const double* g, const double* arnoldi_noise, int32_t use_precond,
                       double* dSdx, double* Minv_g, elph_solve_info* info);

/* -------------------------------------------------------------------- dynamics */
/* evolve!(model,dyn,fa,P) src/LangevinDynamics.jl:81,162,272.  eta (Ndof), g1, g2 (Ndim; g2 unused for Euler),
 * arnoldi1/2 (2*Nsites each, NULL without preconditioner).  x stays device-resident; read it with elph_get_x.
 * iters = the reference's return value; info1/info2 (may be NULL) = both solves. */
int32_t elph_langevin_step(elph_handle* h, int32_t method, double dt, const double* eta, const double* g1,
                           const double* g2, const double* arnoldi1, const double* arnoldi2, int32_t use_precond,
                           int64_t* iters, elph_solve_info* info1, elph_solve_info* info2);

/* ------------------------------------------------------------------------- HMC */
/* HybridMonteCarlo state (v, phi+-, Lambda phi+-, O^-1 Lambda phi+-, Lambda; src/HMC.jl:20-279) lives in the handle.
 * elph_hmc_get ids: 0 v, 1 phi+, 2 phi-, 3 Lambda phi+, 4 Lambda phi-, 5 O^-1 Lambda phi+, 6 O^-1 Lambda phi-,
 * 7 Lambda, 8 dSdx of the last elph_hmc_calc_dSdx.  (For SSH the Lambda operators are no-ops: Lambda phi = M^T R.) */
int32_t elph_hmc_set_v(elph_handle* h, const double* v);
int32_t elph_hmc_get(elph_handle* h, int32_t which, double* out);
/* refresh_v!(hmc,model,fa) src/HMC.jl:648: v = alpha v + sqrt(1-alpha^2) sqrt(M^-1) R, R (Ndof) injected */
int32_t elph_hmc_refresh_v(elph_handle* h, double alpha, const double* R);
/* refresh_phi!(hmc,model) src/HMC.jl:666: phi+- = Lambda^-1 M^T R+-, returns S = (R+^2+R-^2)/2 + Sb */
int32_t elph_hmc_refresh_phi(elph_handle* h, const double* R_plus, const double* R_minus, double* S);
/* calc_O^-1 Lambda phi!(hmc,model,P,power) src/HMC.jl:820: setup!(P) (2*Nsites Arnoldi values or NULL), two solves
 * with tol^power; iters = cld(sum,2) when both converge; flag as ldiv! */
int32_t elph_hmc_calc_Oinv(elph_handle* h, int32_t use_precond, const double* arnoldi_noise, double power, int64_t* iters,
                           int32_t* flag);
/* One proposal of special_update!(model,hmc,update,P) src/SpecialUpdates.jl:97-160 (ReflectionUpdate, kind 0: x[:,i] = -x[:,i],
 * Holstein only) and :233-366 (SwapUpdate, kind 1: swap!(x[:,i], x[:,j]); Holstein: the two sites of the sampled bond,
 * SSH: two sampled phonons), entirely on the device: S0 = refresh_phi! with the injected R_plus / R_minus (Ndim each),
 * the move, update_model!, calc_O^-1 Lambda phi! at tol^2 (arnoldi_noise: 2*Nsites values or NULL), S1 = calc_S,
 * accept iff uniform < min(1, exp(-(S1-S0))) and flag == 0, else the move is undone.  i, j: 0-based phonon columns.
 * The sampling of sites / bonds and the acceptance ratio stay with the caller (the RNG lives in Julia). */
int32_t elph_hmc_special_update(elph_handle* h, int32_t kind, int64_t i, int64_t j, const double* R_plus, const double* R_minus,
                                const double* arnoldi_noise, int32_t use_precond, double uniform, int32_t* accepted, double* S0,
                                double* S1, int64_t* iters, int32_t* flag);
/* calc_H(hmc,model,fa) src/HMC.jl:698: H = S + K, S = Sf + Sb, K = v.M.v/2 (SSH: primary fields only) */
int32_t elph_hmc_calc_H(elph_handle* h, double* H, double* S, double* K);
/* fill!(dSdx,0); calc_dSfdx!(hmc,model) [+ calc_dSbdx!(dSdx,model)] src/HMC.jl:749-814 */
int32_t elph_hmc_calc_dSdx(elph_handle* h, int32_t fermion_only, double* dSdx);
/* update!(model,hmc,fa,P) src/HMC.jl:310: one whole trajectory on the device (standard leapfrog for Nb == 1,
 * multi-timestep otherwise).  Injected: R_v (Ndof), R_plus/R_minus (Ndim), arnoldi_noise ((Nt+2) x 2*Nsites values in
 * call order, NULL without preconditioner), the Metropolis uniform.  On rejection x is restored and v = -v0.
 * iters = cld(total, Nt+2) like the reference (which drops the first solve's count in the multi-timestep path, :515). */
int32_t elph_hmc_update(elph_handle* h, double dt, int64_t Nt, int64_t Nb, double alpha, const double* R_v, const double* R_plus,
                        const double* R_minus, const double* arnoldi_noise, int32_t use_precond, double uniform,
                        int32_t* accepted, double* iters, double* H0, double* H1, int32_t* flag);

/* ------------------------------------------------- Green's-function estimator */
/* EstimateGreensFunction (src/GreensFunctions.jl:23-188).  elph_greens_load keeps the random vectors R and the solutions
 * M^-1 R of update!(Gr, model, P) (:201-234; computed with elph_Minv_batch) on the device: nv vectors of Ndim doubles
 * each, host layout, vector k at offset k * Ndim.  elph_greens_setup = setup!(estimator, n1, n2) (:239-296) with 0-based
 * vector indices: the four convolve! calls (:361-414) -- antiperiodic_copy! / periodic_product! (:420-457), forward
 * transforms over (omega, k1, k2, k3), a'[w,s2,k] b'[-w,s1,-k] / V, inverse transform -- entirely on the device.  Each
 * output is a complex array (re, im interleaved) of Julia dimensions (2 Ltau, norbits, norbits, L1, L2, L3) in the
 * reference's column-major order, i.e. exactly estimator.G... after setup!; NULL skips an output.  FFTW's conventions:
 * forward unnormalised exp(-2 pi i jk/n), inverse scaled by 1/n.  Any Ltau; lattice extents up to 64 per axis. */
int32_t elph_greens_load(elph_handle* h, int64_t nv, const double* R, const double* MinvR);
int32_t elph_greens_setup(elph_handle* h, int64_t n1, int64_t n2, int64_t L1, int64_t L2, int64_t L3, int64_t norbits, double* G_D0,
                          double* G_D0_G_D0, double* G_DD_G_00, double* G_D0_G_0D);

/* --------------------------------------------------- device-resident (bench) API */
/* Device pointers, engine layout [tau][site], asynchronous on the handle's stream. */
int32_t elph_dev_mulMTM(elph_handle* h, const double* v_dev, double* y_dev);
int32_t elph_dev_mulM(elph_handle* h, const double* v_dev, double* y_dev);
int32_t elph_dev_mulMT(elph_handle* h, const double* v_dev, double* y_dev);
/* nrep independent replicas (own expnV table + own vector each, strides in doubles): the reference's only
 * scale-out is independent runs distinguished by `id` (src/ElPhDynamics.jl:90-95). */
int32_t elph_dev_mulMTM_replicas(elph_handle* h, int64_t nrep, const double* expnV_dev, int64_t expnV_stride,
                                 const double* v_dev, double* y_dev, int64_t vec_stride);
/* The same for the SSH model on periodic square lattices (config C).  A replica's operator is its (cosh, sinh)(dtau t') table
 * (update_model!, src/SSHModels.jl:510-540; sweeps :581-701): elph_dev_ssh_replica_tables evaluates it for nrep phonon
 * fields x_dev + r*x_stride (engine layout [tau][phonon]) into tab_dev + r*tab_stride, 4*Ltau*Nsites doubles per replica in
 * the engine's tile layout [tau][direction][site](cosh, sinh); elph_dev_mulMTM_replicas_ssh applies M^T M of every replica to
 * its own vector (48 B per lattice point of compulsory traffic).  Strides in doubles, pointers 16-byte aligned. */
int32_t elph_dev_ssh_replica_tables(elph_handle* h, int64_t nrep, const double* x_dev, int64_t x_stride, double* tab_dev,
                                    int64_t tab_stride);
int32_t elph_dev_mulMTM_replicas_ssh(elph_handle* h, int64_t nrep, const double* tab_dev, int64_t tab_stride,
                                     const double* v_dev, double* y_dev, int64_t vec_stride);
/* ---- tau-sharding across GPUs: one handle per rank, created with Ltau = the rank's slab length.
 * elph_set_shard tells it which global slices it owns (tau0 .. tau0+Ltau-1 of Lglob) so that the antiperiodic sign
 * lands on GLOBAL slice 0; expnV is re-homed with one halo slice on each side ([halo_lo][own...][halo_hi], own start =
 * elph_dev_ptr_expnV).  The shard entry points take pointers to the FIRST OWN slice of vectors laid out the same way:
 * the caller (one process per GPU, torch.distributed/NCCL) fills v[-1] with the left neighbour's last slice and v[Ltau]
 * (and expnV[Ltau]) with the right neighbour's first slice before the call -- one halo exchange per product.
 * mode: 0 = M v, 1 = M^T v, 2 = M^T M v.  No reference counterpart (the reference is single-process).
 * SSH model (src/SSHModels.jl:581-701): expmu has no time index; what couples to the neighbour slab is the per-slice
 * (cosh, sinh) table, which is re-homed the same way in rows of 2*Ncolumns doubles (own start = elph_dev_ptr_cosh_sinh); the
 * caller fills row Ltau with the right neighbour's first row after every update_model!.  Products, the force (elph_dev_shard_
 * muldMdx, src/SSHModels.jl:745-830) and the bosonic gradient work on SSH slabs; the peer-memory kernels are Holstein-only. */
int32_t elph_set_shard(elph_handle* h, int64_t tau0, int64_t Lglob);
int32_t elph_dev_shard_matvec(elph_handle* h, int32_t mode, const double* v_own, double* y_own);
int32_t elph_dev_shard_muldMdx(elph_handle* h, const double* u_own, const double* v_own, double* out, double scale);
/* Peer-memory CG of the tau-sharded lattice: solve!(x,A,b,cg) src/IterativeSolvers.jl:239-314 with x0 = 0 as ONE
 * persistent cooperative kernel per GPU; the halo of p and the two scalar all-reduces of every iteration travel over
 * NVLink peer memory inside the kernel (no NCCL call, no launch per iteration).  Set-up, once per handle after
 * elph_set_shard: every rank calls elph_shard_p2p_export (allocates the exchange arena, returns its 64-byte CUDA IPC
 * handle), the caller all-gathers the handles and the slab lengths in rank order (torch.distributed / MPI / files) and
 * passes them to elph_shard_p2p_open.  Then every rank of the ring calls elph_dev_shard_cg_p2p with its own slices of
 * b and x ([Ltau][Nsites], engine layout); all ranks return the same iteration count and eps.  ELPH_ERR_UNSUPPORTED if
 * the slab's time slices are not all co-resident on the GPU; ELPH_ERR_STATE if a peer never reached a barrier. */
int32_t elph_shard_p2p_export(elph_handle* h, int32_t rank, int32_t world, unsigned char* ipc_handle_out);
int32_t elph_shard_p2p_open(elph_handle* h, const unsigned char* ipc_handles, const int64_t* slab_lengths);
/* 1 when a persistent CG kernel can serve this rank's slab (every time slice co-resident on the GPU); elph_shard_p2p_open
 * succeeds either way (the halo exchange elph_dev_shard_halo needs only the arenas) */
int32_t elph_shard_cg_available(elph_handle* h, int32_t* available);
int32_t elph_dev_shard_cg_p2p(elph_handle* h, const double* b_own, double* x_own, double tol, int64_t maxiter,
                              int64_t* iters, double* eps);
int32_t elph_dev_update_model(elph_handle* h);
/* calc_dSbdx! on a slab: dSbdx_own += dSb/dx, x_own = first own slice of a halo'd copy of the field (periodic in tau) */
int32_t elph_dev_shard_dSbdx(elph_handle* h, double* dSbdx_own, const double* x_own, int32_t shifted);
/* Halo exchange of one halo'd slab vector through peer memory (after elph_shard_p2p_open): fills the slice before v_own with the
 * left neighbour's last own slice and the slice after the own ones with the right neighbour's first own slice -- the exchange
 * every product of the sharded lattice needs (mulM!: tau-1, src/HolsteinModels.jl:594-601; mulMT!: tau+1, :671-677).  One
 * kernel launch on the handle's stream, no host synchronisation; every rank of the ring must make the same call. */
int32_t elph_dev_shard_halo(elph_handle* h, double* v_own);
/* elph_dev_shard_halo on v followed by elph_dev_shard_matvec: one product of the sharded lattice (mode 0 = M, 1 = M^T, 2 = M^T M)
 * with its halo exchange in ONE call (two launches, nothing returns to the host in between) */
int32_t elph_dev_shard_matvec_halo(elph_handle* h, int32_t mode, double* v_own, double* y_own);
/* fourier_accelerate! for `ncols` columns in [k][col] layout with an explicit diagonal (same layout): after the
 * all-to-all transpose of the tau-sharded driver a rank holds all Ltau slices of a subset of the sites.  The handle's
 * Ltau must be the GLOBAL time extent (the driver keeps a 1-site handle just for this plan). */
int32_t elph_dev_fourier_accelerate_cols(elph_handle* h, const double* vin_dev, double* vout_dev, int64_t ncols,
                                         const double* diag_dev, double power);
/* ---- KPM preconditioner of a tau-sharded lattice (reference: ldiv!(v', P, v), src/KPMPreconditioners.jl:426-481, on one
 * process).  The apply is three stages on three shardings with an all-to-all between them: tau-sharded ->
 * site-sharded (tau_to_omega! on all slices of a subset of the sites, src/TimeFreqFFTs.jl:31-75) -> omega-sharded (the
 * Chebyshev recurrences of this rank's frequencies on ALL sites, :606-679) -> back.  These entries act on an auxiliary handle
 * created for the GLOBAL lattice (Nsites, global Ltau, kpm_n > 0): it owns the FFT plan, the coefficients and the chain kernels.
 *   elph_dev_tau_to_omega_cols / elph_dev_omega_to_tau_cols: [tau][col] real <-> [omega][col] complex (interleaved re, im)
 *   elph_dev_kpm_setup_bar: setup!(P) (:269-321) with the tau-mean supplied by the caller (local sums + all-reduce replace
 *     update_A!): Holstein: the mean of expnV, [Nsites] (:332-350); SSH: the mean of the (cosh, sinh) pairs, [Ncolumns][2] in the
 *     order of elph_dev_ptr_cosh_sinh (:355-381).  Arnoldi bounds, hysteresis and coefficients as elph_kpm_setup.
 *   elph_kpm_set_omega_subset: this handle's chain kernels run the frequencies w = first, first + stride, ... < cld(Ltau, 2)
 *   elph_dev_kpm_chains: nu_out[w], nu_out[Ltau-1-w] = conj for every frequency of the subset from nu_in[w]; both buffers are
 *     [Ltau][Nsites] complex indexed by the GLOBAL frequency, rows of other frequencies are left untouched. */
int32_t elph_dev_tau_to_omega_cols(elph_handle* h, const double* vin_dev, double* nu_dev, int64_t ncols);
int32_t elph_dev_omega_to_tau_cols(elph_handle* h, const double* nu_dev, double* vout_dev, int64_t ncols);
int32_t elph_dev_kpm_setup_bar(elph_handle* h, const double* bar_dev, const double* arnoldi_noise, elph_kpm_info* info);
int32_t elph_kpm_set_omega_subset(elph_handle* h, int64_t first, int64_t stride);
int32_t elph_dev_kpm_chains(elph_handle* h, const double* nu_in_dev, double* nu_out_dev);
/* The same application as ONE call per rank with the transposes through peer memory instead of all-to-alls (csrc/kpm_shard.cu):
 * every rank exports an arena holding the output of each of its stages (CUDA IPC handle, 64 bytes), all ranks open all arenas
 * (handles in rank order; slab_starts[q] = first global slice of rank q's slab, near-equal contiguous slabs; the site blocks are
 * the near-equal contiguous split of Nsites), and elph_dev_kpm_shard_apply runs  copy-in | forward FFT | gather + chains | inverse
 * FFT | gather  with the loads of each stage pulling from the producers' arenas over NVLink and one 32-thread cross-GPU barrier
 * kernel between stages.  r_own / z_own: [Lloc][Nsites] device vectors of this rank's slab; every rank must make the same calls
 * in the same order.  Requires elph_kpm_set_omega_subset(rank, world) and an active preconditioner (the caller copies r to z
 * otherwise, src/KPMPreconditioners.jl:475-478).  A barrier that is not reached within 2 s raises a failure flag, reported by
 * the next call and by elph_kpm_shard_check (which synchronises the stream). */
int32_t elph_kpm_shard_export(elph_handle* h, int32_t rank, int32_t world, int64_t tau0, int64_t lloc, unsigned char* ipc_handle_out);
int32_t elph_kpm_shard_open(elph_handle* h, const unsigned char* ipc_handles, const int64_t* slab_starts);
int32_t elph_dev_kpm_shard_apply(elph_handle* h, const double* r_own_dev, double* z_own_dev);
int32_t elph_kpm_shard_check(elph_handle* h);
/* BLAS-1 on device pointers for the sharded solver: out = a X + b Y + c Z (Y, Z may be NULL); out_dev[0] = a.b */
int32_t elph_dev_lincomb(elph_handle* h, double* out, double a, const double* X, double b, const double* Y, double c,
                         const double* Z, int64_t n);
int32_t elph_dev_dot(elph_handle* h, const double* a, const double* b, int64_t n, double* out_dev);
/* host layout (N x Ltau, tau fastest) <-> engine layout (Ltau x N) on the device */
int32_t elph_dev_to_engine_layout(elph_handle* h, const double* host_layout_dev, double* engine_dev, int64_t ncols);
int32_t elph_dev_from_engine_layout(elph_handle* h, const double* engine_dev, double* host_layout_dev, int64_t ncols);
/* device pointers to resident state (engine layout) */
int32_t elph_dev_ptr_x(elph_handle* h, double** x_dev);
int32_t elph_dev_ptr_expnV(elph_handle* h, double** expnV_dev);
int32_t elph_dev_ptr_cosh_sinh(elph_handle* h, double** cosh_sinh_dev);   /* Holstein [Ncolumns][2]; SSH [Ltau][Ncolumns][2] */
/* CG on device pointers; asynchronous until the result scalars are read (blocks). */
int32_t elph_dev_cg_solve(elph_handle* h, const double* b_dev, double* x_dev, int32_t use_precond, double tol,
                          int64_t maxiter, int64_t* iters, double* eps);
/* elph_solve_batch on device buffers (engine layout): right-hand side k at b_dev + k*N*Ltau, solution at x_dev + k*N*Ltau */
int32_t elph_dev_solve_batch(elph_handle* h, int64_t nrhs, const double* b_dev, double* x_dev, int32_t use_precond,
                             double tol_power, elph_solve_info* infos);
/* ldiv!(vout,P,vin) and fourier_accelerate! on device pointers (engine layout), asynchronous */
int32_t elph_dev_kpm_apply(elph_handle* h, const double* vin_dev, double* vout_dev);
int32_t elph_dev_fourier_accelerate(elph_handle* h, const double* vin_dev, double* vout_dev, double power, int32_t use_mass);
/* number of kernels this handle has launched since creation (bench `gpu_launches`) */
int64_t elph_launch_count(const elph_handle* h);
/* tuning knob: tau-slices per CTA for the fused matvec kernels (0 = auto) */
int32_t elph_set_chunk(elph_handle* h, int32_t slices_per_cta);
/* tuning / test knobs: key 0 = slices per CTA, 1 = disable the register/shuffle square-lattice kernel (use the
 * generic shared-memory kernel), 2 = rows per warp of the square kernel (0 = auto), 3 = CUDA-graph replay of CG
 * iteration blocks (default 1), 4 = KPM apply on 2-CTA clusters, one CTA per real/imaginary chain (default 1),
 * 5 = unpreconditioned CG as one cooperative persistent kernel when all time slices are co-resident (default 1),
 * 6 = preconditioned CG with the vector updates fused into the FFT kernels of the KPM apply (default 1),
 * 7 = single-reduction form of the persistent unpreconditioned CG (one grid barrier per iteration; Holstein on square
 *     lattices; same iterates to rounding, iteration counts within +-2 of the two-reduction loop; -1 = auto (default: on for
 *     32-wide lattices, where it is measured faster), 0 = off, 1 = on),
 * 8 = replicas per stage of the H2D | kernels | D2H pipeline behind elph_mulMTM_batch (1..8, default 8),
 * 9 = multi-timestep HMC: the Nb inner (bosonic) steps of an outer step in one kernel (default 1),
 * 10 = pipelined form of the persistent unpreconditioned CG (csrc/cg_pipe.cu: the all-reduce of an iteration overlaps the
 *      next product, a time slice may be split over several SMs; periodic square lattices; iteration counts within +-2 of
 *      the reference loop; -1 = auto (default: on unless key 7 forces one of the older forms), 0 = off, 1 = on),
 * 11 = CTAs per time slice of the pipelined kernel (0 = auto, else 1, 2, 4, 8),
 * 12 = per-phase cycle counters of the pipelined kernel (development aid), 13 = force one variant of the pipelined kernel
 *      (0 = auto), 14 = time slices per CTA of its multi-slice variants (0 = the smallest number that makes the slab co-resident),
 * 15 = cluster barrier flavour of the pipelined kernel (development aid; default = full cluster barrier),
 * 16 = KPM chains with the sweeps in tanh form and the constants folded (default 1; differs from the plain form by rounding),
 * 17 = the KPM-preconditioned solve (solve!(x,A,b,cg,P), src/IterativeSolvers.jl:153-234) as ONE persistent cooperative kernel
 *      where served (csrc/pcg_fused.cu: Holstein on 32-wide square lattices; default 1), 18 = its tau-FFTs at half length for
 *      even Ltau (default 1), 20 = its CTA count (0 = one per SM; SMs / K when K chains share one GPU),
 * 19 = Arnoldi eigenvalue bounds of setup!(P) (src/KPMPreconditioners.jl:845-942) on the device (default 1; 0 = host loops),
 * 21 = register-tile kernels for the honeycomb lattice 32 cells wide (default 1),
 * 22 = tau-sharded M^T M with the halo exchange inside the product kernel (default 1; 0 = exchange kernel + product kernel),
 * 23 = elph_langevin_step: eta and g2 travel host-to-device on a second stream during the first solve (default 1),
 * 24 = fused M^T M on square lattices with the sweeps in tanh form (default 1),
 * 25 = speculative setup!(P) in the force evaluation (calc_dSfdx!, src/LangevinDynamics.jl:350-384; default 1): the solve is queued
 *      right behind update_A! with the polynomials of the previous set-up -- which the hysteresis of setup! keeps unless the
 *      spectral window moved by more than `buf` (src/KPMPreconditioners.jl:296-309) -- while the Arnoldi kernel runs beside it and a
 *      host thread reduces its result; the solve is repeated when the set-up does change the polynomials or the active flag, so
 *      the results are those of the reference order (0 = set-up strictly before the solve),
 * 26 = KPM chains on 64-wide lattices with every frequency on an 8-CTA cluster, (re | im) x 4 row strips with the strip edges through
 *      distributed shared memory (default 0: measured no faster than the 2-CTA clusters, see csrc/kpm_square.cu) */
int32_t elph_set_tuning(elph_handle* h, int32_t key, int32_t value);
/* read-back of a tuning key; key 100 = which kernel served the last unpreconditioned persistent solve: 0 = none yet /
 * other kernels, else variant * 100 + CTAs per slice * 10 + warps per CTA of the pipelined kernel; key 101 = its time slices per CTA */
int32_t elph_get_tuning(elph_handle* h, int32_t key, int32_t* value);
/* which kernel family serves the fused M^T M product of this model, and the number of bond colours found */
int32_t elph_get_kernel_info(elph_handle* h, int32_t* square_kernel, int32_t* ngroups);

#ifdef __cplusplus
}
#endif
#endif /* ELPH_B200_H */
