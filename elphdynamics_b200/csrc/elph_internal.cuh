// Internal declarations shared by the translation units of libelph_b200.so.
// Nothing here is part of the C ABI (see include/elph_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <complex>
#include <future>
#include <set>
#include <string>
#include <vector>

#include "../../include/elph_b200.h"

typedef double2 cplx;  // (re, im)

// ----------------------------------------------------------------------------
// error plumbing: no exception crosses the ABI; every entry point is wrapped in
// ELPH_TRY { ... } ELPH_CATCH(h)
// ----------------------------------------------------------------------------
struct elph_error {
    int32_t code;
    std::string msg;
};

std::string& elph_global_error();

#define ELPH_CUDA(call)                                                                        \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            throw elph_error{ELPH_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)}; \
    } while (0)

#define ELPH_REQUIRE(cond, code, message)           \
    do {                                            \
        if (!(cond)) throw elph_error{(code), (message)}; \
    } while (0)

#define ELPH_TRY try
#define ELPH_CATCH(h)                                                   \
    catch (const elph_error& e) {                                       \
        if (h) (h)->err = e.msg; else elph_global_error() = e.msg;      \
        return e.code;                                                  \
    } catch (const std::exception& e) {                                 \
        if (h) (h)->err = e.what(); else elph_global_error() = e.what(); \
        return ELPH_ERR_INVALID;                                        \
    } catch (...) {                                                     \
        if (h) (h)->err = "unknown error"; else elph_global_error() = "unknown error"; \
        return ELPH_ERR_INVALID;                                        \
    }

// ----------------------------------------------------------------------------
// device-side CG scalar block (lives in device memory; mirrored to pinned host)
// ----------------------------------------------------------------------------
struct CgScalars {
    double rdotz;      // r.z (or r.r) of the current iterate
    double pAp;        // p.Ap of the current iteration
    double normb;      // |b|
    double eps0;       // initial relative residual
    double eps;        // current relative residual
    double kappa_min;  // running lower bound of the condition number
    double alpha;
    double beta;
    double tol;
    double kappa_max;
    long long iter;      // completed iterations
    long long maxiter;
    int done;            // latch: 1 once the stop rule fired (later launches are no-ops)
    int pad;
};

// ----------------------------------------------------------------------------
// KPM host-side state (coefficients, hysteresis), src/KPMPreconditioners.jl:21-146
// ----------------------------------------------------------------------------
struct KpmState {
    bool configured = false;
    bool ever_setup = false;
    bool active = true;
    int n = 0;
    double buf = 0.05, c1 = 1.0, c2 = 1.0;
    double lam_lo = 0.0, lam_hi = 2.0, lam_avg = 1.0, lam_mag = 1.0;
    double e_min = 0.0, e_max = 0.0;
    int Lo2 = 0;
    std::vector<double> phis;
    std::vector<int> order;                 // per omega
    std::vector<int> coeff_off;             // prefix offsets into coeff
    std::vector<std::complex<double>> coeff;  // concatenated c_m per omega
    std::vector<int> schedule;              // omegas sorted by order, longest first
    cudaStream_t spec_stream = nullptr;     // speculative set-up: Arnoldi kernel + read-back beside the solve
    cudaEvent_t spec_fork = nullptr, spec_done = nullptr;
    bool spec_stale = false;                // the last set-up changed the polynomials / the active flag
    std::future<std::pair<double, double>> spec_future;   // (e_max, 1/e_min) of the set-up in flight
    int nsched = 0;                         // frequencies the chain kernels run (= Lo2 unless an omega subset is set)
    int sub_first = 0, sub_stride = 1;      // omega-sharded apply (sharded.py): this handle runs w = first, first+stride, ...
    // host copies of the tau-averaged operator (for the Arnoldi iteration)
    std::vector<double> eVbar, cbar, sbar;
    // device
    double* d_eVbar = nullptr;    // [N]
    double2* d_csbar = nullptr;   // [Nb] (c,s)
    double2* d_csbar_tile = nullptr;   // SSH on square lattices: the same in the tile layout [direction][site]
    cplx* d_coeff = nullptr;      // concatenated coefficients
    int* d_order = nullptr;       // [Lo2]
    int* d_coeff_off = nullptr;   // [Lo2]
    int* d_schedule = nullptr;    // [Lo2]
    size_t d_coeff_cap = 0;
    cplx* d_nu = nullptr;         // [L][N] complex work vector (frequency space)
    double* d_noise = nullptr;    // [2][N] Arnoldi start vectors (device Arnoldi)
    double* d_hm = nullptr;       // [2][(n+1) n + 1] Hessenberg matrices + completed steps
    double* h_hm = nullptr;       // page-locked copy
    double* d_Q = nullptr;        // Krylov bases when they do not fit in shared memory
};

// HybridMonteCarlo work vectors (src/HMC.jl:20-279), engine layout, allocated on first use
struct HmcState {
    bool init = false;
    double *v = nullptr, *v0 = nullptr, *x0 = nullptr, *dS = nullptr, *y = nullptr, *Q = nullptr;   // Ndof
    double *Lam = nullptr, *Rp = nullptr, *Rm = nullptr, *phip = nullptr, *phim = nullptr;        // Ndim
    double *Lphip = nullptr, *Lphim = nullptr, *Op = nullptr, *Om = nullptr, *u = nullptr;        // Ndim
};

// a block of CG iterations captured as a CUDA graph (cg.cu); the key fields decide whether it can be replayed
struct CgGraph {
    cudaGraphExec_t exec = nullptr;
    const double* x = nullptr;
    bool precond = false;
    int64_t kpm_version = 0;
    int chunk = 0;
    bool sq_disable = false;
    cudaStream_t stream = nullptr;
    int64_t nlaunch = 0;
};

// copy/compute pipeline of the batched host-buffer entry points (api.cu: host_matvec_pipelined)
struct HostPipe {
    bool init = false;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    double* buf[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
};

struct elph_handle {
    std::string err;
    HmcState hmc;
    struct { int cap = 0; double* V = nullptr; double* partial = nullptr; } g1r;   // cg1r_generic work space (cg_persistent.cu)
    void* greens = nullptr;    // GreensState of greens.cu (Green's-function convolutions), allocated on first use
    HostPipe pipe;
    std::vector<CgGraph> cg_graphs;
    int64_t kpm_version = 0;   // bumped whenever the KPM kernels' launch parameters change
    bool use_graphs = true;
    bool trace = false;        // ELPH_TRACE=1: phase timings of the dynamics entry points on stderr (synchronises the stream)
    double trace_t0 = 0.0;
    bool kpm_split = true;     // KPM apply: one 2-CTA cluster per frequency (re / im chains), see kpm_square.cu
    bool kpm_exclusive = true; // KPM apply: one chain CTA per SM (shared-memory request padded)
    bool mtm_tanh = true;        // fused M^T M on square lattices: sweeps in tanh form (tuning key 24)
    bool overlap_uploads = true; // elph_langevin_step: eta and g2 travel on a second stream during the first solve (tuning key 23)
    cudaStream_t upload_stream = nullptr;
    cudaEvent_t upload_event = nullptr, upload_fence = nullptr;
    bool upload_pending = false;
    bool halo_fused = true;      // sharded M^T M: halo exchange inside the product kernel (tuning key 22)
    bool hc_tiles = true;        // honeycomb lattices: register-tile kernels (tuning key 21)
    int pcg_grid = 0;            // fused PCG: CTAs of the persistent kernel (0 = one per SM); tuning key 20
    bool kpm_dev_arnoldi = true; // KPM set-up: Arnoldi eigenvalue bounds on the device (tuning key 19)
    bool kpm_wide = false;       // KPM apply on 64-wide lattices: 8-CTA clusters, (re | im) x 4 row strips per frequency (key 26;
                                 // measured no faster than the 2-CTA kernel: the strip-edge round trip through DSMEM costs what it saves)
    bool kpm_speculate = true;   // force evaluation: solve with the previous polynomials while the Arnoldi bounds are computed (key 25)
    bool spec_running = false;   // a speculative set-up is in flight: the one-kernel PCG leaves two SMs to the Arnoldi kernel
    bool pcg_half_fft = true;   // fused PCG: tau-FFTs at length L/2 for even L (tuning key 18)
    bool pcg_persistent = true; // KPM-preconditioned CG as one persistent kernel where served (pcg_fused.cu, tuning key 17)
    bool kpm_fast = true;      // KPM apply: sweeps in tanh form with folded constants (tuning key 16)
    bool use_persistent = true;  // unpreconditioned CG as one cooperative persistent kernel (cg_persistent.cu)
    int cg_single_reduction = -1;  // unpreconditioned CG, Holstein square: one barrier per iteration (cg_p2p.cu); -1 = auto
                                   // (on where measured faster: 32-wide lattices), 0 = off, 1 = on
    bool hmc_fused_inner = true; // multi-timestep HMC: the Nb inner steps of an outer step in one kernel (fft.cu)
    int pipe_chunk = 8;          // replicas per stage of the host-buffer pipeline (elph_mulMTM_batch); 8 measured best
    bool pcg_fuse = true;        // preconditioned CG: vector updates fused into the FFT kernels of the KPM apply
    unsigned int* d_bar = nullptr;  // grid-barrier arrival counter of the persistent CG
    bool own_stream = false;
    std::set<const void*> smem_enabled;  // kernels that already have the opt-in shared-memory attribute
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;

    int model = ELPH_MODEL_HOLSTEIN;
    int L = 0, N = 0, Nb = 0, Nph = 0;
    int64_t Ndim = 0, Ndof = 0;
    double dtau = 0.0;
    int ngroups = 0;
    int max_group = 0;
    std::vector<int> goff_host;
    std::vector<int2> bonds_host;

    // device tables
    int2* d_bonds = nullptr;      // [Nb] 0-based (i,j), checkerboard order
    int* d_goff = nullptr;        // [ngroups+1]
    double2* d_cs = nullptr;      // Holstein: [Nb] static (cosh,sinh); SSH: [L][Nb] per-tau table
    double* d_lam = nullptr;      // [N]
    double* d_lam2 = nullptr;     // [N]
    double* d_mu = nullptr;       // [N]
    double* d_omega = nullptr;    // [Nph]
    double* d_omega4 = nullptr;   // [Nph]
    double* d_x = nullptr;        // [L][Nph]
    double* d_D = nullptr;        // Holstein: expnV [L][N]; SSH: expmu [N]
    double* d_Q = nullptr;        // [L(k)][Nph]  fourier acceleration diagonals, engine layout
    double* d_Mass = nullptr;     // [L(k)][Nph]
    bool have_Q = false, have_M = false;

    // SSH maps (device, 0-based)
    double* d_t = nullptr;        // [Nb] original bond order
    double* d_alpha = nullptr;    // [Nph]
    double* d_alpha2 = nullptr;   // [Nph]
    int* d_ph_col = nullptr;      // [Nph] phonon -> neighbor_table column
    int* d_col_ph = nullptr;      // [Nb] column -> phonon or -1
    int* d_col_bond = nullptr;    // [Nb] column -> original bond (inv_checkerboard_perm)
    int* d_primary_ph = nullptr;  // [Nph] phonon -> primary phonon (tau-independent part of primary_field)
    std::vector<int> primary_ph_host;
    int* d_grp_start = nullptr;   // [Nph+1] CSR over primary phonons: phonons sharing that primary ...
    int* d_grp_members = nullptr; // [Nph]   ... listed in neighbour-table column order (the reference's accumulation order)
    double* d_tprime = nullptr;   // [L][Nb] modulated hopping t' per column (kept for parity getters)

    // solver configuration
    double cg_tol = 1e-5;
    int64_t cg_maxiter = 0;
    double cg_kappa_max = 1e12;

    // scratch vectors (engine layout), what the reference keeps in model.v',v'',v''' / cg.r,p,z / dyn.*
    double* d_va = nullptr;       // staging for host-buffer entry points
    double* d_vb = nullptr;
    double* d_vc = nullptr;
    double* d_b = nullptr;        // right-hand side M^T g (model.v'')
    double* d_res = nullptr;      // true-residual scratch (model.v''')
    double* d_r = nullptr;        // cg.r
    double* d_p[2] = {nullptr, nullptr};  // cg.p, double-buffered
    double* d_z = nullptr;        // cg.z
    double* d_partial = nullptr;  // per-CTA partial sums
    int partial_cap = 0;
    unsigned int* d_ticket = nullptr;
    CgScalars* d_cg = nullptr;
    CgScalars* h_cg = nullptr;    // pinned
    double* h_scal = nullptr;     // pinned scalars (8 doubles)
    double* d_scal = nullptr;

    // dynamics scratch (Ndof each)
    double* d_dSdx = nullptr;
    double* d_dSdx2 = nullptr;
    double* d_eta = nullptr;
    double* d_dx = nullptr;
    double* d_g = nullptr;        // Ndim
    double* d_g2 = nullptr;       // Ndim
    double* d_stage[4] = {nullptr, nullptr, nullptr, nullptr};  // grow-only staging (host layout) for host-buffer calls
    size_t stage_cap[4] = {0, 0, 0, 0};
    double* d_Minv = nullptr;     // Ndim
    double* d_tmp = nullptr;      // Ndof

    // FFT
    cplx* d_twiddle = nullptr;    // exp(-2 pi i k / L), k = 0..L-1
    cplx* d_theta = nullptr;      // exp(-i pi tau / L)
    std::vector<int> fft_radices;

    KpmState kpm;
    cplx* d_nu2 = nullptr;          // second frequency-space work vector [L][N]
    bool kpm_skip_enabled = false;  // inside the CG loop the KPM kernels honour the convergence latch

    // register/shuffle kernel for periodic square lattices (mtm_square.cu)
    struct {
        bool enabled = false;
        int Lx = 0, Ly = 0;
        double c[4] = {0, 0, 0, 0}, s[4] = {0, 0, 0, 0};
    } sq;
    // register/shuffle kernels for the periodic honeycomb lattice 32 cells wide (hc tiles of square_tiles.cuh): config D
    struct {
        bool enabled = false;
        int L1 = 0, L2 = 0;
        double c[3] = {0, 0, 0}, s[3] = {0, 0, 0};
    } hc;
    // SSH on the periodic square lattice (ssh_square.cu): second copy of the per-(tau,bond) table in the
    // register-tile layout [tau][dir][site] (dir 0 = +x bond leaving `site`, 1 = +y bond), written by update_model
    struct {
        bool enabled = false;
        int Lx = 0, Ly = 0;
        int* d_slot = nullptr;        // [Nb] column -> dir*N + origin site
        double2* d_tab = nullptr;     // [L][2][N] (cosh, sinh)
    } ssq;
    // batched solves (elph_solve_batch_device): grow-only buffers for `cap` right-hand sides
    struct {
        int cap = 0;
        double* x = nullptr;
        double* r = nullptr;
        double* p0 = nullptr;
        double* p1 = nullptr;
        double* partial = nullptr;
        unsigned int* bar = nullptr;
        CgScalars* S = nullptr;
        std::vector<CgScalars> hS;
    } batch;
    // tau-sharding (multi-GPU): this handle owns global slices [shard_tau0, shard_tau0 + L) of shard_Lglob
    bool sharded = false;
    int shard_tau0 = 0, shard_Lglob = 0;
    // peer-memory CG (cg_p2p.cu): arena exported over CUDA IPC, the other ranks' arenas opened in rank order
    struct {
        void* arena = nullptr;
        std::vector<void*> peer;
        std::vector<int64_t> peer_L;
        int rank = 0, world = 1, Lmax = 0;
        unsigned int seq = 0;       // sequence number of the last cross-GPU barrier executed
        bool opened = false;
        bool failed = false;        // a solve timed out: tags on the peers are undefined until the arenas are re-opened
        // pipelined CG (cg_pipe.cu): its region of the same arena, own tag counter
        size_t pipe_off = 0;
        unsigned int hx_seq = 0;    // halo exchanges of the products executed (tag of the last one)
        unsigned int pipe_seq = 0;
        bool pipe_failed = false;
    } p2p;
    // KPM preconditioner of a tau-sharded lattice with the transposes through peer memory (kpm_shard.cu); lives on the auxiliary
    // handle of the GLOBAL lattice
    struct {
        void* arena = nullptr;
        size_t bytes = 0, offA = 0, offB = 0, offC = 0, offD = 0;
        std::vector<void*> peer;
        std::vector<int> tau0s, s0s;     // [world + 1] first slice / first site of every rank
        int rank = 0, world = 1, tau0 = 0, lloc = 0;
        unsigned long long seq = 0;      // number of the last cross-GPU barrier issued
        bool opened = false;
        cplx* nu_in = nullptr;           // [L][N] gathered input rows of this rank's frequencies
        unsigned int* h_fail = nullptr;  // pinned + mapped: a barrier timed out
        unsigned int* d_fail = nullptr;
    } kshard;
    unsigned int* h_hx_flag = nullptr;   // pinned: failure flag of the peer-memory halo exchange
    int cg_pipeline = -1;          // unpreconditioned CG on square lattices: pipelined persistent kernel (cg_pipe.cu); -1 = auto, 0 = off
    int pipe_ys = 0;               // tuning: CTAs per time slice of the pipelined kernel (0 = automatic)
    int pipe_sync_mode = 0;        // tuning key 15: how the edge exchange across the CTAs of a slice is synchronised (see cg_pipe.cu)
    int pipe_spc = 0;              // tuning key 14: time slices per CTA of the multi-slice variants (0 = smallest that fits)
    int pipe_last_spc = 1;
    int pipe_variant = 0;          // tuning key 13: force one variant of the pipelined kernel (0 = automatic)
    bool pipe_prof = false;        // tuning key 12: per-phase cycle counters of the pipelined kernel (development aid)
    unsigned long long* pipe_prof_buf = nullptr;   // [8192][8]
    int pipe_last_variant = 0;     // variant * 100 + ys * 10 + warps of the last pipelined solve (diagnostics)
    double* d_D_alloc = nullptr;   // sharded: d_D points one slice into this allocation (halo slices around it)
    double2* d_cs_alloc = nullptr; // sharded SSH: d_cs points one slice into this allocation ([halo][own ...][halo] rows of Nb)
    double* d_x_alloc = nullptr;   // sharded: same for the phonon field (the force needs no x halo; kept symmetric)
    bool sq_disable = false;
    int sq_py = 0;
    int chunk_override = 0;
    unsigned smem_attr_mask = 0;  // which matvec kernel instances already have the opt-in smem attribute
    int threads = 256;
};

// ----------------------------------------------------------------------------
// launch helpers (matvec.cu)
// ----------------------------------------------------------------------------
enum MatvecMode { MODE_M = 0, MODE_MT = 1, MODE_MTM = 2 };

// the halo exchange of a sharded product done inside the product kernel (mtm_square.cu, HALO): arena rows of this GPU and of its
// two neighbours, the tag of this exchange and the failure flag (filled by elph_shard_halo_args, cg_pipe.cu)
struct HaloArgs {
    bool enabled = false;
    unsigned long long* mine = nullptr;
    unsigned long long* left = nullptr;
    unsigned long long* right = nullptr;
    unsigned int* fail = nullptr;
    double* v_out = nullptr;
    unsigned int tag = 0;
};

struct MatvecArgs {
    HaloArgs halo;
    const double* v = nullptr;
    double* y = nullptr;
    const double* D = nullptr;   // nullptr -> handle's table
    int64_t nbatch = 1;
    int64_t v_stride = 0, y_stride = 0, D_stride = 0;
    const double2* ssh_tab = nullptr;   // SSH replicas: (cosh, sinh) tables in the tile layout of ssh_square.cu, one per replica
    int64_t ssh_tab_stride = 0;         // in double2 elements
    bool open = false;              // tau-sharded slab: v (and D) carry one halo slice before and after the own slices
    double* partial_dot = nullptr;  // if set: per-CTA partial sums of dot(v, y); returns count via *npartial
    int* npartial = nullptr;
    // CG fusion (see matvec.cu): v := cg_pr + S->beta * cg_pold on the fly, stored to cg_pnew
    const double* cg_pr = nullptr;
    const double* cg_pold = nullptr;
    double* cg_pnew = nullptr;
    CgScalars* cg_S = nullptr;
    unsigned int* cg_ticket = nullptr;
};

void elph_launch_matvec(elph_handle* h, MatvecMode mode, const MatvecArgs& a);
void elph_launch_update_model(elph_handle* h);
bool elph_pcg_fused(elph_handle* h, double* x_dev, double* z_dev);   // pcg_fused.cu
void elph_launch_ssh_replica_tables(elph_handle* h, int64_t nrep, const double* x_dev, int64_t x_stride, double2* tab_dev, int64_t tab_stride);
void elph_detect_square(elph_handle* h, const std::vector<double2>& cs);
void elph_detect_honeycomb(elph_handle* h, const std::vector<double2>& cs);
bool elph_launch_mtm_square(elph_handle* h, const MatvecArgs& a);
bool elph_match_square(const elph_handle* h, int* Lx, int* Ly, std::vector<int>* slot);
void elph_detect_ssh_square(elph_handle* h);                          // ssh_square.cu
bool elph_launch_ssh_square(elph_handle* h, const MatvecArgs& a);     // ssh_square.cu
void elph_launch_transpose(elph_handle* h, const double* in, double* out, int rows, int cols, int64_t nbatch);
void elph_launch_transpose_c(elph_handle* h, const cplx* in, cplx* out, int rows, int cols);
// host layout (ncols rows of length L) -> engine layout [L][ncols]
inline void elph_to_engine(elph_handle* h, const double* host_layout_dev, double* engine_dev, int ncols, int64_t nbatch = 1) {
    elph_launch_transpose(h, host_layout_dev, engine_dev, ncols, h->L, nbatch);
}
inline void elph_from_engine(elph_handle* h, const double* engine_dev, double* host_layout_dev, int ncols, int64_t nbatch = 1) {
    elph_launch_transpose(h, engine_dev, host_layout_dev, h->L, ncols, nbatch);
}

// cg.cu
void elph_cg_device(elph_handle* h, const double* b_dev, double* x_dev, bool use_precond, double tol, int64_t maxiter,
                    int64_t* iters, double* eps);
void elph_solve_device(elph_handle* h, const double* b_dev, double* x_dev, bool use_precond, double tol_power,
                       elph_solve_info* info);
bool elph_cg_persistent(elph_handle* h, double* x_dev);   // cg_persistent.cu
#ifdef __CUDACC__
// dSb/dx at (tau, i): dtau w^2 x + 4 dtau w4 x^3 - (x(tau+1)+x(tau-1)-2x)/dtau [- dtau lam * shifted]
// (src/PhononAction.jl:114-233); x: [L][ncols], periodic in tau.  Shared by force.cu and the fused HMC inner loop (fft.cu).
__device__ __forceinline__ double dSb_term(const double* __restrict__ x, int tau, int i, int ncols, int L, double dtau, double w,
                                           double w4, double lam_shift) {
    const int tp = (tau + 1 == L) ? 0 : tau + 1;
    const int tm = (tau == 0) ? L - 1 : tau - 1;
    const double xt = x[(size_t)tau * ncols + i];
    double d = dtau * w * w * xt - lam_shift;
    d += dtau * 4.0 * w4 * xt * xt * xt;
    d -= (x[(size_t)tp * ncols + i] + x[(size_t)tm * ncols + i] - 2.0 * xt) / dtau;
    return d;
}
#endif

// ELPH_TRACE=1 development aid: wall time since the previous mark, after draining the stream
void elph_trace_mark(elph_handle* h, const char* label);
bool elph_hmc_inner_dev(elph_handle* h, double* x, double* v, double dtp, int Nb);   // fft.cu
// greens.cu
void elph_greens_free(elph_handle* h);
void elph_greens_load_impl(elph_handle* h, int nv, const double* R, const double* MinvR);
void elph_greens_setup_impl(elph_handle* h, int n1, int n2, int L1, int L2, int L3, int ns, double* const out[4]);
// cg_p2p.cu
void elph_shard_p2p_export_impl(elph_handle* h, int rank, int world, unsigned char* handle_out);
void elph_shard_p2p_open_impl(elph_handle* h, const unsigned char* handles, const int64_t* slab_lengths);
void elph_shard_p2p_close_impl(elph_handle* h);
bool elph_shard_cg_available_impl(elph_handle* h);
bool elph_shard_cg_p2p_impl(elph_handle* h, const double* b_own, double* x_own, double tol, int64_t maxiter, int64_t* iters,
                            double* eps);
bool elph_cg_single_reduction(elph_handle* h, double* x_dev);
// cg_pipe.cu
size_t elph_pipe_arena_bytes(int N, int Lmax);
bool elph_cg_pipe_fits(elph_handle* h);
void elph_shard_halo_impl(elph_handle* h, double* v_own);
HaloArgs elph_shard_halo_args(elph_handle* h, double* v_own, bool advance);
bool elph_cg_pipe_run(elph_handle* h, const double* r0, double* x, bool x0_given, bool scalars_on_device, double tol, int64_t maxiter);
// buffers of nrhs independent solves for the persistent kernels (right-hand side k at + k*vstride / k*pstride / k)
struct CgBatchBufs {
    double* x = nullptr;
    double* R = nullptr;
    double* P0 = nullptr;
    double* P1 = nullptr;
    double* partial = nullptr;      // 2L doubles per right-hand side
    unsigned int* bar = nullptr;
    CgScalars* S = nullptr;
    long long vstride = 0;
    int pstride = 0;
};
bool elph_cg_persistent_batch(elph_handle* h, int nrhs, const CgBatchBufs& B);
// nrhs solves A x_k = b_k on the same field with zero initial guesses (device pointers, one per right-hand side);
// unpreconditioned solves run together in the persistent kernels, everything else falls back to a loop of elph_solve_device
void elph_solve_batch_device(elph_handle* h, int nrhs, const double* const* b_dev, double* const* x_dev, bool use_precond,
                             double tol_power, elph_solve_info* infos);
// reductions: out[0] = sum a*b  (deterministic two-stage); blocking read helpers
void elph_dot_async(elph_handle* h, const double* a, const double* b, int64_t n, double* d_out);
void elph_diffnorm2_async(elph_handle* h, const double* a, const double* b, int64_t n, double* d_out2);  // |a-b|^2, |b|^2

// fft.cu
void elph_fft_init(elph_handle* h);
void elph_tau_to_omega_dev(elph_handle* h, const double* vin, cplx* vout);       // twisted forward, [tau][N] -> [omega][N]
void elph_omega_to_tau_dev(elph_handle* h, const cplx* vin, double* vout);       // inverse + conj twist + real part
void elph_fourier_accelerate_dev(elph_handle* h, const double* vin, double* vout, double power, bool use_mass);

// kpm.cu
void elph_kpm_init(elph_handle* h, int n, double buf, double c1, double c2);
void elph_kpm_setup_impl(elph_handle* h, const double* arnoldi_noise_host, elph_kpm_info* info, const double* ext_eVbar_dev = nullptr,
                         int phase = 0);
bool elph_kpm_can_speculate(const elph_handle* h);
void elph_kpm_setup_begin(elph_handle* h, const double* arnoldi_noise_host);
bool elph_kpm_setup_finish(elph_handle* h, elph_kpm_info* info);
void elph_kpm_set_omega_subset(elph_handle* h, int first, int stride);
void elph_kpm_chains_dev(elph_handle* h, const cplx* nu_in, cplx* nu_out);
// kpm_shard.cu
void elph_kpm_shard_export_impl(elph_handle* h, int rank, int world, int tau0, int lloc, unsigned char* handle_out);
void elph_kpm_shard_open_impl(elph_handle* h, const unsigned char* handles, const int64_t* tau0s);
void elph_kpm_shard_close_impl(elph_handle* h);
void elph_kpm_shard_free(elph_handle* h);
void elph_kpm_shard_apply_impl(elph_handle* h, const double* r_own, double* z_own);
bool elph_kpm_shard_ok(elph_handle* h);
void elph_kpm_apply_dev(elph_handle* h, const double* vin, double* vout);
struct KpmCgFuse {   // vectors of the running preconditioned CG iteration (see fft.cu: CgFuse)
    double* x;
    double* r;
    const double* p;
    const double* ap;
};
void elph_kpm_apply_dev_cg(elph_handle* h, const double* vin, double* vout, const KpmCgFuse* cgf);
void elph_tau_to_omega_cols_dev(elph_handle* h, const double* vin, cplx* vout, int ncols);
void elph_omega_to_tau_cols_dev(elph_handle* h, const cplx* vin, double* vout, int ncols);
void elph_tau_to_omega_dev_cg(elph_handle* h, double* x, double* r, const double* p, const double* ap, cplx* vout);
void elph_omega_to_tau_dev_cg(elph_handle* h, const cplx* vin, double* z, double* r);
void elph_kpm_free(elph_handle* h);
bool elph_launch_kpm_square(elph_handle* h, const cplx* nu_in, cplx* nu_out, const int* skip);
void elph_tau_to_omega_dev_skip(elph_handle* h, const double* vin, cplx* vout, const int* skip);
void elph_omega_to_tau_dev_skip(elph_handle* h, const cplx* vin, double* vout, const int* skip);

// force.cu
void elph_muldMdx_dev(elph_handle* h, const double* u, const double* v, double* out, double scale, bool add_dSb,
                      bool shifted);
void elph_dSbdx_dev(elph_handle* h, double* dSbdx, bool shifted);
void elph_dSbdx_open_dev(elph_handle* h, double* dSbdx, const double* x_own, bool shifted);
void elph_fourier_accelerate_cols_dev(elph_handle* h, const double* vin, double* vout, int ncols, const double* diag, double power);
void elph_Sb_dev(elph_handle* h, bool shifted, double* host_out);

// hmc.cu
void elph_hmc_ensure(elph_handle* h);
void elph_hmc_free(elph_handle* h);
void elph_hmc_refresh_v_dev(elph_handle* h, double alpha, const double* R_dev);
double elph_hmc_refresh_phi_dev(elph_handle* h);
void elph_hmc_calc_Oinv_dev(elph_handle* h, bool use_precond, const double* arnoldi_host, double power, int64_t* iters, int* flag);
void elph_hmc_special_update_dev(elph_handle* h, int kind, int i, int j, bool use_precond, const double* arnoldi_host, double uniform,
                                 int* accepted, double* S0_out, double* S1_out, int64_t* iters, int* flag);
void elph_hmc_calc_H_dev(elph_handle* h, double* H, double* S, double* K);
void elph_hmc_calc_dSfdx_dev(elph_handle* h, double* dS);
void elph_hmc_update_dev(elph_handle* h, double dt, int Nt, int Nb, double alpha, const double* Rv_dev, bool use_precond,
                         const double* arnoldi_host, double uniform, int32_t* accepted, double* iters_out, double* H0out,
                         double* H1out, int32_t* flag_out);

// dynamics.cu
void elph_calc_dSdx_dev(elph_handle* h, const double* g_dev, const double* arnoldi_host, bool use_precond, double* dSdx_dev,
                        double* Minv_dev, elph_solve_info* info);

// opt a kernel in to the large dynamic shared-memory carve-out ONCE per handle (the driver call costs
// microseconds of host time; doing it on every launch made the solver loop launch bound)
template <typename K>
static inline void elph_enable_smem(elph_handle* h, K kernel) {
    if (h->smem_enabled.insert((const void*)kernel).second)
        ELPH_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin));
}

template <typename T>
static inline T* elph_dalloc(size_t n) {
    T* p = nullptr;
    ELPH_CUDA(cudaMalloc(&p, (n ? n : 1) * sizeof(T)));
    return p;
}
template <typename T>
static inline void elph_upload(T* dst, const T* src, size_t n, cudaStream_t s) {
    ELPH_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, s));
}
