// extern "C" entry points of libelph_b200.so (see include/elph_b200.h).
// Host-buffer entry points copy in, convert between the reference host layout
// (tau fastest) and the engine layout ([tau][site]) on the device, run the
// device path and copy out.  There is no CPU implementation of any operator here.
#include "elph_internal.cuh"

#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cstring>

std::string& elph_global_error() {
    static thread_local std::string e;
    return e;
}

void elph_lincomb(elph_handle* h, double* out, double a, const double* X, double b, const double* Y, double c, const double* Z,
                  int64_t n);
void elph_langevin_step_dev(elph_handle* h, int method, double dt, const double* eta_dev, const double* g1_dev,
                            const double* g2_dev, const double* arn1, const double* arn2, bool use_precond, int64_t* iters,
                            elph_solve_info* info1, elph_solve_info* info2);
std::vector<std::complex<double>> elph_debug_hess_eig(const std::vector<double>& h, int n);

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (dev != prev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (cur != prev && prev >= 0) cudaSetDevice(prev);
    }
};

template <typename T>
void up(T*& dst, const std::vector<T>& src) {
    dst = elph_dalloc<T>(src.size());
    if (!src.empty()) ELPH_CUDA(cudaMemcpy(dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
}

double* zeros(size_t n) {
    double* p = elph_dalloc<double>(n);
    ELPH_CUDA(cudaMemset(p, 0, (n ? n : 1) * sizeof(double)));
    return p;
}

// grow-only staging area (device, host layout) for the host-buffer entry points
double* stage(elph_handle* h, int slot, size_t ndoubles) {
    if (h->stage_cap[slot] < ndoubles) {
        if (h->d_stage[slot]) ELPH_CUDA(cudaFree(h->d_stage[slot]));
        h->d_stage[slot] = elph_dalloc<double>(ndoubles);
        h->stage_cap[slot] = ndoubles;
    }
    return h->d_stage[slot];
}

// host vector (host layout, ncols columns of length L) -> engine buffer
void upload_vec(elph_handle* h, const double* host, double* engine_dev, int ncols, int64_t nbatch = 1) {
    ELPH_REQUIRE(host != nullptr, ELPH_ERR_INVALID, "null host input pointer");
    const size_t n = (size_t)ncols * h->L * nbatch;
    double* st = stage(h, 0, n);
    ELPH_CUDA(cudaMemcpyAsync(st, host, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    elph_to_engine(h, st, engine_dev, ncols, nbatch);
}

void download_vec(elph_handle* h, const double* engine_dev, double* host, int ncols, int64_t nbatch = 1) {
    ELPH_REQUIRE(host != nullptr, ELPH_ERR_INVALID, "null host output pointer");
    const size_t n = (size_t)ncols * h->L * nbatch;
    double* st = stage(h, 1, n);
    elph_from_engine(h, engine_dev, st, ncols, nbatch);
    ELPH_CUDA(cudaMemcpyAsync(host, st, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    ELPH_CUDA(cudaStreamSynchronize(h->stream));
}

void build(elph_handle* h, const elph_config* c) {
    ELPH_REQUIRE(c->model == ELPH_MODEL_HOLSTEIN || c->model == ELPH_MODEL_SSH, ELPH_ERR_INVALID, "unknown model kind");
    ELPH_REQUIRE(c->index_base == 0 || c->index_base == 1, ELPH_ERR_INVALID, "index_base must be 0 or 1");
    ELPH_REQUIRE(c->Ltau >= 1 && c->Nsites >= 1 && c->Nbonds >= 0 && c->Nph >= 0, ELPH_ERR_INVALID, "bad dimensions");
    ELPH_REQUIRE(c->Ltau < (1 << 24) && c->Nsites < (1 << 24) && c->Nbonds < (1LL << 30), ELPH_ERR_INVALID, "dimensions too large");
    ELPH_REQUIRE(c->dtau > 0.0, ELPH_ERR_INVALID, "dtau must be positive");
    ELPH_REQUIRE(c->Nbonds == 0 || c->neighbor_table, ELPH_ERR_INVALID, "neighbor_table is NULL");
    ELPH_REQUIRE(c->mu, ELPH_ERR_INVALID, "mu is NULL");
    h->model = c->model;
    h->L = (int)c->Ltau;
    h->N = (int)c->Nsites;
    h->Nb = (int)c->Nbonds;
    h->Nph = (int)c->Nph;
    if (h->model == ELPH_MODEL_HOLSTEIN) ELPH_REQUIRE(c->Nph == c->Nsites, ELPH_ERR_INVALID, "Holstein: Nph must equal Nsites");
    h->Ndim = (int64_t)h->N * h->L;
    h->Ndof = (int64_t)h->Nph * h->L;
    h->dtau = c->dtau;
    h->cg_tol = c->cg_tol > 0 ? c->cg_tol : 1e-4;
    h->cg_maxiter = c->cg_maxiter >= 1 ? c->cg_maxiter : h->Ndim;  // ConjugateGradient(): maxiter<1 -> N (:47-49)
    h->cg_kappa_max = c->cg_kappa_max > 0 ? c->cg_kappa_max : 1e12;

    int dev = c->device;
    if (dev < 0) ELPH_CUDA(cudaGetDevice(&dev));
    ELPH_CUDA(cudaSetDevice(dev));
    h->device = dev;
    cudaDeviceProp prop;
    ELPH_CUDA(cudaGetDeviceProperties(&prop, dev));
    ELPH_REQUIRE(prop.major >= 10, ELPH_ERR_UNSUPPORTED, "libelph_b200 requires a Blackwell (sm_100a) GPU");
    h->sm_count = prop.multiProcessorCount;
    // the handle works on its own non-blocking stream (CUDA graphs cannot be captured on the legacy default stream);
    // elph_set_stream replaces it
    ELPH_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
    // dynamic shared memory available to the slice kernels: the opt-in maximum minus 1 KB kept for their
    // (small) static shared arrays -- cudaFuncAttributeMaxDynamicSharedMemorySize counts static + dynamic
    h->smem_optin = prop.sharedMemPerBlockOptin - 1024;

    // ---- bonds: 0-based pairs in checkerboard order; colour groups recovered from the order:
    // a new group starts at the first bond that touches a site already used in the current group
    // (identical to the reference's greedy groups; any disjoint split gives the same numbers)
    const int64_t base = c->index_base;
    h->bonds_host.resize(h->Nb);
    for (int b = 0; b < h->Nb; ++b) {
        const int64_t i = c->neighbor_table[2 * (size_t)b] - base, j = c->neighbor_table[2 * (size_t)b + 1] - base;
        ELPH_REQUIRE(i >= 0 && i < h->N && j >= 0 && j < h->N && i != j, ELPH_ERR_INVALID, "neighbor_table entry out of range");
        h->bonds_host[b] = make_int2((int)i, (int)j);
    }
    {
        std::vector<int> stamp(h->N, -1);
        h->goff_host.clear();
        h->goff_host.push_back(0);
        int g = 0;
        for (int b = 0; b < h->Nb; ++b) {
            const int2 ij = h->bonds_host[b];
            if (stamp[ij.x] == g || stamp[ij.y] == g) {
                h->goff_host.push_back(b);
                ++g;
            }
            stamp[ij.x] = g;
            stamp[ij.y] = g;
        }
        if (h->Nb > 0) h->goff_host.push_back(h->Nb);
        h->ngroups = (int)h->goff_host.size() - 1;
        h->max_group = 0;
        for (int k = 0; k < h->ngroups; ++k) h->max_group = std::max(h->max_group, h->goff_host[k + 1] - h->goff_host[k]);
    }
    up(h->d_bonds, h->bonds_host);
    up(h->d_goff, h->goff_host);

    auto vec = [&](const double* p, size_t n, double fill) {
        std::vector<double> v(n, fill);
        if (p) std::copy(p, p + n, v.begin());
        return v;
    };
    up(h->d_mu, vec(c->mu, h->N, 0.0));
    up(h->d_omega, vec(c->omega, h->Nph, 0.0));
    up(h->d_omega4, vec(c->omega4, h->Nph, 0.0));
    h->d_x = zeros(h->Ndof);

    if (h->model == ELPH_MODEL_HOLSTEIN) {
        ELPH_REQUIRE(h->Nb == 0 || (c->cosht && c->sinht), ELPH_ERR_INVALID, "cosht/sinht are NULL");
        std::vector<double2> cs(h->Nb);
        for (int b = 0; b < h->Nb; ++b) cs[b] = make_double2(c->cosht[b], c->sinht[b]);
        up(h->d_cs, cs);
        elph_detect_square(h, cs);
        elph_detect_honeycomb(h, cs);
        up(h->d_lam, vec(c->lambda, h->N, 0.0));
        up(h->d_lam2, vec(c->lambda2, h->N, 0.0));
        h->d_D = zeros(h->Ndim);
    } else {
        ELPH_REQUIRE(h->Nb == 0 || (c->t && c->checkerboard_perm && c->inv_checkerboard_perm && c->bond_to_phonon),
                     ELPH_ERR_INVALID, "SSH tables (t, checkerboard_perm, inv_checkerboard_perm, bond_to_phonon) are NULL");
        ELPH_REQUIRE(h->Nph == 0 || (c->alpha && c->phonon_to_bond), ELPH_ERR_INVALID, "SSH phonon tables are NULL");
        up(h->d_t, vec(c->t, h->Nb, 0.0));
        up(h->d_alpha, vec(c->alpha, h->Nph, 0.0));
        up(h->d_alpha2, vec(c->alpha2, h->Nph, 0.0));
        std::vector<int> col_bond(h->Nb), col_ph(h->Nb, -1), ph_col(h->Nph, -1);
        for (int col = 0; col < h->Nb; ++col) {
            const int64_t bond = c->inv_checkerboard_perm[col] - base;
            ELPH_REQUIRE(bond >= 0 && bond < h->Nb, ELPH_ERR_INVALID, "inv_checkerboard_perm out of range");
            ELPH_REQUIRE(c->checkerboard_perm[bond] - base == col, ELPH_ERR_INVALID,
                         "checkerboard_perm and inv_checkerboard_perm are not inverse permutations");
            col_bond[col] = (int)bond;
            const int64_t ph = c->bond_to_phonon[bond] - base;  // 0 (1-based) / -1 (0-based) = no phonon
            if (ph >= 0) {
                ELPH_REQUIRE(ph < h->Nph, ELPH_ERR_INVALID, "bond_to_phonon out of range");
                ELPH_REQUIRE(c->phonon_to_bond[ph] - base == bond, ELPH_ERR_INVALID, "phonon_to_bond/bond_to_phonon mismatch");
                col_ph[col] = (int)ph;
                ph_col[ph] = col;
            }
        }
        for (int ph = 0; ph < h->Nph; ++ph) ELPH_REQUIRE(ph_col[ph] >= 0, ELPH_ERR_INVALID, "phonon without a bond");
        up(h->d_col_bond, col_bond);
        up(h->d_col_ph, col_ph);
        up(h->d_ph_col, ph_col);
        // primary_field (host layout field = ph*L + tau) must be tau-diagonal: field (ph,tau) -> (ph',tau)
        h->primary_ph_host.resize(h->Nph);
        for (int ph = 0; ph < h->Nph; ++ph) {
            int pr = ph;
            if (c->primary_field) {
                for (int tau = 0; tau < h->L; ++tau) {
                    const int64_t f = c->primary_field[(size_t)ph * h->L + tau] - base;
                    ELPH_REQUIRE(f >= 0 && f < h->Ndof && f % h->L == tau, ELPH_ERR_INVALID, "primary_field is not tau-diagonal");
                    if (tau == 0) pr = (int)(f / h->L);
                    ELPH_REQUIRE(f / h->L == pr, ELPH_ERR_INVALID, "primary_field differs between time slices");
                }
            }
            h->primary_ph_host[ph] = pr;
        }
        for (int ph = 0; ph < h->Nph; ++ph)
            ELPH_REQUIRE(h->primary_ph_host[h->primary_ph_host[ph]] == h->primary_ph_host[ph], ELPH_ERR_INVALID,
                         "primary_field does not map onto primary fields");
        up(h->d_primary_ph, h->primary_ph_host);
        // CSR of equivalent phonons per primary, members in neighbour-table column order
        std::vector<int> start(h->Nph + 1, 0), members(h->Nph);
        for (int ph = 0; ph < h->Nph; ++ph) start[h->primary_ph_host[ph] + 1]++;
        for (int k = 0; k < h->Nph; ++k) start[k + 1] += start[k];
        std::vector<int> order(h->Nph);
        for (int k = 0; k < h->Nph; ++k) order[k] = k;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ph_col[a] < ph_col[b]; });
        std::vector<int> fill(start.begin(), start.end() - 1);
        for (int k : order) members[fill[h->primary_ph_host[k]]++] = k;
        up(h->d_grp_start, start);
        up(h->d_grp_members, members);
        h->d_cs = elph_dalloc<double2>((size_t)h->L * h->Nb);
        h->d_tprime = elph_dalloc<double>((size_t)h->L * h->Nb);
        elph_detect_ssh_square(h);
        h->d_D = zeros(h->N);
        h->d_lam = zeros(h->N);
        h->d_lam2 = zeros(h->N);
    }

    // scratch
    const size_t nd = (size_t)std::max(h->Ndim, h->Ndof);
    h->d_va = zeros(nd);
    h->d_vb = zeros(nd);
    h->d_vc = zeros(nd);
    h->d_b = zeros(h->Ndim);
    h->d_res = zeros(h->Ndim);
    h->d_r = zeros(h->Ndim);
    h->d_p[0] = zeros(h->Ndim);
    h->d_p[1] = zeros(h->Ndim);
    h->d_z = zeros(h->Ndim);
    { const char* tr = getenv("ELPH_TRACE"); h->trace = tr && tr[0] == '1'; }
    h->partial_cap = std::max(std::max(4 * h->sm_count, 4 * h->L + 8), h->N / 4 + 8);
    h->d_partial = zeros(h->partial_cap);
    h->d_ticket = elph_dalloc<unsigned int>(1);
    ELPH_CUDA(cudaMemset(h->d_ticket, 0, sizeof(unsigned int)));
    h->d_bar = elph_dalloc<unsigned int>(1);
    ELPH_CUDA(cudaMemset(h->d_bar, 0, sizeof(unsigned int)));
    h->d_cg = elph_dalloc<CgScalars>(1);
    ELPH_CUDA(cudaMemset(h->d_cg, 0, sizeof(CgScalars)));
    ELPH_CUDA(cudaMallocHost(&h->h_cg, sizeof(CgScalars)));
    ELPH_CUDA(cudaMallocHost(&h->h_scal, 8 * sizeof(double)));
    h->d_scal = zeros(8);
    h->d_dSdx = zeros(h->Ndof);
    h->d_dSdx2 = zeros(h->Ndof);
    h->d_eta = zeros(h->Ndof);
    h->d_dx = zeros(h->Ndof);
    h->d_tmp = zeros(h->Ndof);
    h->d_g = zeros(h->Ndim);
    h->d_g2 = zeros(h->Ndim);
    h->d_Minv = zeros(h->Ndim);
    h->d_nu2 = elph_dalloc<cplx>((size_t)h->L * h->N);

    elph_fft_init(h);
    if (c->kpm_n > 0) elph_kpm_init(h, (int)c->kpm_n, c->kpm_buf, c->kpm_c1, c->kpm_c2);

    // Fourier-acceleration diagonals -> engine layout [k][phonon]
    auto load_diag = [&](const double* src, double*& dst, bool& have) {
        if (!src) return;
        dst = elph_dalloc<double>(h->Ndof);
        double* st = stage(h, 0, h->Ndof);
        ELPH_CUDA(cudaMemcpy(st, src, h->Ndof * sizeof(double), cudaMemcpyHostToDevice));
        elph_to_engine(h, st, dst, h->Nph);
        have = true;
    };
    ELPH_CUDA(cudaDeviceSynchronize());   // cudaMemset on the legacy stream does not order with the handle's stream
    load_diag(c->fa_Q, h->d_Q, h->have_Q);
    load_diag(c->fa_M, h->d_Mass, h->have_M);
    elph_launch_update_model(h);
    ELPH_CUDA(cudaStreamSynchronize(h->stream));
}

void destroy(elph_handle* h) {
    DeviceGuard g(h->device);
    cudaDeviceSynchronize();
    elph_kpm_free(h);
    elph_hmc_free(h);
    elph_greens_free(h);
    if (h->g1r.V) cudaFree(h->g1r.V);
    if (h->g1r.partial) cudaFree(h->g1r.partial);
    for (auto& g : h->cg_graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    h->cg_graphs.clear();
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    if (h->pipe.init) {
        for (int b = 0; b < 2; ++b) {
            cudaEventDestroy(h->pipe.ev_in[b]);
            cudaEventDestroy(h->pipe.ev_comp[b]);
            cudaEventDestroy(h->pipe.ev_out[b]);
            for (int k = 0; k < 4; ++k) cudaFree(h->pipe.buf[b][k]);
        }
        cudaStreamDestroy(h->pipe.s_in);
        cudaStreamDestroy(h->pipe.s_out);
    }
    elph_shard_p2p_close_impl(h);
    elph_kpm_shard_free(h);
    if (h->pipe_prof_buf) cudaFree(h->pipe_prof_buf);
    if (h->h_hx_flag) cudaFreeHost(h->h_hx_flag);
    if (h->upload_stream) {
        cudaStreamDestroy(h->upload_stream);
        cudaEventDestroy(h->upload_event);
        cudaEventDestroy(h->upload_fence);
    }
    if (h->d_D_alloc) {  // sharded: d_D points one slice into this allocation
        cudaFree(h->d_D_alloc);
        h->d_D = nullptr;
    }
    if (h->d_cs_alloc) {
        cudaFree(h->d_cs_alloc);
        h->d_cs = nullptr;
    }
    void* ptrs[] = {h->d_bonds, h->d_goff, h->d_cs, h->d_lam, h->d_lam2, h->d_mu, h->d_omega, h->d_omega4, h->d_x, h->d_D,
                    h->d_Q, h->d_Mass, h->d_t, h->d_alpha, h->d_alpha2, h->d_ph_col, h->d_col_ph, h->d_col_bond, h->ssq.d_slot, h->ssq.d_tab,
                    h->d_primary_ph, h->d_grp_start, h->d_grp_members, h->d_tprime, h->d_va, h->d_vb, h->d_vc, h->d_b,
                    h->d_res, h->d_r, h->d_p[0], h->d_p[1], h->d_z, h->d_partial, h->d_ticket, h->d_bar, h->d_cg, h->d_scal,
                    h->d_dSdx, h->d_dSdx2, h->d_eta, h->d_dx, h->d_tmp, h->d_g, h->d_g2, h->d_Minv, h->d_nu2,
                    h->d_twiddle, h->d_theta, h->d_stage[0], h->d_stage[1], h->d_stage[2], h->d_stage[3],
                    h->batch.x, h->batch.r, h->batch.p0, h->batch.p1, h->batch.partial, h->batch.bar, h->batch.S};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (h->h_cg) cudaFreeHost(h->h_cg);
    if (h->h_scal) cudaFreeHost(h->h_scal);
    delete h;
}

}  // namespace

#define ENTER(h)                                                      \
    if (!(h)) {                                                       \
        elph_global_error() = "null handle";                          \
        return ELPH_ERR_INVALID;                                      \
    }                                                                 \
    DeviceGuard guard__((h)->device);                                 \
    ELPH_TRY

extern "C" {

const char* elph_version(void) { return "elph_b200 0.1.0 (sm_100a)"; }

const char* elph_last_error(const elph_handle* h) { return h ? h->err.c_str() : elph_global_error().c_str(); }

int32_t elph_create(const elph_config* cfg, elph_handle** out) {
    elph_handle* h = nullptr;
    elph_handle* none = nullptr;
    ELPH_TRY {
        ELPH_REQUIRE(cfg && out, ELPH_ERR_INVALID, "elph_create: null argument");
        *out = nullptr;
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        ELPH_REQUIRE(e == cudaSuccess && ndev > 0, ELPH_ERR_CUDA,
                     std::string("no CUDA device available (libelph_b200 has no CPU fallback): ") + cudaGetErrorString(e));
        h = new elph_handle();
        try {
            build(h, cfg);
        } catch (...) {
            std::string keep;
            try { throw; } catch (const elph_error& er) { keep = er.msg; } catch (...) { keep = "elph_create failed"; }
            destroy(h);
            h = nullptr;
            throw elph_error{ELPH_ERR_INVALID, keep};
        }
        *out = h;
        return ELPH_OK;
    }
    ELPH_CATCH(none)
}

int32_t elph_destroy(elph_handle* h) {
    if (!h) return ELPH_OK;
    destroy(h);
    return ELPH_OK;
}

int32_t elph_set_stream(elph_handle* h, void* cuda_stream) {
    ENTER(h) {
        ELPH_CUDA(cudaStreamSynchronize(h->stream));
        if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
        h->own_stream = false;
        h->stream = (cudaStream_t)cuda_stream;
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_host_register(elph_handle* h, void* host_ptr, int64_t bytes) {
    ENTER(h) {
        ELPH_REQUIRE(host_ptr && bytes > 0, ELPH_ERR_INVALID, "null pointer or empty range");
        ELPH_CUDA(cudaHostRegister(host_ptr, (size_t)bytes, cudaHostRegisterDefault));
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_host_unregister(elph_handle* h, void* host_ptr) {
    ENTER(h) {
        ELPH_REQUIRE(host_ptr, ELPH_ERR_INVALID, "null pointer");
        ELPH_CUDA(cudaHostUnregister(host_ptr));
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_synchronize(elph_handle* h) {
    ENTER(h) {
        ELPH_CUDA(cudaStreamSynchronize(h->stream));
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_set_solver(elph_handle* h, double tol, int64_t maxiter, double kappa_max) {
    ENTER(h) {
        if (tol > 0.0) h->cg_tol = tol;
        if (maxiter > 0) h->cg_maxiter = maxiter;
        if (kappa_max > 0.0) h->cg_kappa_max = kappa_max;
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_kpm_configure(elph_handle* h, int64_t n, double buf, double c1, double c2) {
    ENTER(h) {
        ELPH_REQUIRE(n >= 1, ELPH_ERR_INVALID, "KPM Krylov dimension n must be >= 1");
        ELPH_CUDA(cudaStreamSynchronize(h->stream));
        elph_kpm_free(h);
        elph_kpm_init(h, (int)n, buf, c1, c2);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_set_fourier_acceleration(elph_handle* h, const double* Q, const double* M) {
    ENTER(h) {
        auto load = [&](const double* src, double*& dst, bool& have) {
            if (!src) return;
            if (!dst) dst = elph_dalloc<double>(h->Ndof);
            double* st = stage(h, 0, h->Ndof);
            ELPH_CUDA(cudaMemcpyAsync(st, src, h->Ndof * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            elph_to_engine(h, st, dst, h->Nph);
            have = true;
        };
        load(Q, h->d_Q, h->have_Q);
        load(M, h->d_Mass, h->have_M);
        ELPH_CUDA(cudaStreamSynchronize(h->stream));
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_set_x(elph_handle* h, const double* x) {
    ENTER(h) {
        upload_vec(h, x, h->d_x, h->Nph);
        ELPH_CUDA(cudaStreamSynchronize(h->stream));
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_get_x(elph_handle* h, double* x) {
    ENTER(h) {
        download_vec(h, h->d_x, x, h->Nph);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_set_mu(elph_handle* h, const double* mu) {
    ENTER(h) {
        ELPH_REQUIRE(mu, ELPH_ERR_INVALID, "mu is NULL");
        ELPH_CUDA(cudaMemcpyAsync(h->d_mu, mu, h->N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        ELPH_CUDA(cudaStreamSynchronize(h->stream));
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_update_model(elph_handle* h) {
    ENTER(h) {
        elph_launch_update_model(h);
        if (h->model == ELPH_MODEL_SSH) {
            // "make sure equivalent fields are equal" (src/SSHModels.jl:547-559): isapprox with default rtol = sqrt(eps)
            std::vector<double> x(h->Ndof);
            ELPH_CUDA(cudaMemcpyAsync(x.data(), h->d_x, h->Ndof * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            ELPH_CUDA(cudaStreamSynchronize(h->stream));
            const double rtol = std::sqrt(2.220446049250313e-16);
            for (int ph = 0; ph < h->Nph; ++ph) {
                const int pr = h->primary_ph_host[ph];
                if (pr == ph) continue;
                for (int tau = 0; tau < h->L; ++tau) {
                    const double a = x[(size_t)tau * h->Nph + ph], b = x[(size_t)tau * h->Nph + pr];
                    if (!(a == b || std::fabs(a - b) <= rtol * std::max(std::fabs(a), std::fabs(b))))
                        throw elph_error{ELPH_ERR_STATE, "equivalent phonon fields differ (update_model!)"};
                }
            }
        }
        ELPH_CUDA(cudaStreamSynchronize(h->stream));
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_get_expnV(elph_handle* h, double* out) {
    ENTER(h) {
        ELPH_REQUIRE(h->model == ELPH_MODEL_HOLSTEIN, ELPH_ERR_INVALID, "expnV exists only for the Holstein model");
        download_vec(h, h->d_D, out, h->N);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_get_cosh_sinh(elph_handle* h, double* cosht, double* sinht) {
    ENTER(h) {
        ELPH_REQUIRE(cosht && sinht, ELPH_ERR_INVALID, "null output");
        if (h->model == ELPH_MODEL_HOLSTEIN) {
            std::vector<double2> cs(h->Nb);
            ELPH_CUDA(cudaMemcpyAsync(cs.data(), h->d_cs, h->Nb * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
            ELPH_CUDA(cudaStreamSynchronize(h->stream));
            for (int b = 0; b < h->Nb; ++b) { cosht[b] = cs[b].x; sinht[b] = cs[b].y; }
        } else {
            std::vector<double2> cs((size_t)h->L * h->Nb);
            ELPH_CUDA(cudaMemcpyAsync(cs.data(), h->d_cs, cs.size() * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
            ELPH_CUDA(cudaStreamSynchronize(h->stream));
            // reference layout: (Ltau, Nbonds) column-major -> index tau + Ltau*col
            for (int tau = 0; tau < h->L; ++tau)
                for (int b = 0; b < h->Nb; ++b) {
                    cosht[(size_t)b * h->L + tau] = cs[(size_t)tau * h->Nb + b].x;
                    sinht[(size_t)b * h->L + tau] = cs[(size_t)tau * h->Nb + b].y;
                }
        }
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

// Batched products from/to HOST buffers: chunks of the batch are pipelined over three streams (H2D copy | layout
// change + kernel + layout change | D2H copy) with double-buffered staging, so that the two PCIe directions and
// the kernel overlap.  The entry point is PCIe-bound: 16 B per point cross the bus for 24 B of HBM traffic.
static void host_matvec_pipelined(elph_handle* h, MatvecMode mode, const double* v, double* y, int64_t nrhs) {
    HostPipe& P = h->pipe;
    // replicas per pipeline stage (tuning key 8).  Measured at 64 replicas per call on B200 / PCIe 5: 8 -> 24.3 k
    // matvecs/s, 4 -> 23.2 k, 2 -> 20.6 k, 1 -> 17.7 k (per-stage launch and event overhead beats the shorter fill and
    // drain); a bare duplex cudaMemcpyAsync of the same bytes reaches 27.7 k/s, i.e. the entry point runs at 88 % of
    // the bus
    const int64_t chunk = std::max(1, std::min(8, h->pipe_chunk));
    const size_t cdoubles = (size_t)8 * h->Ndim;
    if (!P.init) {
        ELPH_CUDA(cudaStreamCreateWithFlags(&P.s_in, cudaStreamNonBlocking));
        ELPH_CUDA(cudaStreamCreateWithFlags(&P.s_out, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            ELPH_CUDA(cudaEventCreateWithFlags(&P.ev_in[b], cudaEventDisableTiming));
            ELPH_CUDA(cudaEventCreateWithFlags(&P.ev_comp[b], cudaEventDisableTiming));
            ELPH_CUDA(cudaEventCreateWithFlags(&P.ev_out[b], cudaEventDisableTiming));
            for (int k = 0; k < 4; ++k) P.buf[b][k] = elph_dalloc<double>(cdoubles);
        }
        P.init = true;
    }
    ELPH_CUDA(cudaStreamSynchronize(h->stream));
    const int64_t nchunks = (nrhs + chunk - 1) / chunk;
    for (int64_t k = 0; k < nchunks; ++k) {
        const int b = (int)(k & 1);
        const int64_t r0 = k * chunk, nr = std::min<int64_t>(chunk, nrhs - r0);
        const size_t nd = (size_t)nr * h->Ndim;
        double *in_host = P.buf[b][0], *in_eng = P.buf[b][1], *out_eng = P.buf[b][2], *out_host = P.buf[b][3];
        if (k >= 2) ELPH_CUDA(cudaStreamWaitEvent(P.s_in, P.ev_comp[b], 0));     // staging slot free again
        ELPH_CUDA(cudaMemcpyAsync(in_host, v + (size_t)r0 * h->Ndim, nd * sizeof(double), cudaMemcpyHostToDevice, P.s_in));
        ELPH_CUDA(cudaEventRecord(P.ev_in[b], P.s_in));
        ELPH_CUDA(cudaStreamWaitEvent(h->stream, P.ev_in[b], 0));
        if (k >= 2) ELPH_CUDA(cudaStreamWaitEvent(h->stream, P.ev_out[b], 0));   // previous D2H of this slot done
        elph_to_engine(h, in_host, in_eng, h->N, nr);
        MatvecArgs a;
        a.v = in_eng;
        a.y = out_eng;
        a.nbatch = nr;
        a.v_stride = h->Ndim;
        a.y_stride = h->Ndim;
        elph_launch_matvec(h, mode, a);
        elph_from_engine(h, out_eng, out_host, h->N, nr);
        ELPH_CUDA(cudaEventRecord(P.ev_comp[b], h->stream));
        ELPH_CUDA(cudaStreamWaitEvent(P.s_out, P.ev_comp[b], 0));
        ELPH_CUDA(cudaMemcpyAsync(y + (size_t)r0 * h->Ndim, out_host, nd * sizeof(double), cudaMemcpyDeviceToHost, P.s_out));
        ELPH_CUDA(cudaEventRecord(P.ev_out[b], P.s_out));
    }
    ELPH_CUDA(cudaStreamSynchronize(P.s_out));
    ELPH_CUDA(cudaStreamSynchronize(h->stream));
}

static int32_t host_matvec(elph_handle* h, MatvecMode mode, const double* v, double* y, int64_t nrhs) {
    ENTER(h) {
        ELPH_REQUIRE(nrhs >= 1, ELPH_ERR_INVALID, "nrhs must be >= 1");
        ELPH_REQUIRE(v && y, ELPH_ERR_INVALID, "null host pointer");
        if (nrhs >= 16) {
            host_matvec_pipelined(h, mode, v, y, nrhs);
            return ELPH_OK;
        }
        double* vin = (nrhs == 1) ? h->d_va : stage(h, 2, (size_t)h->Ndim * nrhs);
        double* vout = (nrhs == 1) ? h->d_vb : stage(h, 3, (size_t)h->Ndim * nrhs);
        upload_vec(h, v, vin, h->N, nrhs);
        MatvecArgs a;
        a.v = vin;
        a.y = vout;
        a.nbatch = nrhs;
        a.v_stride = h->Ndim;
        a.y_stride = h->Ndim;
        elph_launch_matvec(h, mode, a);
        download_vec(h, vout, y, h->N, nrhs);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_mulM(elph_handle* h, const double* v, double* y) { return host_matvec(h, MODE_M, v, y, 1); }
int32_t elph_mulMT(elph_handle* h, const double* v, double* y) { return host_matvec(h, MODE_MT, v, y, 1); }
int32_t elph_mulMTM(elph_handle* h, const double* v, double* y) { return host_matvec(h, MODE_MTM, v, y, 1); }
int32_t elph_mulMTM_batch(elph_handle* h, int64_t nrhs, const double* v, double* y) { return host_matvec(h, MODE_MTM, v, y, nrhs); }

int32_t elph_muldMdx(elph_handle* h, const double* u, const double* v, double* dMdx) {
    ENTER(h) {
        upload_vec(h, u, h->d_va, h->N);
        upload_vec(h, v, h->d_vb, h->N);
        elph_muldMdx_dev(h, h->d_va, h->d_vb, h->d_vc, 1.0, false, false);
        download_vec(h, h->d_vc, dMdx, h->Nph);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_kpm_setup(elph_handle* h, const double* arnoldi_noise, elph_kpm_info* info) {
    ENTER(h) {
        elph_kpm_setup_impl(h, arnoldi_noise, info);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_kpm_apply(elph_handle* h, const double* vin, double* vout) {
    ENTER(h) {
        upload_vec(h, vin, h->d_va, h->N);
        elph_kpm_apply_dev(h, h->d_va, h->d_vb);
        download_vec(h, h->d_vb, vout, h->N);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_kpm_get_orders(elph_handle* h, int64_t* orders) {
    ENTER(h) {
        ELPH_REQUIRE(h->kpm.configured && orders, ELPH_ERR_STATE, "KPM not configured");
        for (int w = 0; w < h->kpm.Lo2; ++w) orders[w] = h->kpm.order[w];
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_kpm_get_coeff(elph_handle* h, int64_t w, double* out) {
    ENTER(h) {
        ELPH_REQUIRE(h->kpm.configured && out && w >= 0 && w < h->kpm.Lo2, ELPH_ERR_INVALID, "bad frequency index");
        for (int m = 0; m < h->kpm.order[w]; ++m) {
            out[2 * m] = h->kpm.coeff[h->kpm.coeff_off[w] + m].real();
            out[2 * m + 1] = h->kpm.coeff[h->kpm.coeff_off[w] + m].imag();
        }
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_cg_solve(elph_handle* h, const double* b, double* x, int32_t use_precond, double tol, int64_t maxiter, int64_t* iters,
                      double* eps) {
    ENTER(h) {
        upload_vec(h, b, h->d_va, h->N);
        upload_vec(h, x, h->d_vb, h->N);
        elph_cg_device(h, h->d_va, h->d_vb, use_precond != 0, tol, maxiter, iters, eps);
        download_vec(h, h->d_vb, x, h->N);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_solve(elph_handle* h, const double* b, double* x, int32_t use_precond, double tol_power, elph_solve_info* info) {
    ENTER(h) {
        upload_vec(h, b, h->d_va, h->N);
        upload_vec(h, x, h->d_vb, h->N);
        elph_solve_device(h, h->d_va, h->d_vb, use_precond != 0, tol_power, info);
        download_vec(h, h->d_vb, x, h->N);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

// nrhs right-hand sides on the current field (host buffers, nrhs consecutive vectors in host layout); X is output only:
// the initial guesses are zero, as in update!(Gr, ...) (src/GreensFunctions.jl:219-225) and calc_O⁻¹Λϕ! (src/HMC.jl:855-885)
int32_t elph_solve_batch(elph_handle* h, int64_t nrhs, const double* B, double* X, int32_t use_precond, double tol_power,
                         elph_solve_info* infos) {
    ENTER(h) {
        ELPH_REQUIRE(nrhs >= 1 && nrhs <= 4096, ELPH_ERR_INVALID, "number of right-hand sides out of range");
        const size_t n = (size_t)h->Ndim;
        double* db = stage(h, 2, n * nrhs);
        double* dx = stage(h, 3, n * nrhs);
        upload_vec(h, B, db, h->N, nrhs);
        std::vector<const double*> bl(nrhs);
        std::vector<double*> xl(nrhs);
        for (int64_t k = 0; k < nrhs; ++k) { bl[k] = db + k * n; xl[k] = dx + k * n; }
        elph_solve_batch_device(h, (int)nrhs, bl.data(), xl.data(), use_precond != 0, tol_power, infos);
        download_vec(h, dx, X, h->N, nrhs);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

// update!(Gr, model, P) (src/GreensFunctions.jl:201-234) for all n_v random vectors at once:
// MinvR[:,k] = (M^T M)^-1 M^T R[:,k], the random vectors R drawn by the caller.  setup!(P) is the caller's (elph_kpm_setup).
int32_t elph_Minv_batch(elph_handle* h, int64_t nrhs, const double* R, double* MinvR, int32_t use_precond,
                        elph_solve_info* infos) {
    ENTER(h) {
        ELPH_REQUIRE(nrhs >= 1 && nrhs <= 4096, ELPH_ERR_INVALID, "number of right-hand sides out of range");
        const size_t n = (size_t)h->Ndim;
        double* dr = stage(h, 2, n * nrhs);
        double* dx = stage(h, 3, n * nrhs);
        upload_vec(h, R, dx, h->N, nrhs);
        MatvecArgs a;   // b_k = M^T r_k, one batched launch
        a.v = dx; a.y = dr; a.nbatch = nrhs; a.v_stride = (int64_t)n; a.y_stride = (int64_t)n;
        elph_launch_matvec(h, MODE_MT, a);
        std::vector<const double*> bl(nrhs);
        std::vector<double*> xl(nrhs);
        for (int64_t k = 0; k < nrhs; ++k) { bl[k] = dr + k * n; xl[k] = dx + k * n; }
        elph_solve_batch_device(h, (int)nrhs, bl.data(), xl.data(), use_precond != 0, 1.0, infos);
        download_vec(h, dx, MinvR, h->N, nrhs);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

// same on device buffers in the engine layout: right-hand side k at b_dev + k*Ndim, solution k at x_dev + k*Ndim
int32_t elph_dev_solve_batch(elph_handle* h, int64_t nrhs, const double* b_dev, double* x_dev, int32_t use_precond,
                             double tol_power, elph_solve_info* infos) {
    ENTER(h) {
        ELPH_REQUIRE(nrhs >= 1 && nrhs <= 4096 && b_dev && x_dev, ELPH_ERR_INVALID, "bad batch arguments");
        const size_t n = (size_t)h->Ndim;
        std::vector<const double*> bl(nrhs);
        std::vector<double*> xl(nrhs);
        for (int64_t k = 0; k < nrhs; ++k) { bl[k] = b_dev + k * n; xl[k] = x_dev + k * n; }
        elph_solve_batch_device(h, (int)nrhs, bl.data(), xl.data(), use_precond != 0, tol_power, infos);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_tau_to_omega(elph_handle* h, const double* vin, double* vout_complex) {
    ENTER(h) {
        ELPH_REQUIRE(vout_complex, ELPH_ERR_INVALID, "null output");
        upload_vec(h, vin, h->d_va, h->N);
        elph_tau_to_omega_dev(h, h->d_va, h->d_nu2);
        // [omega][site] complex -> host layout (site-major, omega fastest)
        cplx* st = reinterpret_cast<cplx*>(stage(h, 1, 2 * (size_t)h->Ndim));
        elph_launch_transpose_c(h, h->d_nu2, st, h->L, h->N);
        ELPH_CUDA(cudaMemcpyAsync(vout_complex, st, h->Ndim * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
        ELPH_CUDA(cudaStreamSynchronize(h->stream));
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_omega_to_tau(elph_handle* h, const double* vin_complex, double* vout) {
    ENTER(h) {
        ELPH_REQUIRE(vin_complex, ELPH_ERR_INVALID, "null input");
        cplx* st = reinterpret_cast<cplx*>(stage(h, 0, 2 * (size_t)h->Ndim));
        ELPH_CUDA(cudaMemcpyAsync(st, vin_complex, h->Ndim * sizeof(cplx), cudaMemcpyHostToDevice, h->stream));
        elph_launch_transpose_c(h, st, h->d_nu2, h->N, h->L);
        elph_omega_to_tau_dev(h, h->d_nu2, h->d_vb);
        download_vec(h, h->d_vb, vout, h->N);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_fourier_accelerate(elph_handle* h, const double* v, double* vout, double power, int32_t use_mass) {
    ENTER(h) {
        upload_vec(h, v, h->d_va, h->Nph);
        elph_fourier_accelerate_dev(h, h->d_va, h->d_vb, power, use_mass != 0);
        download_vec(h, h->d_vb, vout, h->Nph);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_Sb(elph_handle* h, int32_t shifted, double* Sb) {
    ENTER(h) {
        ELPH_REQUIRE(Sb, ELPH_ERR_INVALID, "null output");
        elph_Sb_dev(h, shifted != 0, Sb);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_dSbdx(elph_handle* h, int32_t shifted, double* dSbdx) {
    ENTER(h) {
        upload_vec(h, dSbdx, h->d_va, h->Nph);
        elph_dSbdx_dev(h, h->d_va, shifted != 0);
        download_vec(h, h->d_va, dSbdx, h->Nph);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_calc_dSdx(elph_handle* h, const double* g, const double* arnoldi_noise, int32_t use_precond, double* dSdx,
                       double* Minv_g, elph_solve_info* info) {
    ENTER(h) {
        upload_vec(h, g, h->d_g, h->N);
        elph_calc_dSdx_dev(h, h->d_g, arnoldi_noise, use_precond != 0, h->d_dSdx, h->d_Minv, info);
        download_vec(h, h->d_dSdx, dSdx, h->Nph);
        if (Minv_g) download_vec(h, h->d_Minv, Minv_g, h->N);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_langevin_step(elph_handle* h, int32_t method, double dt, const double* eta, const double* g1, const double* g2,
                           const double* arnoldi1, const double* arnoldi2, int32_t use_precond, int64_t* iters,
                           elph_solve_info* info1, elph_solve_info* info2) {
    ENTER(h) {
        ELPH_REQUIRE(method == ELPH_LANGEVIN_EULER || g2, ELPH_ERR_INVALID, "g2 is required for the two-stage updates");
        elph_trace_mark(h, nullptr);
        upload_vec(h, g1, h->d_g, h->N);          // needed by the first solve
        if (method == ELPH_LANGEVIN_HEUN || !h->overlap_uploads) {
            upload_vec(h, eta, h->d_vc, h->Nph);
            if (g2) upload_vec(h, g2, h->d_g2, h->N);
        } else {
            // eta and g2 are first read after the first solve (src/LangevinDynamics.jl:188,198): their host-to-device copies and
            // layout changes run on a second stream while that solve is under way; elph_langevin_step_dev waits for the event
            if (!h->upload_stream) {
                ELPH_CUDA(cudaStreamCreateWithFlags(&h->upload_stream, cudaStreamNonBlocking));
                ELPH_CUDA(cudaEventCreateWithFlags(&h->upload_event, cudaEventDisableTiming));
                ELPH_CUDA(cudaEventCreateWithFlags(&h->upload_fence, cudaEventDisableTiming));
            }
            // the second stream starts after everything already queued on the main one (the buffers may still be read by it)
            ELPH_CUDA(cudaEventRecord(h->upload_fence, h->stream));
            ELPH_CUDA(cudaStreamWaitEvent(h->upload_stream, h->upload_fence, 0));
            cudaStream_t main = h->stream;
            h->stream = h->upload_stream;          // upload_vec / the transposes queue on h->stream
            try {
                const size_t ne = (size_t)h->Nph * h->L, ng = (size_t)h->N * h->L;
                double* st2 = stage(h, 2, ne);
                ELPH_CUDA(cudaMemcpyAsync(st2, eta, ne * sizeof(double), cudaMemcpyHostToDevice, h->stream));
                elph_to_engine(h, st2, h->d_vc, h->Nph, 1);
                if (g2) {
                    double* st3 = stage(h, 3, ng);
                    ELPH_CUDA(cudaMemcpyAsync(st3, g2, ng * sizeof(double), cudaMemcpyHostToDevice, h->stream));
                    elph_to_engine(h, st3, h->d_g2, h->N, 1);
                }
                ELPH_CUDA(cudaEventRecord(h->upload_event, h->stream));
            } catch (...) {
                h->stream = main;
                throw;
            }
            h->stream = main;
            h->upload_pending = true;
        }
        elph_trace_mark(h, "upload eta, g1, g2");
        elph_langevin_step_dev(h, method, dt, h->d_vc, h->d_g, h->d_g2, arnoldi1, arnoldi2, use_precond != 0, iters, info1, info2);
        ELPH_CUDA(cudaStreamSynchronize(h->stream));
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

// ------------------------------------------------------------------------------- HMC (src/HMC.jl)
int32_t elph_hmc_set_v(elph_handle* h, const double* v) {
    ENTER(h) {
        elph_hmc_ensure(h);
        upload_vec(h, v, h->hmc.v, h->Nph);
        ELPH_CUDA(cudaStreamSynchronize(h->stream));
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_hmc_get(elph_handle* h, int32_t which, double* out) {
    ENTER(h) {
        elph_hmc_ensure(h);
        HmcState& S = h->hmc;
        const double* src[] = {S.v, S.phip, S.phim, S.Lphip, S.Lphim, S.Op, S.Om, S.Lam, S.dS};
        ELPH_REQUIRE(which >= 0 && which <= 8, ELPH_ERR_INVALID, "unknown HMC vector id");
        const bool dof = (which == 0 || which == 8);
        download_vec(h, src[which], out, dof ? h->Nph : h->N);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_hmc_refresh_v(elph_handle* h, double alpha, const double* R) {
    ENTER(h) {
        ELPH_REQUIRE(alpha >= 0.0 && alpha < 1.0, ELPH_ERR_INVALID, "alpha must be in [0,1)");
        elph_hmc_ensure(h);
        upload_vec(h, R, h->d_vc, h->Nph);
        elph_hmc_refresh_v_dev(h, alpha, h->d_vc);
        ELPH_CUDA(cudaStreamSynchronize(h->stream));
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_hmc_refresh_phi(elph_handle* h, const double* R_plus, const double* R_minus, double* S) {
    ENTER(h) {
        elph_hmc_ensure(h);
        upload_vec(h, R_plus, h->hmc.Rp, h->N);
        upload_vec(h, R_minus, h->hmc.Rm, h->N);
        const double s = elph_hmc_refresh_phi_dev(h);
        if (S) *S = s;
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_hmc_calc_Oinv(elph_handle* h, int32_t use_precond, const double* arnoldi_noise, double power, int64_t* iters,
                           int32_t* flag) {
    ENTER(h) {
        elph_hmc_ensure(h);
        int64_t it = 0;
        int fl = 0;
        elph_hmc_calc_Oinv_dev(h, use_precond != 0, arnoldi_noise, power, &it, &fl);
        if (iters) *iters = it;
        if (flag) *flag = fl;
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_hmc_special_update(elph_handle* h, int32_t kind, int64_t i, int64_t j, const double* R_plus, const double* R_minus,
                                const double* arnoldi_noise, int32_t use_precond, double uniform, int32_t* accepted, double* S0,
                                double* S1, int64_t* iters, int32_t* flag) {
    ENTER(h) {
        ELPH_REQUIRE(kind == 0 || kind == 1, ELPH_ERR_INVALID, "kind must be 0 (reflection) or 1 (swap)");
        ELPH_REQUIRE(kind == 1 || h->model == ELPH_MODEL_HOLSTEIN, ELPH_ERR_UNSUPPORTED,
                     "reflection updates exist for the Holstein model only (src/SpecialUpdates.jl:162-165)");
        ELPH_REQUIRE(i >= 0 && i < h->Nph && (kind == 0 || (j >= 0 && j < h->Nph)), ELPH_ERR_INVALID, "phonon index out of range");
        elph_hmc_ensure(h);
        upload_vec(h, R_plus, h->hmc.Rp, h->N);
        upload_vec(h, R_minus, h->hmc.Rm, h->N);
        int acc = 0, fl = 0;
        int64_t it = 0;
        double s0 = 0.0, s1 = 0.0;
        elph_hmc_special_update_dev(h, kind, (int)i, (int)j, use_precond != 0, arnoldi_noise, uniform, &acc, &s0, &s1, &it, &fl);
        ELPH_CUDA(cudaStreamSynchronize(h->stream));
        if (accepted) *accepted = acc;
        if (S0) *S0 = s0;
        if (S1) *S1 = s1;
        if (iters) *iters = it;
        if (flag) *flag = fl;
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

// ------------------------------------------------------------ Green's-function estimator (src/GreensFunctions.jl)
int32_t elph_greens_load(elph_handle* h, int64_t nv, const double* R, const double* MinvR) {
    ENTER(h) {
        ELPH_REQUIRE(nv >= 1 && R && MinvR, ELPH_ERR_INVALID, "nv >= 1 and non-null R, MinvR required");
        elph_greens_load_impl(h, (int)nv, R, MinvR);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_greens_setup(elph_handle* h, int64_t n1, int64_t n2, int64_t L1, int64_t L2, int64_t L3, int64_t norbits, double* G_D0,
                          double* G_D0_G_D0, double* G_DD_G_00, double* G_D0_G_0D) {
    ENTER(h) {
        double* const out[4] = {G_D0, G_D0_G_D0, G_DD_G_00, G_D0_G_0D};
        elph_greens_setup_impl(h, (int)n1, (int)n2, (int)L1, (int)L2, (int)L3, (int)norbits, out);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_hmc_calc_H(elph_handle* h, double* H, double* S, double* K) {
    ENTER(h) {
        elph_hmc_ensure(h);
        double a, b, c;
        elph_hmc_calc_H_dev(h, &a, &b, &c);
        if (H) *H = a;
        if (S) *S = b;
        if (K) *K = c;
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_hmc_calc_dSdx(elph_handle* h, int32_t fermion_only, double* dSdx) {
    ENTER(h) {
        elph_hmc_ensure(h);
        ELPH_CUDA(cudaMemsetAsync(h->hmc.dS, 0, h->Ndof * sizeof(double), h->stream));
        elph_hmc_calc_dSfdx_dev(h, h->hmc.dS);
        if (!fermion_only) elph_dSbdx_dev(h, h->hmc.dS, false);
        download_vec(h, h->hmc.dS, dSdx, h->Nph);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_hmc_update(elph_handle* h, double dt, int64_t Nt, int64_t Nb, double alpha, const double* R_v, const double* R_plus,
                        const double* R_minus, const double* arnoldi_noise, int32_t use_precond, double uniform,
                        int32_t* accepted, double* iters, double* H0, double* H1, int32_t* flag) {
    ENTER(h) {
        elph_trace_mark(h, nullptr);
        ELPH_REQUIRE(Nt >= 0 && Nb >= 1 && dt > 0.0, ELPH_ERR_INVALID, "bad HMC parameters");
        ELPH_REQUIRE(alpha >= 0.0 && alpha < 1.0, ELPH_ERR_INVALID, "alpha must be in [0,1)");
        ELPH_REQUIRE(!use_precond || arnoldi_noise, ELPH_ERR_INVALID, "arnoldi_noise needs (Nt+2)*2*Nsites values");
        if (h->Ndof == 0) {  // update! returns (true, 0) when there is nothing to update (:313,:331-333)
            if (accepted) *accepted = 1;
            if (iters) *iters = 0.0;
            return ELPH_OK;
        }
        elph_hmc_ensure(h);
        upload_vec(h, R_v, h->d_vc, h->Nph);
        upload_vec(h, R_plus, h->hmc.Rp, h->N);
        upload_vec(h, R_minus, h->hmc.Rm, h->N);
        elph_hmc_update_dev(h, dt, (int)Nt, (int)Nb, alpha, h->d_vc, use_precond != 0, arnoldi_noise, uniform, accepted, iters, H0,
                            H1, flag);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

// ------------------------------------------------------------------------------- device-resident API
static int32_t dev_matvec(elph_handle* h, MatvecMode mode, const double* v, double* y) {
    ENTER(h) {
        ELPH_REQUIRE(v && y, ELPH_ERR_INVALID, "null device pointer");
        MatvecArgs a;
        a.v = v;
        a.y = y;
        elph_launch_matvec(h, mode, a);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}
int32_t elph_dev_mulMTM(elph_handle* h, const double* v, double* y) { return dev_matvec(h, MODE_MTM, v, y); }
int32_t elph_dev_mulM(elph_handle* h, const double* v, double* y) { return dev_matvec(h, MODE_M, v, y); }
int32_t elph_dev_mulMT(elph_handle* h, const double* v, double* y) { return dev_matvec(h, MODE_MT, v, y); }

int32_t elph_dev_mulMTM_replicas(elph_handle* h, int64_t nrep, const double* expnV_dev, int64_t expnV_stride, const double* v_dev,
                                 double* y_dev, int64_t vec_stride) {
    ENTER(h) {
        ELPH_REQUIRE(v_dev && y_dev && nrep >= 1, ELPH_ERR_INVALID, "bad replica arguments");
        ELPH_REQUIRE(h->model == ELPH_MODEL_HOLSTEIN, ELPH_ERR_UNSUPPORTED, "replica batches are implemented for the Holstein model");
        MatvecArgs a;
        a.v = v_dev;
        a.y = y_dev;
        a.D = expnV_dev;
        a.nbatch = nrep;
        a.v_stride = vec_stride;
        a.y_stride = vec_stride;
        a.D_stride = expnV_dev ? expnV_stride : 0;
        elph_launch_matvec(h, MODE_MTM, a);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

// SSH: the replicas' tables come from elph_dev_ssh_replica_tables (tile layout of ssh_square.cu), strides in doubles
int32_t elph_dev_ssh_replica_tables(elph_handle* h, int64_t nrep, const double* x_dev, int64_t x_stride, double* tab_dev,
                                    int64_t tab_stride) {
    ENTER(h) {
        ELPH_REQUIRE(x_dev && tab_dev && nrep >= 1, ELPH_ERR_INVALID, "bad replica arguments");
        ELPH_REQUIRE(tab_stride % 2 == 0 && tab_stride >= 4LL * h->L * h->N && x_stride >= (int64_t)h->L * h->Nph, ELPH_ERR_INVALID,
                     "replica strides too small (table: 4*Ltau*Nsites doubles, field: Ltau*Nph doubles)");
        elph_launch_ssh_replica_tables(h, nrep, x_dev, x_stride, reinterpret_cast<double2*>(tab_dev), tab_stride / 2);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}
int32_t elph_dev_mulMTM_replicas_ssh(elph_handle* h, int64_t nrep, const double* tab_dev, int64_t tab_stride, const double* v_dev,
                                     double* y_dev, int64_t vec_stride) {
    ENTER(h) {
        ELPH_REQUIRE(v_dev && y_dev && tab_dev && nrep >= 1, ELPH_ERR_INVALID, "bad replica arguments");
        ELPH_REQUIRE(h->model == ELPH_MODEL_SSH && h->ssq.enabled && !h->sq_disable, ELPH_ERR_UNSUPPORTED,
                     "SSH replica batches need a periodic square lattice (register-tile kernel)");
        ELPH_REQUIRE(tab_stride % 2 == 0 && tab_stride >= 4LL * h->L * h->N && vec_stride >= h->Ndim, ELPH_ERR_INVALID,
                     "replica strides too small");
        ELPH_REQUIRE((reinterpret_cast<uintptr_t>(tab_dev) % 16) == 0 && (tab_stride * sizeof(double)) % 16 == 0 &&
                         (reinterpret_cast<uintptr_t>(v_dev) % 16) == 0 && (vec_stride * sizeof(double)) % 16 == 0,
                     ELPH_ERR_INVALID, "replica tables and vectors must be 16-byte aligned (TMA bulk copies)");
        MatvecArgs a;
        a.v = v_dev;
        a.y = y_dev;
        a.nbatch = nrep;
        a.v_stride = vec_stride;
        a.y_stride = vec_stride;
        a.ssh_tab = reinterpret_cast<const double2*>(tab_dev);
        a.ssh_tab_stride = tab_stride / 2;
        ELPH_REQUIRE(elph_launch_ssh_square(h, a), ELPH_ERR_UNSUPPORTED, "lattice shape not served by the SSH register-tile kernel");
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

// ---- tau-sharding (multi-GPU): this handle owns global slices [tau0, tau0 + Ltau) of Lglob -------------------
int32_t elph_set_shard(elph_handle* h, int64_t tau0, int64_t Lglob) {
    ENTER(h) {
        ELPH_REQUIRE(Lglob >= h->L && tau0 >= 0 && tau0 + h->L <= Lglob, ELPH_ERR_INVALID, "bad shard bounds");
        ELPH_CUDA(cudaStreamSynchronize(h->stream));
        if (!h->sharded && h->model == ELPH_MODEL_SSH) {
            // SSH: the per-tau (cosh, sinh) table is what couples to the neighbour slab (K(b) of the right neighbour's first slice
            // is needed to recompute (M v)(b), src/SSHModels.jl:581-701); expmu has no time index.  Same re-homing, rows of Nb.
            double2* alloc = elph_dalloc<double2>((size_t)(h->L + 2) * h->Nb);
            ELPH_CUDA(cudaMemset(alloc, 0, (size_t)(h->L + 2) * h->Nb * sizeof(double2)));
            ELPH_CUDA(cudaMemcpy(alloc + h->Nb, h->d_cs, (size_t)h->L * h->Nb * sizeof(double2), cudaMemcpyDeviceToDevice));
            ELPH_CUDA(cudaFree(h->d_cs));
            h->d_cs_alloc = alloc;
            h->d_cs = alloc + h->Nb;
        } else if (!h->sharded) {
            // re-home expnV with one halo slice on each side: [halo_lo][own ...][halo_hi]
            double* alloc = elph_dalloc<double>((size_t)(h->L + 2) * h->N);
            ELPH_CUDA(cudaMemset(alloc, 0, (size_t)(h->L + 2) * h->N * sizeof(double)));
            ELPH_CUDA(cudaMemcpy(alloc + h->N, h->d_D, h->Ndim * sizeof(double), cudaMemcpyDeviceToDevice));
            ELPH_CUDA(cudaFree(h->d_D));
            h->d_D_alloc = alloc;
            h->d_D = alloc + h->N;
        }
        ELPH_CUDA(cudaDeviceSynchronize());
        h->sharded = true;
        h->shard_tau0 = (int)tau0;
        h->shard_Lglob = (int)Lglob;
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_dev_shard_matvec(elph_handle* h, int32_t mode, const double* v_own, double* y_own) {
    ENTER(h) {
        ELPH_REQUIRE(h->sharded, ELPH_ERR_STATE, "elph_set_shard has not been called");
        ELPH_REQUIRE(mode >= 0 && mode <= 2 && v_own && y_own, ELPH_ERR_INVALID, "bad arguments");
        MatvecArgs a;
        a.v = v_own;
        a.y = y_own;
        a.open = true;
        elph_launch_matvec(h, (MatvecMode)mode, a);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

// ---- peer-memory CG over NVLink (cg_p2p.cu): one process per GPU, arenas shared through CUDA IPC ------------------
int32_t elph_shard_p2p_export(elph_handle* h, int32_t rank, int32_t world, unsigned char* ipc_handle_out) {
    ENTER(h) {
        ELPH_REQUIRE(ipc_handle_out, ELPH_ERR_INVALID, "null output");
        elph_shard_p2p_export_impl(h, rank, world, ipc_handle_out);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_shard_p2p_open(elph_handle* h, const unsigned char* ipc_handles, const int64_t* slab_lengths) {
    ENTER(h) {
        elph_shard_p2p_open_impl(h, ipc_handles, slab_lengths);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_dev_shard_cg_p2p(elph_handle* h, const double* b_own, double* x_own, double tol, int64_t maxiter, int64_t* iters,
                              double* eps) {
    ENTER(h) {
        ELPH_REQUIRE(b_own && x_own, ELPH_ERR_INVALID, "null device pointer");
        ELPH_REQUIRE(elph_shard_cg_p2p_impl(h, b_own, x_own, tol, maxiter, iters, eps), ELPH_ERR_UNSUPPORTED,
                     "peer-memory CG: the slab's time slices are not all co-resident on this GPU (or unsupported lattice)");
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_dev_shard_matvec_halo(elph_handle* h, int32_t mode, double* v_own, double* y_own) {
    ENTER(h) {
        ELPH_REQUIRE(h->sharded, ELPH_ERR_STATE, "elph_set_shard has not been called");
        ELPH_REQUIRE(mode >= 0 && mode <= 2 && v_own && y_own, ELPH_ERR_INVALID, "bad arguments");
        if (mode == MODE_MTM && h->halo_fused && (h->sq.enabled && !h->sq_disable)) {
            // one launch: the exchange travels inside the product kernel, behind the interior of the slab
            MatvecArgs f;
            f.v = v_own;
            f.y = y_own;
            f.open = true;
            f.halo = elph_shard_halo_args(h, v_own, false);
            if (elph_launch_mtm_square(h, f)) {
                h->p2p.hx_seq++;
                return ELPH_OK;
            }
        }
        elph_shard_halo_impl(h, v_own);
        MatvecArgs a;
        a.v = v_own;
        a.y = y_own;
        a.open = true;
        elph_launch_matvec(h, (MatvecMode)mode, a);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_shard_cg_available(elph_handle* h, int32_t* available) {
    ENTER(h) {
        ELPH_REQUIRE(available, ELPH_ERR_INVALID, "null output");
        *available = elph_shard_cg_available_impl(h) ? 1 : 0;
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_dev_shard_halo(elph_handle* h, double* v_own) {
    ENTER(h) {
        ELPH_REQUIRE(h->sharded && v_own, ELPH_ERR_INVALID, "elph_set_shard has not been called / null pointer");
        elph_shard_halo_impl(h, v_own);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_dev_shard_muldMdx(elph_handle* h, const double* u_own, const double* v_own, double* out, double scale) {
    ENTER(h) {
        ELPH_REQUIRE(h->sharded, ELPH_ERR_STATE, "elph_set_shard has not been called");
        elph_muldMdx_dev(h, u_own, v_own, out, scale, false, false);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_dev_shard_dSbdx(elph_handle* h, double* dSbdx_own, const double* x_own, int32_t shifted) {
    ENTER(h) {
        ELPH_REQUIRE(dSbdx_own && x_own, ELPH_ERR_INVALID, "null device pointer");
        elph_dSbdx_open_dev(h, dSbdx_own, x_own, shifted != 0);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_dev_fourier_accelerate_cols(elph_handle* h, const double* vin_dev, double* vout_dev, int64_t ncols,
                                         const double* diag_dev, double power) {
    ENTER(h) {
        elph_fourier_accelerate_cols_dev(h, vin_dev, vout_dev, (int)ncols, diag_dev, power);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

// ---- tau-sharded KPM preconditioner (sharded.py: ShardedKPM): site-sharded FFT stage and omega-sharded chain stage ----
int32_t elph_dev_tau_to_omega_cols(elph_handle* h, const double* vin_dev, double* nu_dev, int64_t ncols) {
    ENTER(h) {
        elph_tau_to_omega_cols_dev(h, vin_dev, reinterpret_cast<cplx*>(nu_dev), (int)ncols);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_dev_omega_to_tau_cols(elph_handle* h, const double* nu_dev, double* vout_dev, int64_t ncols) {
    ENTER(h) {
        elph_omega_to_tau_cols_dev(h, reinterpret_cast<const cplx*>(nu_dev), vout_dev, (int)ncols);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_dev_kpm_setup_bar(elph_handle* h, const double* eVbar_dev, const double* arnoldi_noise, elph_kpm_info* info) {
    ENTER(h) {
        ELPH_REQUIRE(eVbar_dev, ELPH_ERR_INVALID, "null tau-mean");
        elph_kpm_setup_impl(h, arnoldi_noise, info, eVbar_dev);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_kpm_set_omega_subset(elph_handle* h, int64_t first, int64_t stride) {
    ENTER(h) {
        elph_kpm_set_omega_subset(h, (int)first, (int)stride);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_dev_kpm_chains(elph_handle* h, const double* nu_in_dev, double* nu_out_dev) {
    ENTER(h) {
        elph_kpm_chains_dev(h, reinterpret_cast<const cplx*>(nu_in_dev), reinterpret_cast<cplx*>(nu_out_dev));
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

// the same application with the three transposes through peer memory (kpm_shard.cu): no NCCL call per application
int32_t elph_kpm_shard_export(elph_handle* h, int32_t rank, int32_t world, int64_t tau0, int64_t lloc, unsigned char* ipc_handle_out) {
    ENTER(h) {
        ELPH_REQUIRE(ipc_handle_out, ELPH_ERR_INVALID, "null output");
        elph_kpm_shard_export_impl(h, rank, world, (int)tau0, (int)lloc, ipc_handle_out);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_kpm_shard_open(elph_handle* h, const unsigned char* ipc_handles, const int64_t* slab_starts) {
    ENTER(h) {
        elph_kpm_shard_open_impl(h, ipc_handles, slab_starts);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_dev_kpm_shard_apply(elph_handle* h, const double* r_own_dev, double* z_own_dev) {
    ENTER(h) {
        elph_kpm_shard_apply_impl(h, r_own_dev, z_own_dev);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_kpm_shard_check(elph_handle* h) {
    ENTER(h) {
        ELPH_REQUIRE(elph_kpm_shard_ok(h), ELPH_ERR_STATE, "sharded KPM apply: a peer GPU did not reach a barrier (timeout)");
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_dev_update_model(elph_handle* h) {
    ENTER(h) {
        elph_launch_update_model(h);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_dev_lincomb(elph_handle* h, double* out, double a, const double* X, double b, const double* Y, double c, const double* Z,
                         int64_t n) {
    ENTER(h) {
        ELPH_REQUIRE(out && X && n >= 0, ELPH_ERR_INVALID, "bad arguments");
        elph_lincomb(h, out, a, X, b, Y, c, Z, n);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_dev_dot(elph_handle* h, const double* a, const double* b, int64_t n, double* out_dev) {
    ENTER(h) {
        ELPH_REQUIRE(a && b && out_dev, ELPH_ERR_INVALID, "bad arguments");
        elph_dot_async(h, a, b, n, out_dev);   // out_dev[0] = a.b (out_dev needs room for 2 doubles)
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_dev_to_engine_layout(elph_handle* h, const double* host_layout_dev, double* engine_dev, int64_t ncols) {
    ENTER(h) {
        elph_to_engine(h, host_layout_dev, engine_dev, (int)ncols);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}
int32_t elph_dev_from_engine_layout(elph_handle* h, const double* engine_dev, double* host_layout_dev, int64_t ncols) {
    ENTER(h) {
        elph_from_engine(h, engine_dev, host_layout_dev, (int)ncols);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}
int32_t elph_dev_ptr_x(elph_handle* h, double** x_dev) {
    ENTER(h) {
        *x_dev = h->d_x;
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}
int32_t elph_dev_ptr_expnV(elph_handle* h, double** p) {
    ENTER(h) {
        *p = h->d_D;
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}
int32_t elph_dev_ptr_cosh_sinh(elph_handle* h, double** p) {
    ENTER(h) {
        *p = reinterpret_cast<double*>(h->d_cs);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}
int32_t elph_dev_cg_solve(elph_handle* h, const double* b_dev, double* x_dev, int32_t use_precond, double tol, int64_t maxiter,
                          int64_t* iters, double* eps) {
    ENTER(h) {
        elph_cg_device(h, b_dev, x_dev, use_precond != 0, tol, maxiter, iters, eps);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}
int32_t elph_dev_kpm_apply(elph_handle* h, const double* vin_dev, double* vout_dev) {
    ENTER(h) {
        elph_kpm_apply_dev(h, vin_dev, vout_dev);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}
int32_t elph_dev_fourier_accelerate(elph_handle* h, const double* vin_dev, double* vout_dev, double power, int32_t use_mass) {
    ENTER(h) {
        elph_fourier_accelerate_dev(h, vin_dev, vout_dev, power, use_mass != 0);
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}
int64_t elph_launch_count(const elph_handle* h) { return h ? h->launches : 0; }
int32_t elph_set_chunk(elph_handle* h, int32_t c) {
    ENTER(h) {
        ELPH_REQUIRE(c >= 0 && c <= 64, ELPH_ERR_INVALID, "slices_per_cta out of range");
        h->chunk_override = c;
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

int32_t elph_set_tuning(elph_handle* h, int32_t key, int32_t value) {
    ENTER(h) {
        switch (key) {
            case 0: ELPH_REQUIRE(value >= 0 && value <= 64, ELPH_ERR_INVALID, "chunk out of range"); h->chunk_override = value; break;
            case 1: h->sq_disable = (value != 0); break;
            case 2: h->sq_py = value; h->kpm_version++; break;
            case 3: h->use_graphs = (value != 0); break;
            case 4: h->kpm_split = (value != 0); h->kpm_version++; break;
            case 5: h->use_persistent = (value != 0); break;
            case 6: h->pcg_fuse = (value != 0); h->kpm_version++; break;
            case 7: h->cg_single_reduction = (value < 0) ? -1 : (value != 0); break;
            case 9: h->hmc_fused_inner = (value != 0); break;
            case 10: h->cg_pipeline = (value < 0) ? -1 : (value != 0); break;
            case 13: ELPH_REQUIRE(value >= 0 && value <= 16, ELPH_ERR_INVALID, "variant out of range"); h->pipe_variant = value; break;
            case 15: h->pipe_sync_mode = value; break;
            case 16: h->kpm_fast = (value != 0); h->kpm_version++; break;
            case 17: h->pcg_persistent = (value != 0); break;
            case 18: h->pcg_half_fft = (value != 0); break;
            case 19: h->kpm_dev_arnoldi = (value != 0); break;
            case 21: h->hc_tiles = (value != 0); break;
            case 22: h->halo_fused = (value != 0); break;
            case 23: h->overlap_uploads = (value != 0); break;
            case 25: h->kpm_speculate = (value != 0); break;
            case 26: h->kpm_wide = (value != 0); h->kpm_version++; break;
            case 24: h->mtm_tanh = (value != 0); break;
            case 20: ELPH_REQUIRE(value >= 0 && value <= 4096, ELPH_ERR_INVALID, "CTA count out of range"); h->pcg_grid = value; break;
            case 14: ELPH_REQUIRE(value >= 0 && value <= 64, ELPH_ERR_INVALID, "slices per CTA out of range"); h->pipe_spc = value; break;
            case 12:
                h->pipe_prof = (value != 0);
                if (h->pipe_prof && !h->pipe_prof_buf) {
                    h->pipe_prof_buf = elph_dalloc<unsigned long long>(8192 * 8);
                    ELPH_CUDA(cudaMemset(h->pipe_prof_buf, 0, 8192 * 8 * sizeof(unsigned long long)));
                }
                break;
            case 11: ELPH_REQUIRE(value >= 0 && value <= 8, ELPH_ERR_INVALID, "CTAs per slice out of range"); h->pipe_ys = value; break;
            case 8: ELPH_REQUIRE(value >= 1 && value <= 8, ELPH_ERR_INVALID, "pipeline stage must hold 1..8 replicas"); h->pipe_chunk = value; break;
            default: ELPH_REQUIRE(false, ELPH_ERR_INVALID, "unknown tuning key");
        }
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}
int32_t elph_get_tuning(elph_handle* h, int32_t key, int32_t* value) {
    ENTER(h) {
        ELPH_REQUIRE(value, ELPH_ERR_INVALID, "null output");
        switch (key) {
            case 0: *value = h->chunk_override; break;
            case 1: *value = h->sq_disable ? 1 : 0; break;
            case 3: *value = h->use_graphs ? 1 : 0; break;
            case 5: *value = h->use_persistent ? 1 : 0; break;
            case 7: *value = h->cg_single_reduction; break;
            case 10: *value = h->cg_pipeline; break;
            case 11: *value = h->pipe_ys; break;
            case 100: *value = h->pipe_last_variant; break;
            case 101: *value = h->pipe_last_spc; break;
            default: ELPH_REQUIRE(false, ELPH_ERR_INVALID, "unknown tuning key");
        }
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}
int32_t elph_get_kernel_info(elph_handle* h, int32_t* square_kernel, int32_t* ngroups) {
    ENTER(h) {
        if (square_kernel) *square_kernel = ((h->sq.enabled || h->ssq.enabled) && !h->sq_disable) ? 1 : 0;
        if (ngroups) *ngroups = h->ngroups;
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

// development hook (not in the public header): per-phase cycle counters of the last pipelined CG solve, [ncta][8]
int32_t elph_debug_pipe_prof(elph_handle* h, int32_t ncta, unsigned long long* out) {
    ENTER(h) {
        ELPH_REQUIRE(h->pipe_prof_buf && out && ncta >= 1 && ncta <= 8192, ELPH_ERR_STATE, "profiling not enabled (tuning key 12)");
        ELPH_CUDA(cudaStreamSynchronize(h->stream));
        ELPH_CUDA(cudaMemcpy(out, h->pipe_prof_buf, (size_t)ncta * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        return ELPH_OK;
    }
    ELPH_CATCH(h)
}

// test hook (not in the public header): eigenvalues of a real upper-Hessenberg matrix, row-major n x n
int32_t elph_debug_hessenberg_eigvals(int32_t n, const double* hmat, double* re, double* im) {
    elph_handle* none = nullptr;
    ELPH_TRY {
        std::vector<double> hv(hmat, hmat + (size_t)n * n);
        auto ev = elph_debug_hess_eig(hv, n);
        for (int i = 0; i < n; ++i) { re[i] = ev[i].real(); im[i] = ev[i].imag(); }
        return ELPH_OK;
    }
    ELPH_CATCH(none)
}

}  // extern "C"
