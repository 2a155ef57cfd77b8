"""GPU parity at BASELINE.json's full sizes (configs B, C, E), through the C ABI.

At these sizes the checks are (i) direct comparison with the fast oracles (the C restatement for products and CG,
NumPy for the force / KPM / PCG, all of which finish in seconds) and (ii) size-independent properties of the
operator: adjointness, fused-vs-composed M^T M, linearity, true-residual of the solves, agreement of the two kernel
families.
"""
import numpy as np
import pytest

from helpers import engine_holstein_like, oracle_holstein, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def config_B():
    """Holstein square 32x32, beta = 20, dtau = 0.1 -> Ltau = 200 (N*Ltau = 204,800), shipped parameters."""
    om, rng = oracle_holstein("square", 32, 20.0, 0.1, mu=-1.0, seed=1234, eps=0.3)
    em = engine_holstein_like(om)
    yield om, em, rng
    em.close()


def test_B_products_against_c_restatement(config_B):
    import elphdynamics_b200 as E
    from oracle.cref import CRef
    om, em, rng = config_B
    c = CRef(om)
    v = rng.normal(size=om.Ndim)
    yc, ye = np.zeros(om.Ndim), np.zeros(om.Ndim)
    for fc, fe in ((c.mulM, E.mulM_), (c.mulMT, E.mulMT_), (c.mulMTM, E.mulMTM_)):
        fc(yc, v)
        fe(ye, em, v)
        assert relerr(ye, yc) <= 1e-12, fe.__name__


def test_B_operator_properties(config_B):
    import elphdynamics_b200 as E
    om, em, rng = config_B
    u, v = rng.normal(size=om.Ndim), rng.normal(size=om.Ndim)
    Mv, Mtu, MtMv, comp = (np.zeros(om.Ndim) for _ in range(4))
    E.mulM_(Mv, em, v)
    E.mulMT_(Mtu, em, u)
    assert abs(u @ Mv - Mtu @ v) <= 1e-12 * np.linalg.norm(u) * np.linalg.norm(Mv)        # adjointness
    E.mulMTM_(MtMv, em, v)
    E.mulMT_(comp, em, Mv)
    assert relerr(MtMv, comp) <= 1e-13                                                        # fused == composed
    a, b = 0.37, -1.9
    lin = np.zeros(om.Ndim)
    E.mulMTM_(lin, em, a * u + b * v)
    MtMu = np.zeros(om.Ndim)
    E.mulMTM_(MtMu, em, u)
    assert relerr(lin, a * MtMu + b * MtMv) <= 1e-13                                          # linearity
    assert v @ MtMv > 0                                                                       # positive definite
    # both kernel families agree (register/shuffle vs generic shared-memory)
    em._call("elph_set_tuning", 1, 1)
    gen = np.zeros(om.Ndim)
    E.mulMTM_(gen, em, v)
    em._call("elph_set_tuning", 1, 0)
    assert relerr(MtMv, gen) <= 1e-14


def test_B_cg_iterations_and_residual(config_B):
    """CG at the shipped tolerance: iteration count within +-2 of the reference algorithm (C restatement) and the
    true residual below sqrt(tol) (src/Models.jl:94-126)."""
    import elphdynamics_b200 as E
    from oracle.cref import CRef
    om, em, rng = config_B
    g = rng.normal(size=om.Ndim)
    b = np.zeros(om.Ndim)
    om.mulMT(b, g)
    xc = np.zeros(om.Ndim)
    it_c, eps_c = CRef(om).cg(xc, b, tol=om.tol, maxiter=om.maxiter)
    xe = np.zeros(om.Ndim)
    it_e, res_e, flag_e = E.ldiv_(xe, em, b)
    assert flag_e == 0 and abs(it_e - it_c) <= 2, (it_e, it_c)
    chk = np.zeros(om.Ndim)
    E.mulMTM_(chk, em, xe)
    true_res = np.linalg.norm(chk - b) / np.linalg.norm(b)
    assert true_res <= np.sqrt(om.tol) and abs(true_res - res_e) <= 1e-8
    assert relerr(xe, xc) <= 1e-3
    # default = single-reduction persistent kernel (csrc/cg_p2p.cu): run-to-run deterministic (fixed-order reductions)
    x3 = np.zeros(om.Ndim)
    E.ldiv_(x3, em, b)
    assert np.array_equal(x3, xe)
    # the two-reduction recurrences of the reference in their three execution strategies: persistent cooperative kernel,
    # CUDA-graph replay of two-kernel iterations, plain launches -- same iteration count (+-1) and the same solution
    em._call("elph_set_tuning", 7, 0)
    xp = np.zeros(om.Ndim)
    it_p, res_p, flag_p = E.ldiv_(xp, em, b)
    assert flag_p == 0 and abs(it_p - it_c) <= 2, (it_p, it_c)
    assert relerr(xp, xe) <= 1e-3                   # the two forms agree to the solve tolerance
    for persistent, graphs in ((0, 1), (0, 0)):
        em._call("elph_set_tuning", 5, persistent)
        em._call("elph_set_tuning", 3, graphs)
        x2 = np.zeros(om.Ndim)
        it2, res2, flag2 = E.ldiv_(x2, em, b)
        assert flag2 == 0 and abs(it2 - it_p) <= 1, (persistent, graphs, it2, it_p)
        assert relerr(x2, xp) <= 1e-6
    em._call("elph_set_tuning", 5, 1)
    em._call("elph_set_tuning", 3, 1)
    x4 = np.zeros(om.Ndim)
    E.ldiv_(x4, em, b)
    assert np.array_equal(x4, xp)
    em._call("elph_set_tuning", 7, -1)


def test_B_cg_with_initial_guess_and_maxiter(config_B):
    """Raw solve!(x, A, b, cg) (src/IterativeSolvers.jl:239-314) from a non-zero initial guess, and a solve cut off by
    maxiter, in both forms of the persistent kernel: iteration counts within +-2 of the C restatement, same iterates."""
    import elphdynamics_b200 as E
    from oracle.cref import CRef
    om, em, rng = config_B
    b = rng.normal(size=om.Ndim)
    x0 = 0.05 * rng.normal(size=om.Ndim)
    xc = x0.copy()
    it_c, eps_c = CRef(om).cg(xc, b, tol=om.tol, maxiter=om.maxiter)
    xm = x0.copy()
    it_m, eps_m = CRef(om).cg(xm, b, tol=om.tol, maxiter=25)
    assert it_m == 25
    for key7 in (1, 0):
        em._call("elph_set_tuning", 7, key7)
        xe = x0.copy()
        it_e = E.solve_(xe, em, b)
        assert abs(it_e - it_c) <= 2, (key7, it_e, it_c)
        assert em.last_eps < om.tol and relerr(xe, xc) <= 1e-3
        xe = x0.copy()
        it_e = E.solve_(xe, em, b, maxiter=25)
        assert it_e == 25 and abs(em.last_eps - eps_m) <= 1e-6 * eps_m, (key7, em.last_eps, eps_m)
        assert relerr(xe, xm) <= 1e-8
    em._call("elph_set_tuning", 7, -1)


def test_B_kpm_pcg_and_force(config_B):
    import elphdynamics_b200 as E
    from oracle.kpm import KPMPreconditioner
    from oracle.solvers import ConjugateGradient, ldiv
    om, em, rng = config_B
    Po, Pe = KPMPreconditioner(om), E.SymmetricKPMPreconditioner(em)
    noise = rng.normal(size=2 * om.N)
    Po.setup(noise)
    info = E.setup_(Pe, noise)
    assert info.active == 1 and Po.active
    assert np.array_equal(Pe.orders(), Po.order)
    # preconditioner apply: all three kernel variants (2-CTA cluster split, single CTA, generic shared-memory)
    # against the oracle at the engine's spectral window
    from oracle.kpm import kpm_coefficients
    Po.lam_lo, Po.lam_hi = info.lambda_lo, info.lambda_hi
    Po.lam_avg, Po.lam_mag = (Po.lam_hi + Po.lam_lo) / 2, (Po.lam_hi - Po.lam_lo) / 2
    Po.coeff = [kpm_coefficients(int(Po.order[w]), Po.lam_lo, Po.lam_hi, Po.phis[w]) for w in range(Po.Lo2)]
    r = rng.normal(size=om.Ndim)
    zo = np.zeros(om.Ndim)
    Po.ldiv(zo, r)
    for key, val in ((4, 1), (4, 0), (1, 1)):
        em._call("elph_set_tuning", key, val)
        ze = np.zeros(om.Ndim)
        E.kpm_ldiv_(ze, Pe, r)
        assert relerr(ze, zo) <= 1e-11, (key, val)
    em._call("elph_set_tuning", 1, 0)
    em._call("elph_set_tuning", 4, 1)
    g = rng.normal(size=om.Ndim)
    b = np.zeros(om.Ndim)
    om.mulMT(b, g)
    cg = ConjugateGradient(om.Ndim, tol=om.tol, maxiter=om.maxiter)
    xo, xe = np.zeros(om.Ndim), np.zeros(om.Ndim)
    it_o, _, fo = ldiv(xo, om, b, cg, Po)
    it_e, res_e, fe = E.ldiv_(xe, em, b, Pe)
    assert fo == fe == 0 and abs(it_o - it_e) <= 2, (it_o, it_e)
    # the same preconditioned solve as launches per phase (key 17 = 0: the default is the one-kernel form of pcg_fused.cu),
    # then also without the FFT-fused vector updates and without CUDA graphs: same iteration count (+-1: the partial sums
    # are folded in a different order) and the same solution to the solver's accuracy
    em._call("elph_set_tuning", 17, 0)
    x1 = np.zeros(om.Ndim)
    it1, _, f1 = E.ldiv_(x1, em, b, Pe)
    assert f1 == 0 and abs(it1 - it_e) <= 1 and relerr(x1, xe) <= 1e-6, (it1, it_e)
    for key in (6, 3):
        em._call("elph_set_tuning", key, 0)
        x2 = np.zeros(om.Ndim)
        it2, _, f2 = E.ldiv_(x2, em, b, Pe)
        em._call("elph_set_tuning", key, 1)
        assert f2 == 0 and it2 == it1 and relerr(x2, x1) <= 1e-9, key
    em._call("elph_set_tuning", 17, 1)
    # force kernel at full size (no solve involved): <dM/dx> with the oracle's vectors
    do, de = np.zeros(om.Ndof), np.zeros(om.Ndof)
    om.muldMdx(do, g, xo)
    E.muldMdx_(de, g, em, xo)
    assert relerr(de, do) <= 1e-9


def test_E_products_64x64_L400():
    """Config E: Holstein 64x64, Ltau = 400 (1,638,400 points)."""
    import elphdynamics_b200 as E
    from oracle.cref import CRef
    om, rng = oracle_holstein("square", 64, 40.0, 0.1, mu=-1.0, seed=5, eps=0.3)
    em = engine_holstein_like(om)
    c = CRef(om)
    u, v = rng.normal(size=om.Ndim), rng.normal(size=om.Ndim)
    yc, ye = np.zeros(om.Ndim), np.zeros(om.Ndim)
    c.mulMTM(yc, v)
    E.mulMTM_(ye, em, v)
    assert relerr(ye, yc) <= 1e-12
    Mv, Mtu = np.zeros(om.Ndim), np.zeros(om.Ndim)
    E.mulM_(Mv, em, v)
    E.mulMT_(Mtu, em, u)
    assert abs(u @ Mv - Mtu @ v) <= 1e-12 * np.linalg.norm(u) * np.linalg.norm(Mv)
    for py in (8, 4):
        em._call("elph_set_tuning", 2, py)
        y2 = np.zeros(om.Ndim)
        E.mulMTM_(y2, em, v)
        assert relerr(y2, yc) <= 1e-12
    em.close()


@pytest.mark.parametrize("geom,Ls,dtau", [("honeycomb", 32, 0.1), ("triangular", 45, 0.05), ("triangular", 46, 0.05)],
                         ids=["honeycomb32", "triangular45-ragged", "triangular46"])
def test_D_hmc_lattices_at_2k_sites(geom, Ls, dtau):
    """Config D: the HMC examples scaled to ~2k sites (honeycomb L=32: 3 colours x 1024; triangular L=45: 8 ragged
    colours 1012..12; L=46: 6 x 1058), beta = 2 as shipped: products, adjointness, the HMC force pieces."""
    import elphdynamics_b200 as E
    from elphdynamics_b200 import hmc as ehmc
    from oracle import hmc as ohmc
    om, rng = oracle_holstein(geom, Ls, 2.0, dtau, mu=0.0, seed=3, eps=0.3)
    em = engine_holstein_like(om)
    assert em.group_sizes.tolist() == np.diff(om.group_offsets).tolist()
    u, v = rng.normal(size=om.Ndim), rng.normal(size=om.Ndim)
    yo, ye = np.zeros(om.Ndim), np.zeros(om.Ndim)
    for fo, fe in ((om.mulM, E.mulM_), (om.mulMT, E.mulMT_), (om.mulMTM, E.mulMTM_)):
        fo(yo, v)
        fe(ye, em, v)
        assert relerr(ye, yo) <= 1e-12, fe.__name__
    ho = ohmc.HybridMonteCarlo(om, 0.01, 1.0, 0.0, 10)
    he = ehmc.HybridMonteCarlo(em, 0.01, 1.0, 0.0, 10)
    Rp, Rm = rng.normal(size=om.Ndim), rng.normal(size=om.Ndim)
    So = ohmc.refresh_phi(ho, om, Rp, Rm)
    Se = ehmc.refresh_phi_(he, em, Rp, Rm)
    assert abs(Se - So) <= 1e-12 * abs(So)
    assert relerr(he.get("phi_plus"), ho.phip) <= 1e-12
    # force pieces with the oracle's vectors in place of the solves (no CG needed at this size)
    do, de = np.zeros(om.Ndof), np.zeros(om.Ndof)
    om.muldMdx(do, u, v)
    E.muldMdx_(de, u, em, v)
    assert relerr(de, do) <= 1e-9
    em.close()


def test_C_ssh_32x32_L200_products():
    """Config C: SSH 32x32, Ltau = 200 (phonon-modulated bonds, per-(tau,bond) tables)."""
    import elphdynamics_b200 as E
    from helpers_ssh import engine_ssh_like, oracle_ssh
    om, rng = oracle_ssh(Lside=32, beta=10.0, dtau=0.05, seed=11)
    em = engine_ssh_like(om)
    u, v = rng.normal(size=om.Ndim), rng.normal(size=om.Ndim)
    yo, ye = np.zeros(om.Ndim), np.zeros(om.Ndim)
    om.mulMTM(yo, v)
    E.mulMTM_(ye, em, v)
    assert relerr(ye, yo) <= 1e-12
    Mv, Mtu = np.zeros(om.Ndim), np.zeros(om.Ndim)
    E.mulM_(Mv, em, v)
    E.mulMT_(Mtu, em, u)
    assert abs(u @ Mv - Mtu @ v) <= 1e-12 * np.linalg.norm(u) * np.linalg.norm(Mv)
    do, de = np.zeros(om.Ndof), np.zeros(om.Ndof)
    om.muldMdx(do, u, v)
    E.muldMdx_(de, u, em, v)
    assert relerr(de, do) <= 1e-9
    em.close()


# ---- iteration counts of the full-size solves against the reference recurrences (src/IterativeSolvers.jl:198-231) -------------
# north_star: "CG solutions to the reference residual tolerance with iteration counts within +-2".  The persistent kernels do
# not run the reference's two-reduction loop literally (cg_p2p.cu: Chronopoulos-Gear, cg_pipe.cu: pipelined), so every form is
# compared with the literal loop at the sizes where rounding has the most room: ~800 (C), ~2000 (E, rough 32x32) iterations.

def _kernel_forms(em):
    """(label, tuning (key, value) pairs) of the forms of the unpreconditioned solve that apply to this handle."""
    return [("pipelined", ((10, 1), (7, -1))), ("single-reduction", ((10, 0), (7, 1))), ("two-reduction", ((10, 0), (7, 0))),
            ("graph replay", ((10, 0), (7, 0), (5, 0))), ("plain launches", ((10, 0), (7, 0), (5, 0), (3, 0)))]


def _reset_tuning(em):
    for key, val in ((10, -1), (7, -1), (5, 1), (3, 1), (13, 0), (14, 0), (11, 0)):
        em._call("elph_set_tuning", key, val)


def _check_forms(em, om, b, it_ref, x_ref, forms=None):
    import elphdynamics_b200 as E
    counts = {}
    for label, keys in (forms or _kernel_forms(em)):
        _reset_tuning(em)
        for key, val in keys:
            em._call("elph_set_tuning", key, val)
        xe = np.zeros(om.Ndim)
        it_e = E.solve_(xe, em, b)
        counts[label] = it_e
        assert abs(it_e - it_ref) <= 2, (label, it_e, it_ref, counts)
        assert em.last_eps < om.tol and relerr(xe, x_ref) <= 1e-3, (label, em.last_eps, relerr(xe, x_ref))
    _reset_tuning(em)
    return counts


def test_E_cg_iterations_64x64_L400():
    """Config E, ~2000 iterations: the literal loop (C restatement, threaded along tau) against the engine's launch-per-iteration
    forms (what one GPU runs at this size) and the multi-slice pipelined kernel (what two GPUs run on their slabs)."""
    from oracle.cref import CRef
    om, rng = oracle_holstein("square", 64, 40.0, 0.1, mu=-1.0, seed=5, eps=0.3)
    em = engine_holstein_like(om)
    try:
        b = rng.normal(size=om.Ndim)
        xc = np.zeros(om.Ndim)
        it_c, eps_c = CRef(om).cg_mt(xc, b, tol=om.tol, maxiter=om.maxiter)
        assert it_c > 1500
        forms = [("graph replay", ((10, 0), (7, 0))), ("plain launches", ((10, 0), (7, 0), (3, 0))),
                 ("pipelined, 6 slices per CTA", ((10, 1), (13, 10)))]
        _check_forms(em, om, b, it_c, xc, forms)
    finally:
        em.close()


def test_B_long_solve_all_forms():
    """32x32xL280 (the longest slab whose slices are all co-resident) with a rough field (eps = 1): a long solve through all five
    forms, the single-reduction and pipelined recurrences included."""
    from oracle.cref import CRef
    om, rng = oracle_holstein("square", 32, 28.0, 0.1, mu=-1.0, seed=1234, eps=1.0)
    em = engine_holstein_like(om)
    try:
        b = rng.normal(size=om.Ndim)
        xc = np.zeros(om.Ndim)
        it_c, eps_c = CRef(om).cg(xc, b, tol=om.tol, maxiter=om.maxiter)
        assert it_c > 1000, it_c
        counts = _check_forms(em, om, b, it_c, xc)
        print("iterations: reference", it_c, counts)
    finally:
        em.close()


def test_C_cg_iterations_ssh_32x32_L200():
    """Config C (SSH, ~800 iterations): NumPy restatement of the literal loop against the engine's forms."""
    from helpers_ssh import engine_ssh_like, oracle_ssh
    from oracle.solvers import ConjugateGradient, solve_cg
    om, rng = oracle_ssh(Lside=32, beta=10.0, dtau=0.05, seed=11)
    em = engine_ssh_like(om)
    try:
        b = rng.normal(size=om.Ndim)
        xo = np.zeros(om.Ndim)
        it_o = solve_cg(xo, om, b, ConjugateGradient(om.Ndim, tol=om.tol, maxiter=om.maxiter))
        assert it_o > 500, it_o
        _check_forms(em, om, b, it_o, xo)
    finally:
        em.close()


@pytest.mark.parametrize("geom,Ls,dtau", [("honeycomb", 32, 0.1), ("triangular", 45, 0.05)], ids=["honeycomb32", "triangular45"])
def test_D_cg_iterations(geom, Ls, dtau):
    """Config D (HMC lattices at ~2k sites): the generic persistent kernels in both forms and the launch-per-iteration forms."""
    from oracle.cref import CRef
    om, rng = oracle_holstein(geom, Ls, 2.0, dtau, mu=0.0, seed=3, eps=0.3)
    em = engine_holstein_like(om)
    try:
        b = rng.normal(size=om.Ndim)
        xc = np.zeros(om.Ndim)
        it_c, eps_c = CRef(om).cg(xc, b, tol=om.tol, maxiter=om.maxiter)
        forms = [("single-reduction", ((7, 1),)), ("two-reduction", ((7, 0),)), ("graph replay", ((7, 0), (5, 0))),
                 ("plain launches", ((7, 0), (5, 0), (3, 0)))]
        _check_forms(em, om, b, it_c, xc, forms)
    finally:
        em.close()


def test_B_sharded_kpm_pcg_against_oracle_and_engine(config_B):
    """Config B through the tau-sharded driver at world 1 (the slab = the whole lattice, the ring closes on the GPU): the
    omega-sharded application of the preconditioner in both forms (all-to-all path degenerate to copies, arena path of
    csrc/kpm_shard.cu) and ``ldiv!(x, model, b, P)`` against the oracle's iteration count (+-2) and the single-GPU engine's solution."""
    import torch
    import elphdynamics_b200 as E
    from oracle.kpm import KPMPreconditioner
    from oracle.solvers import ConjugateGradient, ldiv
    from elphdynamics_b200.sharded import CudaSlabBackend, RingComm, ShardedKPM, ShardedOperator
    om, em, rng = config_B
    noise = rng.normal(size=2 * om.N)
    g = rng.normal(size=om.Ndim)
    b = np.zeros(om.Ndim)
    om.mulMT(b, g)
    Po = KPMPreconditioner(om)
    Po.setup(noise)
    xo = np.zeros(om.Ndim)
    it_o, _, fo = ldiv(xo, om, b, ConjugateGradient(om.Ndim, tol=om.tol, maxiter=om.maxiter), Po)
    Pe = E.SymmetricKPMPreconditioner(em)
    E.setup_(Pe, noise)
    xe = np.zeros(om.Ndim)
    it_e, _, fe = E.ldiv_(xe, em, b, Pe)
    assert fo == fe == 0
    slab, aux = engine_holstein_like(om), engine_holstein_like(om)
    be = CudaSlabBackend(slab, 0, om.L)
    op = ShardedOperator(be, RingComm(0, 1), tol=om.tol, maxiter=om.maxiter)
    op.update_model()
    be.kpm_init(aux)
    P = ShardedKPM(op, om.N, om.L)
    eng = lambda a: np.ascontiguousarray(a.reshape(om.N, om.L).T)
    bt = be.empty()
    bt[1:om.L + 1] = torch.from_numpy(eng(b)).cuda()
    for fused in (False, True):
        if fused:
            assert P.enable_fused(0)
        P.setup(noise)
        assert P.active and np.array_equal(be.kpm_orders(), Po.order)
        x = be.empty()
        it_s, res_s, fs = op.ldiv(x, bt, P=P)
        assert fs == 0 and abs(it_s - it_o) <= 2 and abs(it_s - it_e) <= 1, (fused, it_s, it_o, it_e)
        assert relerr(x[1:om.L + 1].cpu().numpy(), eng(xe)) <= 1e-6, fused
    be.kpm_shard_check()
    slab.close()
    aux.close()
