"""CPU tests that PIN the oracle: the reference ships no tests or golden vectors, so the
restatement is pinned by the invariants its own debug constructors define."""
import numpy as np
import pytest

from helpers import oracle_holstein
from oracle import checkerboard as cb
from oracle import lattice as olat
from oracle.action import calc_dSbdx, calc_Sb
from oracle.fourier import FourierAccelerator, TimeFreqFFT
from oracle.holstein import HolsteinModel
from oracle.kpm import KPMPreconditioner, kpm_coefficients
from oracle.solvers import ConjugateGradient, ldiv, solve_cg, solve_pcg

GEOM_CASES = [("square", 4), ("square", 3), ("honeycomb", 3), ("triangular", 3), ("chain", 5), ("square", 2)]


@pytest.fixture(params=GEOM_CASES, ids=lambda c: f"{c[0]}{c[1]}")
def model(request):
    geom, Ls = request.param
    m, rng = oracle_holstein(geom, Ls, beta=0.8, dtau=0.1, mu=-0.4, lam2=0.1, omega4=0.05)
    return m, rng


def test_dense_M_two_constructions_agree(model):
    """(1) matvec vs the dense construct_M-style matrix (src/Models.jl:300-341) and vs the block
    structure documented at src/HolsteinModels.jl:575-589."""
    m, _ = model
    M1 = m.construct_M()
    M2 = m.construct_M_blocks()
    assert np.abs(M1 - M2).max() < 5e-15


def test_MT_and_MTM_against_dense(model):
    m, rng = model
    M = m.construct_M()
    v = rng.normal(size=m.Ndim)
    y = np.zeros(m.Ndim)
    m.mulMT(y, v)
    assert np.abs(y - M.T @ v).max() < 1e-13
    m.mulMTM(y, v)
    assert np.abs(y - M.T @ (M @ v)).max() < 1e-13


def test_adjointness(model):
    """(2) <u, M v> = <M^T u, v>."""
    m, rng = model
    u, v = rng.normal(size=m.Ndim), rng.normal(size=m.Ndim)
    a, b = np.zeros(m.Ndim), np.zeros(m.Ndim)
    m.mulM(a, v)
    m.mulMT(b, u)
    assert abs(u @ a - b @ v) < 1e-12 * np.linalg.norm(u) * np.linalg.norm(a)


def test_checkerboard_literal_equals_grouped_and_inverse(model):
    """(3) inverse o forward = I; grouped sweeps are bit-identical to the bond-by-bond reference loops."""
    m, rng = model
    nt, c, s, off = m.neighbor_table, m.cosht, m.sinht, m.group_offsets
    for lit, grp in ((cb.checkerboard_mul_literal, cb.checkerboard_mul),
                     (cb.checkerboard_transpose_mul_literal, cb.checkerboard_transpose_mul),
                     (cb.checkerboard_inverse_mul_literal, cb.checkerboard_inverse_mul),
                     (cb.checkerboard_inverse_transpose_mul_literal, cb.checkerboard_inverse_transpose_mul)):
        Y1 = rng.normal(size=(m.N, m.L))
        Y2 = Y1.copy()
        lit(Y1, nt, c, s)
        grp(Y2, nt, c, s, off)
        assert np.array_equal(Y1, Y2)
    Y = rng.normal(size=(m.N, m.L))
    Z = Y.copy()
    cb.checkerboard_mul(Z, nt, c, s, off)
    cb.checkerboard_inverse_mul(Z, nt, c, s, off)
    assert np.abs(Z - Y).max() < 1e-13
    K = cb.checkerboard_matrix(nt, c, s, m.N)
    Kt = cb.checkerboard_matrix(nt, c, s, m.N, transposed=True)
    assert np.abs(K.T - Kt).max() < 1e-15
    # per-tau tables (SSH shape) reduce to the shared-table sweep when constant in tau
    c2 = np.repeat(c[:, None], m.L, axis=1)
    s2 = np.repeat(s[:, None], m.L, axis=1)
    A, B = Y.copy(), Y.copy()
    cb.checkerboard_mul(A, nt, c, s, off)
    cb.checkerboard_mul(B, nt, c2, s2, off)
    assert np.array_equal(A, B)


def test_frequency_block_identity():
    """(4) for a tau-independent field, F Theta M v = (I - e^{-i phi_w} A) F Theta v blockwise,
    phi_w = 2 pi (w + 1/2)/L  (src/TimeFreqFFTs.jl:37,55-73; src/KPMPreconditioners.jl:117,948-951)."""
    m, rng = oracle_holstein("square", 4, beta=1.2, dtau=0.1, eps=0.0)
    P = KPMPreconditioner(m)
    P.update_A()
    A = P.construct_Abar()
    fft = TimeFreqFFT(m.N, m.L)
    v = rng.normal(size=m.Ndim)
    Mv = np.zeros(m.Ndim)
    m.mulM(Mv, v)
    lhs = fft.tau_to_omega(Mv).reshape(m.N, m.L)
    nu = fft.tau_to_omega(v).reshape(m.N, m.L)
    for w in range(m.L):
        phi = 2 * np.pi * (w + 0.5) / m.L
        rhs = nu[:, w] - np.exp(-1j * phi) * (A @ nu[:, w])
        assert np.abs(lhs[:, w] - rhs).max() < 1e-12


def test_force_matches_finite_differences():
    """(5) muldMdx = u^T (dM/dx) v by central differences (src/HolsteinModels.jl:691-755)."""
    m, rng = oracle_holstein("square", 3, beta=0.5, dtau=0.1, lam2=0.07)
    u, v = rng.normal(size=m.Ndim), rng.normal(size=m.Ndim)
    d = np.zeros(m.Ndof)
    m.muldMdx(d, u, v)
    x0 = m.x.copy()
    y = np.zeros(m.Ndim)
    h = 1e-5
    for k in rng.choice(m.Ndof, size=12, replace=False):
        vals = []
        for sgn in (+1, -1):
            m.x[:] = x0
            m.x[k] += sgn * h
            m.update_model()
            m.mulM(y, v)
            vals.append(u @ y)
        fd = (vals[0] - vals[1]) / (2 * h)
        assert abs(fd - d[k]) < 1e-6 * max(1.0, abs(d[k]))
    m.x[:] = x0
    m.update_model()


def test_single_site_closed_form():
    """(6) examples/holstein_hmc_single_site.toml: N=1, no bonds -> det M = 1 + prod_tau exp(-dtau V_tau)."""
    lat = olat.Lattice(1, 1, 1)
    m = HolsteinModel(lat, [], np.zeros(0), beta=1.0, dtau=0.1, omega=1.0, lam=1.0, mu=0.0)
    rng = np.random.default_rng(3)
    m.x[:] = rng.normal(size=m.Ndof)
    m.update_model()
    M = m.construct_M()
    assert abs(np.linalg.det(M) - (1 + np.prod(m.expnV))) < 1e-12


def test_tau_fft_roundtrip_and_conventions():
    """(7) omega_to_tau o tau_to_omega = id; forward convention sum_j x_j e^{-2 pi i jk/L}."""
    rng = np.random.default_rng(0)
    N, L = 5, 12
    fft = TimeFreqFFT(N, L)
    v = rng.normal(size=N * L)
    nu = fft.tau_to_omega(v)
    assert np.abs(fft.omega_to_tau_real(nu) - v).max() < 1e-14
    k = 3
    direct = sum(np.exp(-1j * np.pi * t / L) * v.reshape(N, L)[2, t] * np.exp(-2j * np.pi * t * k / L) for t in range(L))
    assert abs(nu.reshape(N, L)[2, k] - direct) < 1e-13


def test_cg_residuals_and_flags(model):
    """(9) CG true residual <= sqrt(tol) and equals the recursive residual to rounding; flag semantics."""
    m, rng = model
    g = rng.normal(size=m.Ndim)
    b = np.zeros(m.Ndim)
    m.mulMT(b, g)
    cg = ConjugateGradient(m.Ndim, tol=1e-8, maxiter=10000)
    x = np.zeros(m.Ndim)
    it, resid, flag = ldiv(x, m, b, cg)
    assert flag == 0 and resid <= 1e-4 and it > 0
    assert abs(resid - cg.history[-1]) < 1e-9
    M = m.construct_M()
    assert np.abs(M @ x - g).max() < 1e-5
    # maxiter hit -> flag 1 and x zeroed (src/Models.jl:156-166)
    cg2 = ConjugateGradient(m.Ndim, tol=1e-12, maxiter=2)
    x = np.zeros(m.Ndim)
    it, resid, flag = ldiv(x, m, b, cg2)
    if resid > 1e-6:
        assert flag == 1 and it == 2 and not x.any()


def test_kpm_preconditioner_is_a_good_inverse_and_speeds_up_cg():
    m, rng = oracle_holstein("square", 4, beta=2.0, dtau=0.1, mu=-1.0)
    cg = ConjugateGradient(m.Ndim, tol=1e-5, maxiter=10000)
    g = rng.normal(size=m.Ndim)
    b = np.zeros(m.Ndim)
    m.mulMT(b, g)
    x = np.zeros(m.Ndim)
    it_plain = solve_cg(x, m, b, cg)
    P = KPMPreconditioner(m)
    P.setup(rng.normal(size=2 * m.N))
    assert P.active and 0 < P.e_min < 1 < P.e_max
    x = np.zeros(m.Ndim)
    it_pre = solve_pcg(x, m, b, cg, P)
    assert it_pre < it_plain / 3
    # symmetric positive: <r, P r> > 0 and <a, P b> = <P a, b>
    a, c = rng.normal(size=m.Ndim), rng.normal(size=m.Ndim)
    Pa, Pc = np.zeros(m.Ndim), np.zeros(m.Ndim)
    P.ldiv(Pa, a)
    P.ldiv(Pc, c)
    assert a @ Pa > 0 and abs(a @ Pc - Pa @ c) < 1e-10 * abs(a @ Pc)


def test_kpm_coefficients_match_scipy_dct_and_approximate_the_inverse():
    """kpm_coefficients! (src/KPMPreconditioners.jl:789-839) through the orthonormal DCT-II the reference
    calls (FFTW.dct!), and the Chebyshev sum reproduces 1/(1 - e^{-i phi} x) on [lam_lo, lam_hi]."""
    from scipy.fft import dct
    order, lo, hi, phi = 40, 0.5, 1.6, 2 * np.pi * 2.5 / 200
    c = kpm_coefficients(order, lo, hi, phi)
    M, NM = order, 2 * order
    avg, mag = (hi + lo) / 2, (hi - lo) / 2
    n = np.arange(NM)
    f = 1.0 / (1.0 - np.exp(-1j * phi) * (mag * np.cos(np.pi * (n + 0.5) / NM) + avg))
    ref = np.zeros(M, dtype=complex)
    for part, unit in ((f.real, 1.0), (f.imag, 1j)):
        cp = dct(part, type=2, norm="ortho")
        cp = cp * np.sqrt(2 * NM) / 2
        cp[0] *= np.sqrt(2)
        q = np.where(np.arange(M) == 0, np.pi, np.pi / 2)
        ref += unit * (np.pi * cp[:M]) / (NM * q)
    assert np.abs(c - ref).max() < 1e-13
    xs = np.linspace(lo + 0.05, hi - 0.05, 7)
    T = np.cos(np.outer(np.arange(M), np.arccos((xs - avg) / mag)))
    approx = c @ T
    exact = 1.0 / (1.0 - np.exp(-1j * phi) * xs)
    assert np.abs(approx - exact).max() < 0.3 * np.abs(exact).max()   # finite order: a preconditioner, not an inverse


def test_bosonic_gradient_matches_finite_differences():
    m, rng = oracle_holstein("square", 3, beta=0.6, dtau=0.1, omega4=0.1)
    for shifted in (False, True):
        d = np.zeros(m.Ndof)
        calc_dSbdx(d, m, shifted)
        x0 = m.x.copy()
        h = 1e-5
        for k in rng.choice(m.Ndof, size=8, replace=False):
            m.x[:] = x0
            m.x[k] += h
            sp = calc_Sb(m, shifted)
            m.x[k] -= 2 * h
            sm = calc_Sb(m, shifted)
            assert abs((sp - sm) / (2 * h) - d[k]) < 1e-6 * max(1.0, abs(d[k]))
        m.x[:] = x0


def test_fourier_acceleration_is_diagonal_in_frequency():
    rng = np.random.default_rng(5)
    N, L, dtau = 3, 10, 0.1
    fa = FourierAccelerator(N, L, dtau, np.full(N, 1.0))
    fa.update_Q(0.0, 10.0, 1.0)
    v = rng.normal(size=N * L)
    a = fa.accelerate(fa.accelerate(v, 0.5), 0.5)
    assert np.abs(a - fa.accelerate(v, 1.0)).max() < 1e-13
    assert np.abs(fa.accelerate(fa.accelerate(v, 1.0), -1.0) - v).max() < 1e-12
    assert fa.Q.reshape(N, L)[0, 0] > fa.Q.reshape(N, L)[0, L // 2] >= 1.0 - 1e-12   # slow modes accelerated most


def test_greens_convolution_is_the_antiperiodic_correlation_sum():
    """convolve! (src/GreensFunctions.jl:361-414) restated with numpy.fft equals the definition it implements,
    (a * b)[D, s2, s1] = (1/V) sum_i a(i + D, s2) b(i, s1) over the doubled time axis and the cells, V = 2 L N / n_s;
    with antiperiodic_copy! inputs this is the antiperiodic G(tau + beta) = -G(tau)."""
    from oracle import greens as og
    om, rng = oracle_holstein("honeycomb", 3, 0.4, 0.1)
    Gr = og.EstimateGreensFunction(om, 2)
    L, ns, L1, L2, L3 = Gr.L, Gr.ns, Gr.L1, Gr.L2, Gr.L3
    x, y = rng.normal(size=om.Ndim), rng.normal(size=om.Ndim)
    a = og.antiperiodic_copy(x, L).reshape(L3, L2, L1, ns, 2 * L)
    b = og.antiperiodic_copy(y, L).reshape(L3, L2, L1, ns, 2 * L)
    c = og.convolve(a, b, Gr)
    V = 2 * L * Gr.N / ns
    for (d3, d2, d1, s1, s2, dt) in ((0, 1, 2, 0, 1, 3), (0, 2, 0, 1, 1, 0), (0, 0, 1, 1, 0, 7)):
        tot = 0.0
        for k2 in range(L2):
            for k1 in range(L1):
                for t in range(2 * L):
                    tot += a[0, (k2 + d2) % L2, (k1 + d1) % L1, s2, (t + dt) % (2 * L)] * b[0, k2, k1, s1, t]
        assert abs(c[d3, d2, d1, s1, s2, dt] - tot / V) <= 1e-13
        # antiperiodicity in the time displacement
        assert abs(c[d3, d2, d1, s1, s2, dt] + c[d3, d2, d1, s1, s2, (dt + L) % (2 * L)]) <= 1e-13
    # periodic_product!: both halves equal
    z = og.periodic_product(x, y, L)
    assert np.array_equal(z[:, :L], z[:, L:]) and np.array_equal(z[:, :L], (x * y).reshape(-1, L))
