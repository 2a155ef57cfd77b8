"""KPM apply on 64-wide square lattices with every frequency on an 8-CTA cluster, (re | im) x 4 row strips with the strip edges
through distributed shared memory (csrc/kpm_square.cu: kpm_square_wide_kernel; tuning key 26) -- against the oracle
(src/KPMPreconditioners.jl:426-481, 606-679) and against the 2-CTA kernel, in the standalone apply and inside the preconditioned
solve."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from helpers import engine_holstein_like, oracle_holstein, relerr  # noqa: E402
from oracle.kpm import KPMPreconditioner, kpm_coefficients  # noqa: E402
from oracle.solvers import ConjugateGradient, ldiv  # noqa: E402


@pytest.mark.gpu
@pytest.mark.parametrize("beta", [2.0, 4.1])
def test_wide_cluster_chains_64x64(beta):
    import elphdynamics_b200 as E
    om, rng = oracle_holstein("square", 64, beta, 0.1, mu=-0.8, seed=5)
    em = engine_holstein_like(om)
    Po, Pe = KPMPreconditioner(om), E.SymmetricKPMPreconditioner(em)
    noise = rng.normal(size=2 * om.N)
    Po.setup(noise)
    info = E.setup_(Pe, noise)
    assert info.active == 1 and Po.active and np.array_equal(Pe.orders(), Po.order)
    assert int(Po.order.max()) >= 8                     # long enough for many sweeps per chain
    Po.lam_lo, Po.lam_hi = info.lambda_lo, info.lambda_hi
    Po.lam_avg, Po.lam_mag = (Po.lam_hi + Po.lam_lo) / 2, (Po.lam_hi - Po.lam_lo) / 2
    Po.coeff = [kpm_coefficients(int(Po.order[w]), Po.lam_lo, Po.lam_hi, Po.phis[w]) for w in range(Po.Lo2)]
    r = rng.normal(size=om.Ndim)
    zo = np.zeros(om.Ndim)
    Po.ldiv(zo, r)
    got = {}
    for wide in (1, 0):
        em._call("elph_set_tuning", 26, wide)
        z = np.zeros(om.Ndim)
        E.kpm_ldiv_(z, Pe, r)
        assert relerr(z, zo) <= 1e-11, wide
        got[wide] = z
    assert relerr(got[1], got[0]) <= 1e-13              # same arithmetic per site, only the transport of the edge rows differs
    # inside the preconditioned solve
    g = rng.normal(size=om.Ndim)
    b = np.zeros(om.Ndim)
    om.mulMT(b, g)
    xo = np.zeros(om.Ndim)
    it_o, _, fo = ldiv(xo, om, b, ConjugateGradient(om.Ndim, tol=om.tol, maxiter=om.maxiter), Po)
    for wide in (1, 0):
        em._call("elph_set_tuning", 26, wide)
        x = np.zeros(om.Ndim)
        it_e, _, fe = E.ldiv_(x, em, b, Pe)
        assert fo == fe == 0 and abs(it_e - it_o) <= 2, (wide, it_e, it_o)
        assert relerr(x, xo) <= 1e-4
    em.close()
