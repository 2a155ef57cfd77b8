"""Shared builders for the tests: the SAME physical model as an oracle object (NumPy, CPU)
and as an engine object (libelph_b200.so, GPU)."""
from __future__ import annotations

import numpy as np

from oracle import lattice as olat
from oracle.holstein import HolsteinModel as OracleHolstein

GEOMS = {
    # name: (ndim, norbits, bond defs (0-based orbits))
    "square": (2, 1, olat.SQUARE_BONDS),
    "honeycomb": (2, 2, olat.HONEYCOMB_BONDS),
    "triangular": (2, 1, olat.TRIANGULAR_BONDS),
    "chain": (1, 1, [(0, 0, (1, 0, 0))]),
}


def synthetic_field(rng, N, L, beta, omega, lam, eps):
    """x0_i = (lam/omega^2) u + sigma n  (init_phonons_half_filled! distribution,
    src/InitializePhonons.jl:71-115), then x[i,tau] = x0_i + eps N(0,1).  Host layout."""
    sig = 1.0 / np.sqrt(2 * omega * np.tanh(beta * omega / 2))
    x0 = (lam / omega ** 2) * rng.integers(-1, 2, size=N) + sig * rng.normal(size=N)
    return (x0[:, None] + eps * rng.normal(size=(N, L))).reshape(-1)


def oracle_holstein(geom="square", Lside=4, beta=2.0, dtau=0.1, t=1.0, omega=1.0, lam=1.0, mu=-1.0, omega4=0.0, lam2=0.0,
                    tol=1e-5, maxiter=10000, seed=1234, eps=0.3):
    ndim, norb, bonds = GEOMS[geom]
    lat = olat.Lattice(ndim, norb, Lside)
    m = OracleHolstein(lat, bonds, t, beta, dtau, omega=omega, lam=lam, mu=mu, omega4=omega4, lam2=lam2, tol=tol, maxiter=maxiter)
    rng = np.random.default_rng(seed)
    m.x[:] = synthetic_field(rng, m.N, m.L, beta, omega, lam, eps)
    m.update_model()
    return m, rng


def engine_holstein_like(om, device=-1):
    """Engine model with exactly the oracle model's parameters, built through the package's own
    host-side geometry code (NOT from the oracle's tables)."""
    import elphdynamics_b200 as E
    lat = om.lat
    uc = E.UnitCell(lat.ndim, lat.norbits)
    elat = E.Lattice(uc, lat.L1, lat.L2, lat.L3)
    em = E.HolsteinModel(elat, om.beta, om.dtau, tol=om.tol, maxiter=om.maxiter, device=device)
    em.assign_omega(om.omega)
    em.assign_mu(om.mu)
    em.assign_omega4(om.omega4)
    em.assign_lambda(om.lam)
    em.assign_lambda2(om.lam2)
    off = 0
    for (o1, o2, d), cnt in zip(om.geom_defs, om.geom.def_counts):
        em.assign_t(om.t[off:off + cnt], o1, o2, d)
        off += cnt
    em.initialize_model_()
    em.x = om.x
    E.update_model_(em)
    return em


def relerr(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    d = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (d if d > 0 else 1.0))
