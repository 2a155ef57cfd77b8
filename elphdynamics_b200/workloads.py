"""Synthetic workloads of SURVEY.md section 8(d): the shipped example geometries and parameters with a
seeded synthetic phonon field, built through the package's own host-side API.

The field follows the distribution of ``init_phonons_half_filled!`` (src/InitializePhonons.jl:71-115)
plus i.i.d. roughness in imaginary time; SSH follows src/InitializePhonons.jl:41-48.  The same seeds
and draw order are used by ``tests/helpers.py`` for the oracle, so that engine, oracle and CPU baseline
consume identical bytes.  Nothing here touches ``oracle/``.
"""
from __future__ import annotations

import numpy as np

from .lattices import Lattice, UnitCell
from .models import HolsteinModel, SSHModel, update_model_

# (ndim, norbits, bond definitions (orbit1, orbit2, displacement)), 0-based orbits
GEOMETRIES = {
    "square": (2, 1, [(0, 0, (1, 0, 0)), (0, 0, (0, 1, 0))]),                         # examples/holstein_langevin_square.toml:45-55
    "honeycomb": (2, 2, [(0, 1, (0, 0, 0)), (0, 1, (-1, 0, 0)), (0, 1, (0, -1, 0))]),  # examples/holstein_hmc_honeycomb.toml:46-64
    "triangular": (2, 1, [(0, 0, (1, 0, 0)), (0, 0, (0, 1, 0)), (0, 0, (1, -1, 0))]),  # examples/holstein_hmc_triangular.toml:45-60
    "chain": (1, 1, [(0, 0, (1, 0, 0))]),
}

# the named configurations of SURVEY.md section 8
CONFIGS = {
    "A": dict(geom="square", Lside=4, beta=2.0, dtau=0.1),
    "B": dict(geom="square", Lside=32, beta=20.0, dtau=0.1),
    "D_honeycomb": dict(geom="honeycomb", Lside=32, beta=2.0, dtau=0.1),
    "D_triangular": dict(geom="triangular", Lside=45, beta=2.0, dtau=0.05),
    "E": dict(geom="square", Lside=64, beta=40.0, dtau=0.1),
}


def synthetic_field(rng, N, L, beta, omega, lam, eps):
    """x0_i = (lam/omega^2) u + sigma n, u in {-1,0,1}; x[i,tau] = x0_i + eps N(0,1).  Host layout (tau fastest)."""
    sig = 1.0 / np.sqrt(2 * omega * np.tanh(beta * omega / 2))
    x0 = (lam / omega ** 2) * rng.integers(-1, 2, size=N) + sig * rng.normal(size=N)
    return (x0[:, None] + eps * rng.normal(size=(N, L))).reshape(-1)


def holstein(geom="square", Lside=4, beta=2.0, dtau=0.1, t=1.0, omega=1.0, lam=1.0, mu=-1.0, omega4=0.0, lam2=0.0,
             tol=1e-5, maxiter=10000, seed=1234, eps=0.3, device=-1):
    """Holstein model on a shipped geometry with the seeded synthetic field; returns (model, rng)."""
    ndim, norb, bonds = GEOMETRIES[geom]
    lat = Lattice(UnitCell(ndim, norb), Lside)
    m = HolsteinModel(lat, beta, dtau, tol=tol, maxiter=maxiter, device=device)
    m.assign_omega(omega)
    m.assign_mu(mu)
    m.assign_omega4(omega4)
    m.assign_lambda(lam)
    m.assign_lambda2(lam2)
    for o1, o2, d in bonds:
        m.assign_t(t, o1, o2, d)
    m.initialize_model_()
    rng = np.random.default_rng(seed)
    m.x = synthetic_field(rng, m.Nsites, m.Ltau, beta, omega, lam, eps)
    update_model_(m)
    return m, rng


def ssh_square(Lside=4, beta=2.0, dtau=0.05, t=1.0, alpha=0.1, alpha2=0.0, omega=0.1, omega4=0.0, mu=0.0, tol=1e-5,
               maxiter=10000, seed=1234, eps=0.3, device=-1):
    """SSH model on the square lattice, one phonon per bond (examples/ssh_langevin_square.toml:39-101)."""
    lat = Lattice(UnitCell(2, 1), Lside)
    m = SSHModel(lat, beta, dtau, tol=tol, maxiter=maxiter, device=device)
    m.assign_mu(mu)
    for d in ((1, 0, 0), (0, 1, 0)):
        m.assign_hopping(t, omega, omega4, alpha, alpha2, 0, 0, d, "")
    m.initialize_model_()
    rng = np.random.default_rng(seed)
    sig = 1.0 / np.sqrt(2 * omega * np.tanh(beta * omega / 2))
    # scaled down so that |alpha x| stays well below t (the synthetic sigma is large for omega = 0.1)
    x0 = 0.3 * sig * rng.normal(size=m.Nph) - 2 * alpha / omega ** 2 * 0.05
    X = x0[:, None] + eps * rng.normal(size=(m.Nph, m.Ltau))
    m.x = X.reshape(-1)
    update_model_(m)
    return m, rng


def config(name: str, **overrides):
    """One of the named configurations A, B, D_honeycomb, D_triangular, E (Holstein) or C (SSH)."""
    if name == "C":
        kw = dict(Lside=32, beta=10.0, dtau=0.05)
        kw.update(overrides)
        return ssh_square(**kw)
    kw = dict(CONFIGS[name])
    kw.update(overrides)
    return holstein(**kw)
