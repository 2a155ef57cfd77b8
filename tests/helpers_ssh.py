"""Builders for the SSH model: same physical model as an oracle object and as an engine object."""
from __future__ import annotations

import numpy as np

from oracle import lattice as olat
from oracle.ssh import SSHBondDef, SSHModel as OracleSSH

# bond definitions of the shipped examples (0-based orbits)
SSH_SQUARE = [dict(o1=0, o2=0, d=(1, 0, 0)), dict(o1=0, o2=0, d=(0, 1, 0))]   # examples/ssh_langevin_square.toml:39-80
SSH_TWO_SITE = [dict(o1=0, o2=1, d=(0, 0, 0))]                                 # examples/ssh_hmc_two_site.toml


def oracle_ssh(Lside=4, beta=2.0, dtau=0.05, t=1.0, alpha=0.1, alpha2=0.0, omega=0.1, omega4=0.0, mu=0.0, tol=1e-5,
               maxiter=10000, seed=1234, eps=0.3, geometry="square", names=None, mixed=False):
    """x0 = sigma n - 2 alpha/omega^2 per bond (src/InitializePhonons.jl:41-48), plus eps N(0,1)
    roughness in tau.  ``names``: per-definition phonon names (equal names = equivalent fields).  ``mixed``: the
    second bond definition carries no phonon (omega = 0)."""
    if geometry == "square":
        lat = olat.Lattice(2, 1, Lside)
        geo = SSH_SQUARE
    elif geometry == "two_site":
        lat = olat.Lattice(1, 2, 1)
        geo = SSH_TWO_SITE
    else:
        raise KeyError(geometry)
    defs = []
    for k, gdef in enumerate(geo):
        om = 0.0 if (mixed and k == 1) else omega
        defs.append(SSHBondDef(gdef["o1"], gdef["o2"], gdef["d"], t=t, alpha=alpha, alpha2=alpha2, omega=om, omega4=omega4,
                               name=(names[k] if names else "")))
    m = OracleSSH(lat, defs, beta, dtau, mu=mu, tol=tol, maxiter=maxiter)
    rng = np.random.default_rng(seed)
    sig = 1.0 / np.sqrt(2 * omega * np.tanh(beta * omega / 2))
    # scaled down so that |alpha x| stays well below t (physical regime; the synthetic sigma is large for omega = 0.1)
    x0 = 0.3 * sig * rng.normal(size=m.Nph) - 2 * alpha / omega ** 2 * 0.05
    X = x0[:, None] + eps * rng.normal(size=(m.Nph, m.L))
    x = X.reshape(-1)
    m.x[:] = x[m.primary_field]          # equivalent fields must be equal
    m.update_model()
    return m, rng


def engine_ssh_like(om, device=-1):
    import elphdynamics_b200 as E
    lat = om.lat
    elat = E.Lattice(E.UnitCell(lat.ndim, lat.norbits), lat.L1, lat.L2, lat.L3)
    em = E.SSHModel(elat, om.beta, om.dtau, tol=om.tol, maxiter=om.maxiter, device=device)
    em.assign_mu(om.mu)
    for bd in om.bond_defs:
        em.assign_hopping(bd.t, bd.omega, bd.omega4, bd.alpha, bd.alpha2, bd.o1, bd.o2, bd.d, bd.name)
    em.initialize_model_()
    em.x = om.x
    E.update_model_(em)
    return em
