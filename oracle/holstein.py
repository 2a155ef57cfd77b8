"""Oracle: Holstein model operator.  TEST INFRASTRUCTURE ONLY.

Follows ``src/HolsteinModels.jl``:
  * ``HolsteinModel`` constructor :196-313 (``Ltau = round(beta/dtau)``)
  * ``assign_t!`` :420-444, ``initialize_model!`` :484-517
  * ``update_model!`` :526-549   expnDtauV = exp(-dtau*(lam*x + lam2*x^2 - mu))
  * ``mulM!`` :569-626, ``mulMT!`` :631-684, ``muldMdx!`` :691-755
and ``src/Models.jl``: ``mulMTM!`` :215-224, ``construct_M`` :300-341.

Vectors are 1-D float64 of length N*Ltau in the reference host layout
``index = site*Ltau + tau`` (0-based; ``src/Utilities.jl:12-15``).
"""
from __future__ import annotations

import numpy as np

from . import checkerboard as cb
from .lattice import BondGeometry, Lattice


class HolsteinModel:
    kind = "holstein"

    def __init__(self, lat: Lattice, bond_defs, t, beta: float, dtau: float,
                 omega=1.0, lam=1.0, mu=0.0, omega4=0.0, lam2=0.0,
                 tol: float = 1e-5, maxiter: int = 10000):
        """``t``: scalar, one value per bond definition, or one per bond in
        definition (TOML) order -- the order ``assign_t!`` appends them."""
        self.lat = lat
        self.beta, self.dtau = float(beta), float(dtau)
        self.L = int(round(beta / dtau))          # src/HolsteinModels.jl:205
        self.N = lat.nsites
        self.Nph = self.N
        self.Ndim = self.N * self.L
        self.Ndof = self.Ndim
        self.geom_defs = list(bond_defs)
        self.geom = BondGeometry(lat, bond_defs)
        self.Nbonds = self.geom.nbonds
        t = np.atleast_1d(np.asarray(t, dtype=np.float64))
        if t.size == 1:
            t = np.full(self.Nbonds, t[0])
        elif t.size == len(bond_defs) and t.size != self.Nbonds:
            t = np.concatenate([np.full(c, tv) for tv, c in zip(t, self.geom.def_counts)])
        assert t.size == self.Nbonds
        self.t = t                                   # definition order
        # initialize_model!: cosh/sinh in definition order, then [perm][new_perm]
        order = self.geom.inv_checkerboard_perm      # column -> original bond
        self.cosht = np.cosh(self.dtau * t)[order]
        self.sinht = np.sinh(self.dtau * t)[order]
        self.neighbor_table = self.geom.neighbor_table
        self.group_offsets = self.geom.group_offsets
        self.checkerboard_perm = self.geom.checkerboard_perm

        def per_site(a):
            a = np.asarray(a, dtype=np.float64)
            return np.full(self.N, float(a)) if a.ndim == 0 else a.copy()
        self.omega, self.lam, self.mu = per_site(omega), per_site(lam), per_site(mu)
        self.omega4, self.lam2 = per_site(omega4), per_site(lam2)
        self.x = np.zeros(self.Ndof)
        self.expnV = np.zeros(self.Ndim)
        self.tol, self.maxiter = float(tol), int(maxiter)
        # scratch owned by the model in the reference (v', v'', v''')
        self.v1 = np.zeros(self.Ndim)
        self.v2 = np.zeros(self.Ndim)
        self.v3 = np.zeros(self.Ndim)
        self.update_model()

    # ------------------------------------------------------------------ update
    def update_model(self):
        """src/HolsteinModels.jl:526-549."""
        x = self.x.reshape(self.N, self.L)
        e = np.exp(-self.dtau * (self.lam[:, None] * x + self.lam2[:, None] * x ** 2 + -self.mu[:, None]))
        self.expnV[:] = e.reshape(-1)

    # ----------------------------------------------------------------- matvecs
    def mulM(self, y, v):
        """y = M v, src/HolsteinModels.jl:569-626.  Requires y is not v."""
        N, L = self.N, self.L
        V = v.reshape(N, L)
        Y = y.reshape(N, L)
        E = self.expnV.reshape(N, L)
        Y[:, :] = E * np.roll(V, 1, axis=1)           # y(tau) = eV(tau) v(tau-1), mod1 wrap
        cb.checkerboard_mul(Y, self.neighbor_table, self.cosht, self.sinht, self.group_offsets)
        Y[:, 0] = V[:, 0] + Y[:, 0]
        Y[:, 1:] = V[:, 1:] - Y[:, 1:]

    def mulMT(self, y, v):
        """y = M^T v, src/HolsteinModels.jl:631-684."""
        N, L = self.N, self.L
        V = v.reshape(N, L)
        Y = y.reshape(N, L)
        E = self.expnV.reshape(N, L)
        Y[:, :] = V
        cb.checkerboard_transpose_mul(Y, self.neighbor_table, self.cosht, self.sinht, self.group_offsets)
        yL = V[:, L - 1] + E[:, 0] * Y[:, 0]
        Y[:, :L - 1] = V[:, :L - 1] - E[:, 1:] * Y[:, 1:]
        Y[:, L - 1] = yL

    def mulMTM(self, y, v):
        """src/Models.jl:215-224 (scratch = model.v')."""
        self.mulM(self.v1, v)
        self.mulMT(y, self.v1)

    def mul(self, y, v):
        """``mul!`` with ``mul_by_M=false, transposed=false`` (CG), src/Models.jl:192-209."""
        self.mulMTM(y, v)

    def muldMdx(self, dMdx, u, v):
        """src/HolsteinModels.jl:691-755."""
        N, L, dt = self.N, self.L, self.dtau
        U = u.reshape(N, L)
        V = v.reshape(N, L)
        X = self.x.reshape(N, L)
        E = self.expnV.reshape(N, L)
        D = dMdx.reshape(N, L)
        lam, lam2 = self.lam[:, None], self.lam2[:, None]
        D[:, 0] = -dt * (self.lam + 2 * self.lam2 * X[:, 0]) * E[:, 0] * V[:, L - 1]
        D[:, 1:] = dt * (lam + 2 * lam2 * X[:, 1:]) * E[:, 1:] * V[:, :L - 1]
        Y = self.v1.reshape(N, L)
        Y[:, :] = U
        cb.checkerboard_transpose_mul(Y, self.neighbor_table, self.cosht, self.sinht, self.group_offsets)
        D[:, :] = Y * D

    # ------------------------------------------------------------- dense debug
    def construct_M(self):
        """Dense N*L x N*L matrix, column by column through ``mulM`` like
        ``construct_M`` (src/Models.jl:300-341)."""
        n = self.Ndim
        M = np.zeros((n, n))
        col = np.zeros(n)
        for c in range(n):
            e = np.zeros(n)
            e[c] = 1.0
            self.mulM(col, e)
            M[:, c] = col
        return M

    def construct_M_blocks(self):
        """Independent dense construction from the block structure documented at
        src/HolsteinModels.jl:575-589: I on the diagonal, -B(tau) on the
        sub-diagonal, +B(1) in the top-right corner, B(tau) = K * diag(eV(tau))
        with K = ``checkerboard_matrix`` (src/Checkerboard.jl:10-49)."""
        N, L = self.N, self.L
        K = cb.checkerboard_matrix(self.neighbor_table, self.cosht, self.sinht, N)
        E = self.expnV.reshape(N, L)
        M = np.eye(N * L)
        idx = lambda site, tau: site * L + tau
        for tau in range(L):
            B = K * E[:, tau][None, :]
            rows = np.array([idx(i, tau) for i in range(N)])
            cols = np.array([idx(i, (tau - 1) % L) for i in range(N)])
            sgn = +1.0 if tau == 0 else -1.0
            M[np.ix_(rows, cols)] += sgn * B
        return M
