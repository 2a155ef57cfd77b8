"""Oracle: optical SSH model operator.  TEST INFRASTRUCTURE ONLY.

Follows ``src/SSHModels.jl``:
  * ``SSHBond`` :14-77 (``has_phonon = omega != 0 || sigma_omega != 0``)
  * ``initialize_model!`` :348-505 (bond/phonon maps, checkerboard order, primary fields)
  * ``update_model!`` :510-562   t' = t - (alpha x + sign(x) alpha2 x^2); cosh/sinh(dtau t') per (tau, column)
  * ``randn!(v, ssh)`` :567-576  (v = v[primary_field])
  * ``mulM!`` :581-640, ``mulMT!`` :646-701, ``muldMdx!`` :707-829

0-based indices.  Tables with a tau axis are stored ``(Nbonds, Ltau)`` = the C-order view of the
reference's column-major ``(Ltau, Nbonds)`` matrices.  Disorder (sigma_*) is not drawn here: pass
per-bond arrays if needed.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import checkerboard as cb
from .lattice import Lattice, calc_neighbor_table, checkerboard_groups, checkerboard_order, sorted_neighbor_table_perm


@dataclass
class SSHBondDef:
    o1: int
    o2: int
    d: tuple
    t: float = 1.0
    alpha: float = 0.0
    alpha2: float = 0.0
    omega: float = 0.0
    omega4: float = 0.0
    name: str = ""

    @property
    def has_phonon(self) -> bool:
        return self.omega != 0.0


class SSHModel:
    kind = "ssh"

    def __init__(self, lat: Lattice, bond_defs, beta: float, dtau: float, mu=0.0, tol: float = 1e-5, maxiter: int = 10000):
        self.lat = lat
        self.bond_defs = list(bond_defs)
        self.beta, self.dtau = float(beta), float(dtau)
        self.L = int(round(beta / dtau))
        self.N = lat.nsites
        self.Ndim = self.N * self.L
        mu = np.asarray(mu, dtype=np.float64)
        self.mu = np.full(self.N, float(mu)) if mu.ndim == 0 else mu.copy()
        self.expmu = np.exp(self.dtau * self.mu)
        self.tol, self.maxiter = float(tol), int(maxiter)

        # ---- initialize_model! (:348-505)
        tables, t, omega, omega4, alpha, alpha2 = [], [], [], [], [], []
        phonon_to_bond, bond_to_phonon, bond_to_definition, names = [], [], [], []
        nph_types = 0
        for i, bd in enumerate(self.bond_defs):
            nn = calc_neighbor_table(lat, bd.o1, bd.o2, bd.d)
            n_new = nn.shape[1]
            tables.append(nn)
            t += [bd.t] * n_new
            bond_to_definition += [i] * n_new
            if bd.has_phonon:
                nph_types += 1
                names.append(bd.name if bd.name else f"__unnamed_{i}")   # the reference draws randstring(5)
                omega += [bd.omega] * n_new
                omega4 += [bd.omega4] * n_new
                alpha += [bd.alpha] * n_new
                alpha2 += [bd.alpha2] * n_new
                # NOTE: assumes every definition yields n_new bonds, like the reference (:424-427)
                phonon_to_bond += list(range(i * n_new, (i + 1) * n_new))
                bond_to_phonon += list(range((nph_types - 1) * n_new, nph_types * n_new))
            else:
                bond_to_phonon += [-1] * n_new
        nt = np.concatenate(tables, axis=1) if tables else np.zeros((2, 0), dtype=np.int64)
        self.t = np.asarray(t, dtype=np.float64)
        self.omega = np.asarray(omega, dtype=np.float64)
        self.omega4 = np.asarray(omega4, dtype=np.float64)
        self.alpha = np.asarray(alpha, dtype=np.float64)
        self.alpha2 = np.asarray(alpha2, dtype=np.float64)
        self.phonon_to_bond = np.asarray(phonon_to_bond, dtype=np.int64)
        self.bond_to_phonon = np.asarray(bond_to_phonon, dtype=np.int64)
        self.bond_to_definition = np.asarray(bond_to_definition, dtype=np.int64)
        self.nph = nph_types
        perm = sorted_neighbor_table_perm(nt)
        nt = nt[:, perm]
        groups = checkerboard_groups(nt)
        new_perm = checkerboard_order(groups)
        self.neighbor_table = np.ascontiguousarray(nt[:, new_perm])
        self.inv_checkerboard_perm = perm[new_perm]
        self.checkerboard_perm = np.argsort(self.inv_checkerboard_perm, kind="stable")
        g = groups[new_perm]
        self.ngroups = int(g.max()) if g.size else 0
        self.group_offsets = np.concatenate([[0], np.cumsum(np.bincount(g)[1:])]).astype(np.int64) if g.size else np.zeros(1, np.int64)
        self.Nbonds = nt.shape[1]
        self.Nph = self.omega.size
        self.Ndof = self.Nph * self.L
        self.x = np.zeros(self.Ndof)
        L = self.L
        self.tprime = np.repeat(self.t[:, None], L, axis=1)                   # (bond, tau), original bond order
        self.cosht = np.zeros((self.Nbonds, L))
        self.sinht = np.zeros((self.Nbonds, L))
        self.cosht[self.checkerboard_perm] = np.cosh(self.dtau * self.t)[:, None]
        self.sinht[self.checkerboard_perm] = np.sinh(self.dtau * self.t)[:, None]
        # primary fields (:479-500): fields are ordered (phonon, tau); types with equal names share the first type's fields
        primary = np.arange(self.Ndof, dtype=np.int64)
        if self.nph > 0:
            per_type = self.Ndof // self.nph
            pf = primary.reshape(self.nph, per_type)
            fields = np.arange(self.Ndof, dtype=np.int64).reshape(self.nph, per_type)
            for a in range(self.nph):
                for b in range(a + 1, self.nph):
                    if names[a] == names[b] and pf[b, 0] > fields[a, 0]:
                        pf[b, :] = fields[a, :]
        self.primary_field = primary
        self.names = names
        self.v1 = np.zeros(self.Ndim)
        self.v2 = np.zeros(self.Ndim)
        self.v3 = np.zeros(self.Ndim)
        self.update_model()

    # ------------------------------------------------------------------ update
    def update_model(self):
        """src/SSHModels.jl:510-562."""
        self.expmu[:] = np.exp(self.dtau * self.mu)
        X = self.x.reshape(self.Nph, self.L)
        for ph in range(self.Nph):
            bond = self.phonon_to_bond[ph]
            col = self.checkerboard_perm[bond]
            xt = X[ph]
            v = self.alpha[ph] * xt + np.sign(xt) * self.alpha2[ph] * xt ** 2
            tp = self.t[bond] - v
            self.tprime[bond] = tp
            self.cosht[col] = np.cosh(self.dtau * tp)
            self.sinht[col] = np.sinh(self.dtau * tp)
        pf = self.primary_field
        bad = pf != np.arange(self.Ndof)
        if bad.any() and not np.allclose(self.x[bad], self.x[pf[bad]], rtol=np.sqrt(np.finfo(float).eps), atol=0):
            raise ValueError("equivalent phonon fields differ (src/SSHModels.jl:549-559)")

    def randn_map(self, v):
        """randn!(v, ssh): v = v[primary_field] (:567-576)."""
        return np.asarray(v)[self.primary_field]

    # ----------------------------------------------------------------- matvecs
    def mulM(self, y, v):
        """src/SSHModels.jl:581-640."""
        N, L = self.N, self.L
        V, Y = v.reshape(N, L), y.reshape(N, L)
        Y[:, :] = self.expmu[:, None] * np.roll(V, 1, axis=1)
        cb.checkerboard_mul(Y, self.neighbor_table, self.cosht, self.sinht, self.group_offsets)
        Y[:, 0] = V[:, 0] + Y[:, 0]
        Y[:, 1:] = V[:, 1:] - Y[:, 1:]

    def mulMT(self, y, v):
        """src/SSHModels.jl:646-701."""
        N, L = self.N, self.L
        V, Y = v.reshape(N, L), y.reshape(N, L)
        Y[:, :] = V
        cb.checkerboard_transpose_mul(Y, self.neighbor_table, self.cosht, self.sinht, self.group_offsets)
        yL = V[:, L - 1] + self.expmu * Y[:, 0]
        Y[:, :L - 1] = V[:, :L - 1] - self.expmu[:, None] * Y[:, 1:]
        Y[:, L - 1] = yL

    def mulMTM(self, y, v):
        self.mulM(self.v1, v)
        self.mulMT(y, self.v1)

    def mul(self, y, v):
        self.mulMTM(y, v)

    def muldMdx(self, dMdx, u, v):
        """src/SSHModels.jl:707-829: bond-sequential recurrence, literal bond order, vectorised over tau."""
        N, L, dt = self.N, self.L, self.dtau
        U, V = u.reshape(N, L), v.reshape(N, L)
        X = self.x.reshape(self.Nph, L)
        b = self.expmu[:, None] * np.roll(V, 1, axis=1)
        c = U.copy()
        cb.checkerboard_transpose_mul(c, self.neighbor_table, self.cosht, self.sinht, self.group_offsets)
        acc = np.zeros(self.Ndof)
        A = acc.reshape(self.Nph, L)
        PF = self.primary_field.reshape(self.Nph, L)
        for n in range(self.Nbonds):
            bond = self.inv_checkerboard_perm[n]
            ph = self.bond_to_phonon[bond]
            i, j = self.neighbor_table[0, n], self.neighbor_table[1, n]
            ch, sh = self.cosht[n], self.sinht[n]
            bi, bj = b[i].copy(), b[j].copy()
            b[i] = ch * bi + sh * bj
            b[j] = ch * bj + sh * bi
            ci, cj = c[i].copy(), c[j].copy()
            c[i] = ch * ci - sh * cj
            c[j] = ch * cj - sh * ci
            if ph >= 0:
                dK = self.alpha[ph] + 2 * self.alpha2[ph] * X[ph]
                dm = c[j] * dt * dK * b[i] + (c[i] * dt * dK) * b[j]
                dm[0] = -dm[0]
                np.add.at(acc, PF[ph], dm)
        dMdx[:] = acc[self.primary_field]

    # ------------------------------------------------------------- dense debug
    def construct_M(self):
        n = self.Ndim
        M = np.zeros((n, n))
        col = np.zeros(n)
        for k in range(n):
            e = np.zeros(n)
            e[k] = 1.0
            self.mulM(col, e)
            M[:, k] = col
        return M
