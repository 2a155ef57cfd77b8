"""Oracle: lattice geometry, neighbour tables and checkerboard bond colouring.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Integer-only; results
must be bit-exact with the reference.  0-based site / bond indices here.

Follows:
  * ``src/Lattices.jl:52-107``   Lattice constructor (site numbering)
  * ``src/Lattices.jl:147-200``  loc_to_cell / loc_to_site / site_to_site
  * ``src/Lattices.jl:265-316``  calc_neighbor_table (duplicate removal)
  * ``src/Lattices.jl:323-340``  sorted_neighbor_table_perm!
  * ``src/Checkerboard.jl:442-446,471-515`` checkerboard_order!/groups!
  * ``src/HolsteinModels.jl:484-517`` initialize_model! assembly
"""
from __future__ import annotations

import numpy as np


class Lattice:
    """``Lattice(unit_cell, L1, L2, L3)``, ``src/Lattices.jl:52-107``.

    Sites are numbered ``norbits*cell + orbit`` with
    ``cell = l1 + l2*L1 + l3*L1*L2`` (0-based).
    """

    def __init__(self, ndim: int, norbits: int, L1: int, L2: int | None = None, L3: int | None = None):
        # ``Lattice(unit_cell, L)`` outer constructor, src/Lattices.jl:121-135
        if L2 is None:
            L2 = L1 if ndim >= 2 else 1
        if L3 is None:
            L3 = L1 if ndim >= 3 else 1
        assert L1 >= 1 and L2 >= 1 and L3 >= 1
        self.ndim, self.norbits = ndim, norbits
        self.L1, self.L2, self.L3 = L1, L2, L3
        self.ncells = L1 * L2 * L3
        self.nsites = self.ncells * norbits
        # cell_loc[:, cell], site_to_orbit, site_to_cell (src/Lattices.jl:84-104)
        cell_loc = np.zeros((3, self.ncells), dtype=np.int64)
        site_to_orbit = np.zeros(self.nsites, dtype=np.int64)
        site_to_cell = np.zeros(self.nsites, dtype=np.int64)
        site = 0
        cell = 0
        for l3 in range(L3):
            for l2 in range(L2):
                for l1 in range(L1):
                    cell_loc[:, cell] = (l1, l2, l3)
                    for orbit in range(norbits):
                        site_to_orbit[site] = orbit
                        site_to_cell[site] = cell
                        site += 1
                    cell += 1
        self.cell_loc, self.site_to_orbit, self.site_to_cell = cell_loc, site_to_orbit, site_to_cell

    def loc_to_cell(self, l1: int, l2: int, l3: int) -> int:
        """src/Lattices.jl:147-151 with ``_pbc!`` (``mod``), :384-391."""
        return (l1 % self.L1) + (l2 % self.L2) * self.L1 + (l3 % self.L3) * self.L1 * self.L2

    def site_to_site(self, isite: int, disp, orbit: int) -> int:
        """src/Lattices.jl:182-200 (orbit 0-based)."""
        cell = self.site_to_cell[isite]
        l1, l2, l3 = (int(self.cell_loc[d, cell]) + int(disp[d]) for d in range(3))
        return self.norbits * self.loc_to_cell(l1, l2, l3) + orbit


def calc_neighbor_table(lat: Lattice, orbit1: int, orbit2: int, disp, remove_duplicates: bool = True) -> np.ndarray:
    """src/Lattices.jl:265-316.  Orbits are 0-based.  Returns int64 (2, n)."""
    assert len(disp) == 3
    assert 0 <= orbit1 < lat.norbits and 0 <= orbit2 < lat.norbits
    N = lat.nsites // lat.norbits
    nt = np.zeros((2, N), dtype=np.int64)
    cnt = 0
    for isite in range(orbit1, lat.nsites, lat.norbits):
        nt[0, cnt] = isite
        nt[1, cnt] = lat.site_to_site(isite, disp, orbit2)
        cnt += 1
    if remove_duplicates:
        # keep the first occurrence of each unordered pair (src/Lattices.jl:297-313)
        keep = np.ones(N, dtype=bool)
        seen = set()
        for i in range(N):
            a, b = int(nt[0, i]), int(nt[1, i])
            key = (a, b) if a <= b else (b, a)
            if key in seen:
                keep[i] = False
            else:
                seen.add(key)
        nt = nt[:, keep]
    return nt


def calc_neighbor_table_literal(lat: Lattice, orbit1: int, orbit2: int, disp) -> np.ndarray:
    """The O(N^2) double loop exactly as written at src/Lattices.jl:297-313."""
    nt = calc_neighbor_table(lat, orbit1, orbit2, disp, remove_duplicates=False)
    N = nt.shape[1]
    keep = np.ones(N, dtype=bool)
    for i in range(N - 1):
        if keep[i]:
            a, b = nt[0, i], nt[1, i]
            for j in range(i + 1, N):
                a2, b2 = nt[0, j], nt[1, j]
                if (a == a2 and b == b2) or (a == b2 and b == a2):
                    keep[j] = False
    return nt[:, keep]


def sorted_neighbor_table_perm(nt: np.ndarray) -> np.ndarray:
    """src/Lattices.jl:323-340.  Mutates ``nt`` in place (row0 <= row1) and
    returns the stable sort permutation.  The reference key is
    ``maximum(nt)*nt[1,:] + nt[2,:]`` on 1-based indices; the 0-based key
    ``(max+1)*(a+1) + (b+1)`` orders identically."""
    assert nt.shape[0] == 2
    if nt.shape[1] == 0:
        return np.zeros(0, dtype=np.int64)
    swap = nt[0] > nt[1]
    nt[:, swap] = nt[::-1, swap]
    m = int(nt.max()) + 1  # = maximum of the 1-based table
    vals = m * (nt[0] + 1) + (nt[1] + 1)
    return np.argsort(vals, kind="stable")


def checkerboard_groups_literal(nt: np.ndarray) -> np.ndarray:
    """src/Checkerboard.jl:471-515 exactly as written (O(Nbonds^2)).
    Returns 1-based group ids like the reference."""
    nb = nt.shape[1]
    groups = np.zeros(nb, dtype=np.int64)
    group = 0
    nassigned = 0
    while nassigned < nb:
        group += 1
        for n in range(nb):
            if groups[n] == 0:
                groups[n] = group
                nassigned += 1
                for p in range(n):
                    if groups[p] == group:
                        if (nt[0, n] == nt[0, p] or nt[1, n] == nt[1, p]
                                or nt[0, n] == nt[1, p] or nt[1, n] == nt[0, p]):
                            groups[n] = 0
                            nassigned -= 1
                            break
    return groups


def checkerboard_groups(nt: np.ndarray) -> np.ndarray:
    """Same result as :func:`checkerboard_groups_literal` in O(Nbonds*ngroups):
    a bond is rejected from the group under construction iff an EARLIER bond
    already in that group touches one of its sites, i.e. iff one of its sites is
    already occupied in this group (src/Checkerboard.jl:492-510)."""
    nb = nt.shape[1]
    nsites = int(nt.max()) + 1 if nb else 0
    groups = np.zeros(nb, dtype=np.int64)
    group = 0
    nassigned = 0
    while nassigned < nb:
        group += 1
        occupied = np.zeros(nsites, dtype=bool)
        for n in range(nb):
            if groups[n] == 0:
                a, b = nt[0, n], nt[1, n]
                if not occupied[a] and not occupied[b]:
                    groups[n] = group
                    occupied[a] = True
                    occupied[b] = True
                    nassigned += 1
    return groups


def checkerboard_order(groups: np.ndarray) -> np.ndarray:
    """``sortperm!(order, groups)`` (stable), src/Checkerboard.jl:442-446."""
    return np.argsort(groups, kind="stable")


class BondGeometry:
    """The assembled, checkerboard-ordered neighbour table and its permutations.

    Mirrors the assembly in ``initialize_model!`` (``src/HolsteinModels.jl:484-517``;
    SSH ``src/SSHModels.jl:436-446``):

      nt      = hcat(calc_neighbor_table(def) for def in TOML order)
      perm    = sorted_neighbor_table_perm!(nt);   nt = nt[:, perm]
      groups  = checkerboard_groups(nt);  new_perm = checkerboard_order(groups)
      nt      = nt[:, new_perm]
      inv_checkerboard_perm = perm[new_perm]           (column -> original bond)
      checkerboard_perm     = sortperm(inv_checkerboard_perm)  (bond -> column)
    """

    def __init__(self, lat: Lattice, bond_defs, literal: bool = False):
        """``bond_defs``: list of ``(orbit1, orbit2, (d1,d2,d3))`` with 0-based orbits."""
        tables = []
        self.def_counts = []
        for (o1, o2, d) in bond_defs:
            t = (calc_neighbor_table_literal if literal else calc_neighbor_table)(lat, o1, o2, d)
            tables.append(t)
            self.def_counts.append(t.shape[1])
        nt = np.concatenate(tables, axis=1) if tables else np.zeros((2, 0), dtype=np.int64)
        self.bond_to_definition = np.concatenate(
            [np.full(c, k, dtype=np.int64) for k, c in enumerate(self.def_counts)]) if tables else np.zeros(0, np.int64)
        self.neighbor_table_unsorted = nt.copy()
        perm = sorted_neighbor_table_perm(nt)
        nt = nt[:, perm]
        groups = (checkerboard_groups_literal if literal else checkerboard_groups)(nt)
        new_perm = checkerboard_order(groups)
        self.neighbor_table = np.ascontiguousarray(nt[:, new_perm])
        self.groups = groups[new_perm]  # 1-based group id per column, non-decreasing
        self.ngroups = int(groups.max()) if groups.size else 0
        self.inv_checkerboard_perm = perm[new_perm]
        self.checkerboard_perm = np.argsort(self.inv_checkerboard_perm, kind="stable")
        self.nbonds = nt.shape[1]
        self.nsites = lat.nsites
        # group_offsets[g]..group_offsets[g+1] = columns of colour g (0-based g)
        self.group_offsets = np.zeros(self.ngroups + 1, dtype=np.int64)
        for g in range(1, self.ngroups + 1):
            self.group_offsets[g] = self.group_offsets[g - 1] + int(np.sum(self.groups == g))


# bond definitions of the shipped examples (0-based orbits)
SQUARE_BONDS = [(0, 0, (1, 0, 0)), (0, 0, (0, 1, 0))]  # examples/holstein_langevin_square.toml:45-55
HONEYCOMB_BONDS = [(0, 1, (0, 0, 0)), (0, 1, (-1, 0, 0)), (0, 1, (0, -1, 0))]  # examples/holstein_hmc_honeycomb.toml:46-64
TRIANGULAR_BONDS = [(0, 0, (1, 0, 0)), (0, 0, (0, 1, 0)), (0, 0, (1, -1, 0))]  # examples/holstein_hmc_triangular.toml:45-60
