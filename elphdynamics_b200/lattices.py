"""Host-side geometry: finite lattices, neighbour tables and the checkerboard bond order.

Mirror of the reference's integer set-up code, which STAYS ON THE HOST (SURVEY.md
section 8a row A0): ``Lattice`` (src/Lattices.jl:17-107), ``calc_neighbor_table``
(:265-316), ``sorted_neighbor_table_perm!`` (:323-340), ``checkerboard_groups``
(src/Checkerboard.jl:471-515) and ``checkerboard_order`` (:442-446).  Results are
integer tables and must be bit-exact with the reference; ``tests/`` compares them
with the oracle's literal restatement and with frozen fixtures.

Indices are 0-based here (``index_base = 0`` at the C ABI).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class UnitCell:
    """``UnitCell(ndim, norbits, lvecs, bvecs)`` (src/UnitCells.jl:10-52); only the
    integers matter on the hot path, the vectors are carried for completeness."""
    ndim: int
    norbits: int
    lvecs: list = field(default_factory=list)
    bvecs: list = field(default_factory=list)


class Lattice:
    """``Lattice(unit_cell, L1[, L2, L3])`` (src/Lattices.jl:52-135)."""

    def __init__(self, unit_cell: UnitCell, L1: int, L2: int | None = None, L3: int | None = None):
        nd = unit_cell.ndim
        if L2 is None:
            L2 = L1 if nd >= 2 else 1
        if L3 is None:
            L3 = L1 if nd >= 3 else 1
        if not (L1 >= 1 and L2 >= 1 and L3 >= 1):
            raise ValueError("lattice dimensions must be >= 1")
        self.unit_cell = unit_cell
        self.L1, self.L2, self.L3 = int(L1), int(L2), int(L3)
        self.dims = np.array([self.L1, self.L2, self.L3], dtype=np.int64)
        self.norbits = unit_cell.norbits
        self.ncells = self.L1 * self.L2 * self.L3
        self.nsites = self.ncells * self.norbits
        cells = np.arange(self.ncells, dtype=np.int64)
        # cell = l1 + l2*L1 + l3*L1*L2
        self.cell_loc = np.stack([cells % self.L1, (cells // self.L1) % self.L2, cells // (self.L1 * self.L2)])
        sites = np.arange(self.nsites, dtype=np.int64)
        self.site_to_cell = sites // self.norbits
        self.site_to_orbit = sites % self.norbits

    def site_to_site(self, isites, displacement, orbit: int):
        """Vectorised ``site_to_site`` (src/Lattices.jl:182-200): periodic displacement in unit cells."""
        loc = self.cell_loc[:, self.site_to_cell[np.asarray(isites)]] + np.asarray(displacement, dtype=np.int64)[:, None]
        loc %= self.dims[:, None]
        return self.norbits * (loc[0] + loc[1] * self.L1 + loc[2] * self.L1 * self.L2) + orbit


def calc_neighbor_table(lattice: Lattice, orbit1: int, orbit2: int, displacement, remove_duplicates: bool = True) -> np.ndarray:
    """``calc_neighbor_table`` (src/Lattices.jl:265-316); orbits 0-based.  (2, n) int64."""
    if len(displacement) != 3 or not (0 <= orbit1 < lattice.norbits and 0 <= orbit2 < lattice.norbits):
        raise ValueError("invalid bond definition")
    isites = np.arange(orbit1, lattice.nsites, lattice.norbits, dtype=np.int64)
    fsites = lattice.site_to_site(isites, displacement, orbit2)
    nt = np.stack([isites, fsites])
    if remove_duplicates and nt.shape[1] > 1:
        lo = np.minimum(nt[0], nt[1])
        hi = np.maximum(nt[0], nt[1])
        key = lo * lattice.nsites + hi
        _, first = np.unique(key, return_index=True)   # first occurrence of every unordered pair
        keep = np.zeros(nt.shape[1], dtype=bool)
        keep[first] = True
        nt = nt[:, keep]
    return nt


def sorted_neighbor_table_perm(neighbor_table: np.ndarray) -> np.ndarray:
    """``sorted_neighbor_table_perm!`` (src/Lattices.jl:323-340): orders each pair (in place) and
    returns the stable permutation sorting by (first site, second site)."""
    nt = neighbor_table
    flip = nt[0] > nt[1]
    nt[:, flip] = nt[::-1, flip]
    if nt.shape[1] == 0:
        return np.zeros(0, dtype=np.int64)
    m = int(nt.max()) + 1
    return np.argsort(m * (nt[0] + 1) + nt[1] + 1, kind="stable")


def checkerboard_groups(neighbor_table: np.ndarray) -> np.ndarray:
    """``checkerboard_groups`` (src/Checkerboard.jl:471-515), 1-based group ids.

    The reference grows one colour at a time, scanning the unassigned bonds in order and
    accepting a bond unless an earlier member of the colour shares a site with it.  That is
    the same as keeping a per-colour site-occupancy mask, which is what is done here."""
    nb = neighbor_table.shape[1]
    groups = np.zeros(nb, dtype=np.int64)
    if nb == 0:
        return groups
    a = neighbor_table[0].tolist()
    b = neighbor_table[1].tolist()
    nsites = int(neighbor_table.max()) + 1
    remaining = list(range(nb))
    g = 0
    while remaining:
        g += 1
        used = bytearray(nsites)
        rest = []
        for n in remaining:
            i, j = a[n], b[n]
            if used[i] or used[j]:
                rest.append(n)
            else:
                used[i] = used[j] = 1
                groups[n] = g
        remaining = rest
    return groups


def checkerboard_order(groups: np.ndarray) -> np.ndarray:
    """``checkerboard_order`` = stable ``sortperm`` of the group ids (src/Checkerboard.jl:442-446)."""
    return np.argsort(groups, kind="stable")


@dataclass
class CheckerboardTables:
    neighbor_table: np.ndarray         # (2, Nbonds) in checkerboard order
    inv_checkerboard_perm: np.ndarray  # column -> original bond  (perm[new_perm])
    checkerboard_perm: np.ndarray      # original bond -> column  (sortperm of the above)
    group_sizes: np.ndarray


def assemble_checkerboard(neighbor_table_unsorted: np.ndarray) -> CheckerboardTables:
    """The table assembly of ``initialize_model!`` (src/HolsteinModels.jl:484-517, src/SSHModels.jl:436-446)."""
    nt = np.array(neighbor_table_unsorted, dtype=np.int64, copy=True)
    perm = sorted_neighbor_table_perm(nt)
    nt = nt[:, perm]
    groups = checkerboard_groups(nt)
    new_perm = checkerboard_order(groups)
    nt = np.ascontiguousarray(nt[:, new_perm])
    inv = perm[new_perm]
    sizes = np.bincount(groups)[1:] if groups.size else np.zeros(0, dtype=np.int64)
    return CheckerboardTables(nt, inv, np.argsort(inv, kind="stable"), sizes)
