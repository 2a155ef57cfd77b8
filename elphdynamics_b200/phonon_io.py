"""Phonon-configuration text files of the reference: the format an existing
ElPhDynamics run leaves on disk, so the engine can start from a reference-equilibrated configuration and hand its own
configurations back.

Holstein (``write_phonons!`` / ``read_phonons!``, src/HolsteinModels.jl:764-853): header ``L3 L2 L1 orbit tau x``, then
one line ``l3 l2 l1 orbit tau x`` per (unit cell, orbital, time slice) with 0-based cell coordinates, 1-based orbit and
tau, l1 fastest among the cells, tau innermost, ``x`` printed with ``%.6f``.
SSH (src/SSHModels.jl:838-913): header ``type loc tau x``, lines ``type loc tau x`` with 1-based phonon type, 1-based
index inside the type and tau, tau innermost; nothing is written when the model has no phonons.

The text format keeps six decimals: a round trip reproduces x to 5e-7, not bit for bit -- exactly as in the reference.
Reading ends with ``update_model!`` like the reference (the device tables follow the new field).
"""
from __future__ import annotations

import numpy as np

from .models import HolsteinModel, SSHModel, update_model_


def _holstein_order(model: HolsteinModel):
    """(l3, l2, l1, orbit) of every site in file order and the matching 0-based site numbers
    (site = norbits * cell + orbit - 1, cell = l1 + L1 l2 + L1 L2 l3; src/Lattices.jl:149-168)."""
    lat = model.lattice
    nor = lat.unit_cell.norbits
    l3, l2, l1, orb = np.meshgrid(np.arange(lat.L3), np.arange(lat.L2), np.arange(lat.L1), np.arange(1, nor + 1), indexing="ij")
    l3, l2, l1, orb = (a.reshape(-1) for a in (l3, l2, l1, orb))
    site = nor * (l1 + lat.L1 * (l2 + lat.L2 * l3)) + orb - 1
    return l3, l2, l1, orb, site


def write_phonons_(model, filename: str) -> None:
    """``write_phonons!(model, filename)``."""
    L = model.Ltau
    x = model.x.reshape(-1, L)              # host layout: tau fastest, (site | phonon) slow
    if isinstance(model, HolsteinModel):
        l3, l2, l1, orb, site = _holstein_order(model)
        with open(filename, "w") as f:
            f.write("L3 L2 L1 orbit tau x\n")
            for a3, a2, a1, o, s in zip(l3.tolist(), l2.tolist(), l1.tolist(), orb.tolist(), site.tolist()):
                row = x[s]
                f.write("".join("%d %d %d %d %d %.6f\n" % (a3, a2, a1, o, t + 1, row[t]) for t in range(L)))
    elif isinstance(model, SSHModel):
        if model.Nph == 0:
            return                          # the reference writes no file at all (src/SSHModels.jl:840)
        n = model.nph
        N = model.Nph // n
        with open(filename, "w") as f:
            f.write("type loc tau x\n")
            for ph in range(n):
                for i in range(N):
                    row = x[ph * N + i]
                    f.write("".join("%d %d %d %.6f\n" % (ph + 1, i + 1, t + 1, row[t]) for t in range(L)))
    else:
        raise TypeError("write_phonons_: unknown model type")


def read_phonons_(model, filename: str) -> None:
    """``read_phonons!(model, filename)``: entries present in the file overwrite x, then ``update_model!``."""
    L = model.Ltau
    x = model.x.reshape(-1, L).copy()
    with open(filename, "r") as f:
        f.readline()                        # header
        if isinstance(model, HolsteinModel):
            lat = model.lattice
            nor = lat.unit_cell.norbits
            for line in f:
                a = line.split(" ")
                if len(a) < 6:
                    continue
                l3, l2, l1, orb, tau = (int(v) for v in a[:5])
                site = nor * (l1 + lat.L1 * (l2 + lat.L2 * l3)) + orb - 1
                x[site, tau - 1] = float(a[5])
        elif isinstance(model, SSHModel):
            N = model.Nph // model.nph if model.Nph else 0
            for line in f:
                a = line.split(" ")
                if len(a) < 4:
                    continue
                ph, i, tau = (int(v) for v in a[:3])
                x[(ph - 1) * N + (i - 1), tau - 1] = float(a[3])
        else:
            raise TypeError("read_phonons_: unknown model type")
    model.x = x.reshape(-1)
    update_model_(model)
