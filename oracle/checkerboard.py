"""Oracle: checkerboard (bond-colour) sweeps.  TEST INFRASTRUCTURE ONLY.

Follows ``src/Checkerboard.jl``:
  * ``checkerboard_mul!``            :57-83 (shared c,s + Ltau), :86-121 (per-tau
    c,s matrices), :123-141 (single slice)
  * ``checkerboard_transpose_mul!``  :149-175, :177-210, :212-230 (reverse bond order)
  * ``checkerboard_inverse_mul!``    :238-264, :266-296, :298-316 (reverse order, -s)
  * ``checkerboard_inverse_transpose_mul!`` :324-436 (forward order, -s)
  * ``checkerboard_matrix``          :10-49 (dense debug constructor)

Data layout: ``y`` has shape ``(N, ...)`` with the site index FIRST; trailing axes
(tau, or nothing for a single slice) are carried along, which is exactly the
reference's inner ``@simd for tau`` loop over ``idx = (site-1)*Ltau + tau``.
``c``/``s`` have shape ``(Nbonds,)`` or ``(Nbonds, Ltau)`` (the reference's
``(Ltau, Nbonds)`` column-major matrix, transposed to C order).

The ``*_literal`` functions loop over bonds one at a time exactly like the
reference.  The grouped variants apply one colour group at a time with fancy
indexing; because bonds inside a group touch disjoint sites the floating-point
operations per element are identical, so both are bit-identical (tested).
"""
from __future__ import annotations

import numpy as np


def _sweep_literal(y, nt, c, s, order, sign):
    for n in order:
        i, j = int(nt[0, n]), int(nt[1, n])
        t1 = y[i].copy()
        t2 = y[j].copy()
        cn, sn = c[n], s[n]
        if sign > 0:
            y[i] = cn * t1 + sn * t2
            y[j] = cn * t2 + np.conj(sn) * t1
        else:
            y[i] = cn * t1 - sn * t2
            y[j] = cn * t2 - np.conj(sn) * t1


def checkerboard_mul_literal(y, nt, c, s):
    """src/Checkerboard.jl:57-141, bond 1 first."""
    _sweep_literal(y, nt, c, s, range(nt.shape[1]), +1)


def checkerboard_transpose_mul_literal(y, nt, c, s):
    """src/Checkerboard.jl:149-230, last bond first."""
    _sweep_literal(y, nt, c, s, range(nt.shape[1] - 1, -1, -1), +1)


def checkerboard_inverse_mul_literal(y, nt, c, s):
    """src/Checkerboard.jl:238-316, last bond first with -s."""
    _sweep_literal(y, nt, c, s, range(nt.shape[1] - 1, -1, -1), -1)


def checkerboard_inverse_transpose_mul_literal(y, nt, c, s):
    """src/Checkerboard.jl:324-436, bond 1 first with -s."""
    _sweep_literal(y, nt, c, s, range(nt.shape[1]), -1)


def _bcast(a, y):
    """Broadcast a per-bond (nb,) or (nb, L) table against y[idx] of shape (nb, ...)."""
    if a.ndim == 1 and y.ndim > 1:
        return a.reshape((-1,) + (1,) * (y.ndim - 1))
    return a


def _sweep_grouped(y, nt, c, s, offsets, reverse, sign):
    ng = len(offsets) - 1
    gs = range(ng - 1, -1, -1) if reverse else range(ng)
    for g in gs:
        lo, hi = int(offsets[g]), int(offsets[g + 1])
        i = nt[0, lo:hi]
        j = nt[1, lo:hi]
        t1 = y[i]
        t2 = y[j]
        cg = _bcast(c[lo:hi], t1)
        sg = _bcast(s[lo:hi], t1)
        if sign > 0:
            y[i] = cg * t1 + sg * t2
            y[j] = cg * t2 + np.conj(sg) * t1
        else:
            y[i] = cg * t1 - sg * t2
            y[j] = cg * t2 - np.conj(sg) * t1


def checkerboard_mul(y, nt, c, s, offsets):
    _sweep_grouped(y, nt, c, s, offsets, False, +1)


def checkerboard_transpose_mul(y, nt, c, s, offsets):
    _sweep_grouped(y, nt, c, s, offsets, True, +1)


def checkerboard_inverse_mul(y, nt, c, s, offsets):
    _sweep_grouped(y, nt, c, s, offsets, True, -1)


def checkerboard_inverse_transpose_mul(y, nt, c, s, offsets):
    _sweep_grouped(y, nt, c, s, offsets, False, -1)


def checkerboard_matrix(nt, c, s, nsites, transposed=False):
    """Dense version of ``checkerboard_matrix`` (src/Checkerboard.jl:10-49):
    column ``col`` = sweep applied to the unit vector e_col."""
    K = np.zeros((nsites, nsites), dtype=np.result_type(c.dtype, np.float64))
    for col in range(nsites):
        e = np.zeros(nsites, dtype=K.dtype)
        e[col] = 1.0
        if transposed:
            checkerboard_transpose_mul_literal(e, nt, c, s)
        else:
            checkerboard_mul_literal(e, nt, c, s)
        K[:, col] = e
    return K
