"""Generate the frozen fixtures under tests/golden/.

The reference is pure Julia and cannot run in this image (no julia binary, no network), and it ships
no golden vectors of its own.  These fixtures are therefore produced by the ORACLE
(oracle/, a line-by-line NumPy restatement) and frozen, so that later changes to the oracle, to the
package's host-side geometry code or to the CUDA engine are all checked against the same bytes.

  python tests/golden/make_golden.py        # rewrites tests/golden/*.json, *.npz

Integer fixtures (neighbour tables in checkerboard order, 1-based like the reference, permutations,
group sizes) are exact.  Float fixtures are the oracle's float64 outputs for seeded inputs on the
shipped example shape A (examples/holstein_langevin_square.toml: 4x4, beta=2, dtau=0.1).
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from helpers import oracle_holstein  # noqa: E402
from oracle import lattice as olat  # noqa: E402
from oracle.action import calc_dSbdx, calc_Sb  # noqa: E402
from oracle.fourier import FourierAccelerator, TimeFreqFFT  # noqa: E402
from oracle.kpm import KPMPreconditioner  # noqa: E402
from oracle.solvers import ConjugateGradient, ldiv  # noqa: E402
from oracle import langevin as olang  # noqa: E402

OUT = Path(__file__).resolve().parent
GEOMS = {"square": (2, 1, olat.SQUARE_BONDS), "honeycomb": (2, 2, olat.HONEYCOMB_BONDS),
         "triangular": (2, 1, olat.TRIANGULAR_BONDS)}


def integer_fixtures():
    full = {}
    for geom, sizes in (("square", (2, 3, 4, 5)), ("honeycomb", (3,)), ("triangular", (3, 4))):
        nd, no, bonds = GEOMS[geom]
        for L in sizes:
            g = olat.BondGeometry(olat.Lattice(nd, no, L), bonds, literal=True)
            full[f"{geom}{L}"] = {
                "neighbor_table_1based": (g.neighbor_table + 1).T.tolist(),
                "checkerboard_perm_1based": (g.checkerboard_perm + 1).tolist(),
                "inv_checkerboard_perm_1based": (g.inv_checkerboard_perm + 1).tolist(),
                "group_sizes": np.diff(g.group_offsets).tolist(),
            }
    sizes_only = {}
    for geom, L in (("square", 32), ("square", 64), ("triangular", 45), ("triangular", 46), ("honeycomb", 32)):
        nd, no, bonds = GEOMS[geom]
        g = olat.BondGeometry(olat.Lattice(nd, no, L), bonds)
        sizes_only[f"{geom}{L}"] = {"group_sizes": np.diff(g.group_offsets).tolist(), "nbonds": int(g.nbonds),
                                    "table_checksum": int(np.sum((g.neighbor_table[0] * 7919 + g.neighbor_table[1]) * (np.arange(g.nbonds) % 1009 + 1)))}
    (OUT / "checkerboard_tables.json").write_text(json.dumps({"full": full, "sizes": sizes_only}, indent=0))


def float_fixtures():
    om, rng = oracle_holstein("square", 4, 2.0, 0.1, mu=-1.0, seed=20240117, eps=0.3)
    d = {"x": om.x.copy(), "expnV": om.expnV.copy()}
    v = rng.normal(size=om.Ndim)
    u = rng.normal(size=om.Ndim)
    d["v"], d["u"] = v, u
    for name, fn in (("mulM", om.mulM), ("mulMT", om.mulMT), ("mulMTM", om.mulMTM)):
        y = np.zeros(om.Ndim)
        fn(y, v)
        d[name] = y
    dm = np.zeros(om.Ndof)
    om.muldMdx(dm, u, v)
    d["muldMdx"] = dm
    d["Sb"] = np.array([calc_Sb(om, False), calc_Sb(om, True)])
    ds = np.zeros(om.Ndof)
    calc_dSbdx(ds, om, True)
    d["dSbdx_shifted"] = ds
    fft = TimeFreqFFT(om.N, om.L)
    d["tau_to_omega"] = fft.tau_to_omega(v)
    fa = FourierAccelerator(om.Nph, om.L, om.dtau, om.omega)
    fa.update_Q(0.0, 10.0, 1.0)
    d["Q"] = fa.Q.copy()
    d["fa_half"] = fa.accelerate(v, 0.5)
    cg = ConjugateGradient(om.Ndim, tol=1e-5, maxiter=10000)
    b = np.zeros(om.Ndim)
    om.mulMT(b, u)
    x = np.zeros(om.Ndim)
    it, res, flag = ldiv(x, om, b, cg)
    d["cg_b"], d["cg_x"], d["cg_info"] = b, x, np.array([it, res, flag])
    P = KPMPreconditioner(om)
    noise = rng.normal(size=2 * om.N)
    P.setup(noise)
    d["arnoldi_noise"] = noise
    d["kpm_bounds"] = np.array([P.e_min, P.e_max, P.lam_lo, P.lam_hi])
    d["kpm_orders"] = P.order.copy()
    z = np.zeros(om.Ndim)
    P.ldiv(z, v)
    d["kpm_apply"] = z
    x = np.zeros(om.Ndim)
    it, res, flag = ldiv(x, om, b, cg, P)
    d["pcg_info"] = np.array([it, res, flag])
    # one Runge-Kutta Langevin step with injected noise (dt = 1e-3, the shipped example's update_method = 2)
    eta, g1, g2 = rng.normal(size=om.Ndof), rng.normal(size=om.Ndim), rng.normal(size=om.Ndim)
    a1, a2 = rng.normal(size=2 * om.N), rng.normal(size=2 * om.N)
    P = KPMPreconditioner(om)
    it = olang.evolve_rk(om, cg, fa, P, 1e-3, eta, g1, g2, a1, a2)
    d.update(rk_eta=eta, rk_g1=g1, rk_g2=g2, rk_a1=a1, rk_a2=a2, rk_x_after=om.x.copy(), rk_iters=np.array([it]))
    np.savez_compressed(OUT / "holstein_square4.npz", **d)


def extras_fixtures():
    """Measurement-side and special-update fixtures: Green's-function convolutions for fixed vectors, the log of
    three reflection and three swap proposals with injected noise, on a small honeycomb lattice (two orbitals)."""
    from oracle import greens as og
    from oracle import hmc as ohmc
    om, rng = oracle_holstein("honeycomb", 2, 0.4, 0.1, mu=-0.3, seed=20240229, eps=0.3, tol=1e-7)
    d = {"x": om.x.copy()}
    Gr = og.EstimateGreensFunction(om, 2)
    Gr.R[:] = rng.normal(size=Gr.R.shape)
    Gr.MinvR[:] = rng.normal(size=Gr.R.shape)
    d["greens_R"], d["greens_MinvR"] = Gr.R.copy(), Gr.MinvR.copy()
    for name, arr in zip(("G_D0", "G_D0_G_D0", "G_DD_G_00", "G_D0_G_0D"), og.setup(Gr, 0, 1)):
        d["greens_" + name] = np.asarray(arr)
    cg = ConjugateGradient(om.Ndim, tol=om.tol, maxiter=om.maxiter)
    h = ohmc.HybridMonteCarlo(om, 0.01, 0.05, 0.0, 1)
    logs = []
    for kind, targets in (("reflect", [1, 5, 2]), ("swap", [(0, 1), (2, 7), (4, 5)])):
        Rp = [rng.normal(size=om.Ndim) for _ in targets]
        Rm = [rng.normal(size=om.Ndim) for _ in targets]
        u = rng.uniform(size=len(targets))
        ratio, log = ohmc.special_update(om, h, cg, None, kind, targets, Rp, Rm, u.tolist())
        d[f"special_{kind}_Rp"], d[f"special_{kind}_Rm"], d[f"special_{kind}_u"] = np.array(Rp), np.array(Rm), u
        d[f"special_{kind}_targets"] = np.array(targets)
        d[f"special_{kind}_log"] = np.array([[float(a), s0, s1, float(it), float(fl)] for (a, s0, s1, it, fl) in log])
        logs.append(ratio)
    d["special_ratios"] = np.array(logs)
    d["x_after_special"] = om.x.copy()
    np.savez_compressed(OUT / "honeycomb2_extras.npz", **d)


def chain_fixture():
    """Row N1: the measurement series of a short fixed-seed chain of the shipped example (tests/helpers_chain.py)."""
    from helpers_chain import Noise, OracleChain, run_chain
    om, _ = oracle_holstein("square", 4, 2.0, 0.1, mu=-1.0, seed=1234, eps=0.3, tol=1e-5)
    series = run_chain(OracleChain(om, 0.02), Noise(202, om.Ndof, om.Ndim, om.N), burnin=10, nsteps=24, meas_freq=4)
    np.savez(OUT / "chain_square4.npz", series=series, x_final=om.x.copy())


if __name__ == "__main__":
    integer_fixtures()
    float_fixtures()
    extras_fixtures()
    chain_fixture()
    print("wrote", sorted(p.name for p in OUT.iterdir()))
