"""CPU tests: frozen fixtures vs the oracle and vs the package's host-side geometry code; the C-ABI
library loads and exports every symbol the header declares (no compute calls without a GPU)."""
import ctypes as C
import json
from pathlib import Path

import numpy as np
import pytest

import elphdynamics_b200 as E
from elphdynamics_b200 import _lib
from helpers import GEOMS, oracle_holstein, relerr
from oracle import lattice as olat

GOLD = Path(__file__).resolve().parent / "golden"
TABLES = json.loads((GOLD / "checkerboard_tables.json").read_text())

# SURVEY.md Appendix B / row A0 known answers (1-based, checkerboard order)
SQUARE4 = [(1, 2), (3, 4), (5, 6), (7, 8), (9, 10), (11, 12), (13, 14), (15, 16),
           (1, 4), (2, 3), (5, 8), (6, 7), (9, 12), (10, 11), (13, 16), (14, 15),
           (1, 5), (2, 6), (3, 7), (4, 8), (9, 13), (10, 14), (11, 15), (12, 16),
           (1, 13), (2, 14), (3, 15), (4, 16), (5, 9), (6, 10), (7, 11), (8, 12)]
GROUP_SIZES = {"square2": [2, 2], "square3": [4, 4, 4, 3, 3], "square4": [8] * 4, "square5": [12, 12, 11, 10, 5],
               "square32": [512] * 4, "square64": [2048] * 4, "triangular4": [8] * 6,
               "triangular45": [1012, 1012, 1001, 992, 990, 1001, 55, 12], "triangular46": [1058] * 6,
               "honeycomb3": [9] * 3, "honeycomb32": [1024] * 3}


def _split(key):
    for g in GEOMS:
        if key.startswith(g):
            return g, int(key[len(g):])
    raise KeyError(key)


def _engine_tables(geom, L):
    nd, no, bonds = GEOMS[geom]
    lat = E.Lattice(E.UnitCell(nd, no), L)
    nt = np.concatenate([E.calc_neighbor_table(lat, o1, o2, d) for (o1, o2, d) in bonds], axis=1)
    return E.assemble_checkerboard(nt)


def test_known_answers_from_the_survey():
    assert [tuple(p) for p in TABLES["full"]["square4"]["neighbor_table_1based"]] == SQUARE4
    for key, sizes in GROUP_SIZES.items():
        src = TABLES["full"].get(key) or TABLES["sizes"][key]
        assert src["group_sizes"] == sizes, key


@pytest.mark.parametrize("key", sorted(TABLES["full"]))
def test_oracle_and_host_code_reproduce_integer_fixtures(key):
    geom, L = _split(key)
    nd, no, bonds = GEOMS[geom]
    gold = TABLES["full"][key]
    for literal in (True, False):
        g = olat.BondGeometry(olat.Lattice(nd, no, L), bonds, literal=literal)
        assert (g.neighbor_table + 1).T.tolist() == gold["neighbor_table_1based"]
        assert (g.checkerboard_perm + 1).tolist() == gold["checkerboard_perm_1based"]
        assert (g.inv_checkerboard_perm + 1).tolist() == gold["inv_checkerboard_perm_1based"]
    t = _engine_tables(geom, L)
    assert (t.neighbor_table + 1).T.tolist() == gold["neighbor_table_1based"]
    assert (t.checkerboard_perm + 1).tolist() == gold["checkerboard_perm_1based"]
    assert (t.inv_checkerboard_perm + 1).tolist() == gold["inv_checkerboard_perm_1based"]
    assert t.group_sizes.tolist() == gold["group_sizes"]


@pytest.mark.parametrize("key", sorted(TABLES["sizes"]))
def test_large_tables_bit_exact(key):
    geom, L = _split(key)
    nd, no, bonds = GEOMS[geom]
    gold = TABLES["sizes"][key]
    g = olat.BondGeometry(olat.Lattice(nd, no, L), bonds)
    t = _engine_tables(geom, L)
    assert np.array_equal(t.neighbor_table, g.neighbor_table)
    assert np.array_equal(t.checkerboard_perm, g.checkerboard_perm)
    assert t.group_sizes.tolist() == gold["group_sizes"]
    chk = int(np.sum((t.neighbor_table[0] * 7919 + t.neighbor_table[1]) * (np.arange(t.neighbor_table.shape[1]) % 1009 + 1)))
    assert chk == gold["table_checksum"]
    # colouring invariants: disjoint sites inside a group, every bond exactly once
    off = np.concatenate([[0], np.cumsum(t.group_sizes)])
    for a, b in zip(off[:-1], off[1:]):
        sites = t.neighbor_table[:, a:b].ravel()
        assert len(np.unique(sites)) == sites.size
    assert sorted(t.inv_checkerboard_perm.tolist()) == list(range(t.neighbor_table.shape[1]))


def test_oracle_reproduces_float_fixtures():
    """Regression pin of the oracle itself: re-run the generating recipe, compare with the frozen bytes."""
    from oracle.action import calc_dSbdx, calc_Sb
    from oracle.fourier import TimeFreqFFT
    from oracle.solvers import ConjugateGradient, ldiv
    gold = np.load(GOLD / "holstein_square4.npz")
    om, rng = oracle_holstein("square", 4, 2.0, 0.1, mu=-1.0, seed=20240117, eps=0.3)
    assert np.array_equal(om.x, gold["x"])
    assert np.allclose(om.expnV, gold["expnV"], rtol=1e-15, atol=0)
    v, u = gold["v"], gold["u"]
    y = np.zeros(om.Ndim)
    for name, fn in (("mulM", om.mulM), ("mulMT", om.mulMT), ("mulMTM", om.mulMTM)):
        fn(y, v)
        assert np.allclose(y, gold[name], rtol=1e-14, atol=1e-15), name
    d = np.zeros(om.Ndof)
    om.muldMdx(d, u, v)
    assert np.allclose(d, gold["muldMdx"], rtol=1e-13, atol=1e-15)
    assert np.allclose([calc_Sb(om, False), calc_Sb(om, True)], gold["Sb"], rtol=1e-14)
    assert np.allclose(TimeFreqFFT(om.N, om.L).tau_to_omega(v), gold["tau_to_omega"], rtol=1e-13, atol=1e-13)
    cg = ConjugateGradient(om.Ndim, tol=1e-5, maxiter=10000)
    x = np.zeros(om.Ndim)
    it, res, flag = ldiv(x, om, gold["cg_b"], cg)
    assert (it, flag) == (int(gold["cg_info"][0]), int(gold["cg_info"][2]))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    assert lib.elph_version().startswith(b"elph_b200")
    declared = _lib.declared_symbols()
    assert len(declared) >= 40
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing


def test_no_cpu_fallback_without_a_gpu():
    """Without a CUDA device the product fails loudly at elph_create (no silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lat = E.Lattice(E.UnitCell(2, 1), 4)
    m = E.HolsteinModel(lat, 2.0, 0.1)
    m.assign_t(1.0, 0, 0, (1, 0, 0))
    m.assign_t(1.0, 0, 0, (0, 1, 0))
    with pytest.raises(_lib.ElphError) as ei:
        m.initialize_model_()
    assert "no CUDA device" in str(ei.value) or ei.value.code == 2


def test_package_never_imports_the_oracle():
    import re
    pkg = Path(E.__file__).resolve().parent
    for f in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*")):
        if f.is_file():
            assert not re.search(r"^\s*(from|import)\s+oracle\b", f.read_text(errors="ignore"), flags=re.M), f


def test_phonon_text_format_host_logic(tmp_path, monkeypatch):
    """write_phonons! / read_phonons! (src/HolsteinModels.jl:764-853) on a stand-in model without a device: line order
    (l1 fastest among cells, orbit, tau innermost; 0-based cells, 1-based orbit and tau), '%.6f', partial overwrite."""
    import elphdynamics_b200 as E
    import elphdynamics_b200.phonon_io as pio
    monkeypatch.setattr(pio, "update_model_", lambda m: None)

    class Stub(E.HolsteinModel):
        def __init__(self, lat, L):
            self.lattice, self.Ltau = lat, L
            self._x = np.sin(np.arange(lat.nsites * L, dtype=float))
        x = property(lambda self: self._x.copy(), lambda self, v: setattr(self, "_x", np.asarray(v, float).copy()))

        def __del__(self):
            pass

    lat = E.Lattice(E.UnitCell(2, 2), 3, 2)
    m = Stub(lat, 4)
    f = tmp_path / "p.out"
    pio.write_phonons_(m, str(f))
    lines = f.read_text().splitlines()
    assert lines[0] == "L3 L2 L1 orbit tau x" and len(lines) == 1 + lat.nsites * 4
    assert lines[1] == "0 0 0 1 1 0.000000" and lines[5] == "0 0 0 2 1 %.6f" % np.sin(4.0)
    assert lines[1 + 2 * 4] == "0 0 1 1 1 %.6f" % np.sin(8.0)              # next cell along l1 = site 3 (1-based)
    assert lines[1 + 6 * 4] == "0 1 0 1 1 %.6f" % np.sin(24.0)             # l2 advances after L1 = 3 cells
    x0 = m.x
    m.x = np.zeros_like(x0)
    pio.read_phonons_(m, str(f))
    assert np.abs(m.x - x0).max() <= 5e-7
    assert np.array_equal(m.x, np.array([float("%.6f" % v) for v in x0]))


def _extras_problem():
    from helpers import oracle_holstein
    gold = np.load(GOLD / "honeycomb2_extras.npz")
    om, _ = oracle_holstein("honeycomb", 2, 0.4, 0.1, mu=-0.3, seed=20240229, eps=0.3, tol=1e-7)
    assert np.array_equal(om.x, gold["x"])
    return om, gold


def test_oracle_reproduces_greens_and_special_update_fixture():
    """tests/golden/honeycomb2_extras.npz (made by make_golden.py): Green's-function convolutions for frozen vectors and
    the log of three reflection + three swap proposals."""
    from oracle import greens as og
    from oracle import hmc as ohmc
    from oracle.solvers import ConjugateGradient
    om, gold = _extras_problem()
    Gr = og.EstimateGreensFunction(om, 2)
    Gr.R[:], Gr.MinvR[:] = gold["greens_R"], gold["greens_MinvR"]
    for name, arr in zip(("G_D0", "G_D0_G_D0", "G_DD_G_00", "G_D0_G_0D"), og.setup(Gr, 0, 1)):
        assert np.allclose(arr, gold["greens_" + name], rtol=1e-13, atol=1e-15), name
    cg = ConjugateGradient(om.Ndim, tol=om.tol, maxiter=om.maxiter)
    h = ohmc.HybridMonteCarlo(om, 0.01, 0.05, 0.0, 1)
    for k, kind in enumerate(("reflect", "swap")):
        tg = gold[f"special_{kind}_targets"]
        targets = [int(t) for t in tg] if kind == "reflect" else [tuple(int(v) for v in t) for t in tg]
        ratio, log = ohmc.special_update(om, h, cg, None, kind, targets, list(gold[f"special_{kind}_Rp"]),
                                         list(gold[f"special_{kind}_Rm"]), gold[f"special_{kind}_u"].tolist())
        ref = gold[f"special_{kind}_log"]
        assert ratio == gold["special_ratios"][k]
        for (ok, s0, s1, it, fl), row in zip(log, ref):
            assert float(ok) == row[0] and it == row[3] and fl == row[4]
            assert abs(s0 - row[1]) <= 1e-12 * abs(row[1]) and abs(s1 - row[2]) <= 1e-10 * abs(row[2])
    assert np.array_equal(om.x, gold["x_after_special"])


@pytest.mark.gpu
def test_engine_reproduces_greens_and_special_update_fixture():
    """The same fixture through the C ABI: convolutions to 1e-12, proposals with the same decisions and final field."""
    from helpers import engine_holstein_like
    import elphdynamics_b200 as E
    from elphdynamics_b200 import greens as eg
    from elphdynamics_b200 import hmc as ehmc
    om, gold = _extras_problem()
    em = engine_holstein_like(om)
    Ge = eg.EstimateGreensFunction(em, 2)
    Ge.R[:], Ge.MinvR[:] = gold["greens_R"], gold["greens_MinvR"]
    em._call("elph_greens_load", 2, E._lib.ptr(Ge.R), E._lib.ptr(Ge.MinvR))
    for name, arr in zip(("G_D0", "G_D0_G_D0", "G_DD_G_00", "G_D0_G_0D"), eg.setup_pair_(Ge, 0, 1)):
        ref = gold["greens_" + name]
        assert np.linalg.norm(arr - ref) <= 1e-12 * np.linalg.norm(ref), name
    he = ehmc.HybridMonteCarlo(em, 0.01, 0.05, 0.0, 1)
    for k, (kind, upd) in enumerate((("reflect", ehmc.ReflectionUpdate(em, 1, 3)), ("swap", ehmc.SwapUpdate(em, 1, 3)))):
        tg = gold[f"special_{kind}_targets"]
        targets = [int(t) for t in tg] if kind == "reflect" else [tuple(int(v) for v in t) for t in tg]
        ratio = ehmc.special_update_(em, he, upd, None, targets=targets, R_plus=list(gold[f"special_{kind}_Rp"]),
                                     R_minus=list(gold[f"special_{kind}_Rm"]), uniforms=gold[f"special_{kind}_u"].tolist())
        assert ratio == gold["special_ratios"][k]
        for (ok, s0, s1, it, fl), row in zip(he.special_log, gold[f"special_{kind}_log"]):
            assert float(ok) == row[0] and abs(it - row[3]) <= 2 and fl == row[4]
            assert abs(s0 - row[1]) <= 1e-12 * abs(row[1]) and abs(s1 - row[2]) <= 1e-8 * abs(row[2])
    assert np.array_equal(em.x, gold["x_after_special"])
    Ge.close()
    em.close()


def test_oracle_chain_regression():
    """Row N1 on the CPU: the oracle's fixed-seed chain of the shipped example (updates + <x>, <x^2>, G(0,0) measurements,
    tests/helpers_chain.py) reproduces the frozen series of tests/golden/chain_square4.npz."""
    from helpers_chain import Noise, OracleChain, run_chain
    gold = np.load(GOLD / "chain_square4.npz")
    om, _ = oracle_holstein("square", 4, 2.0, 0.1, mu=-1.0, seed=1234, eps=0.3, tol=1e-5)
    series = run_chain(OracleChain(om, 0.02), Noise(202, om.Ndof, om.Ndim, om.N), burnin=10, nsteps=24, meas_freq=4)
    assert series.shape == gold["series"].shape
    assert np.allclose(series, gold["series"], rtol=1e-7, atol=1e-9)
    assert relerr(om.x, gold["x_final"]) <= 1e-7
