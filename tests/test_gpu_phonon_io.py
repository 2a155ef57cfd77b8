"""Phonon-configuration text files: the package's reader / writer against the reference's
format statement (src/HolsteinModels.jl:764-853, src/SSHModels.jl:838-913), restated line by line below, and the
effect of ``read_phonons!`` on the device tables (``update_model!`` runs at the end of a read)."""
import numpy as np
import pytest

from helpers import engine_holstein_like, oracle_holstein, relerr
from helpers_ssh import engine_ssh_like, oracle_ssh

pytestmark = pytest.mark.gpu


def _expected_holstein(om):
    """The loops of write_phonons!(holstein, filename), literally (1-based orbit / tau, 0-based cell coordinates)."""
    lat, L = om.lat, om.L
    lines = ["L3 L2 L1 orbit tau x\n"]
    for l3 in range(lat.L3):
        for l2 in range(lat.L2):
            for l1 in range(lat.L1):
                for orbit in range(1, lat.norbits + 1):
                    cell = l1 + l2 * lat.L1 + l3 * lat.L1 * lat.L2 + 1           # loc_to_cell, src/Lattices.jl:149-157
                    site = lat.norbits * (cell - 1) + orbit                       # loc_to_site, :164-168
                    for tau in range(1, L + 1):
                        i = (site - 1) * L + tau                                  # get_index
                        lines.append("%d %d %d %d %d %.6f\n" % (l3, l2, l1, orbit, tau, om.x[i - 1]))
    return "".join(lines)


@pytest.mark.parametrize("geom,Ls", [("square", 4), ("honeycomb", 3), ("chain", 5)])
def test_holstein_write_read(tmp_path, geom, Ls):
    import elphdynamics_b200 as E
    om, rng = oracle_holstein(geom, Ls, 0.7, 0.1, mu=-0.3)
    em = engine_holstein_like(om)
    f = tmp_path / "phonons.out"
    E.write_phonons_(em, str(f))
    assert f.read_text() == _expected_holstein(om)
    # a fresh field, then read the file back: x returns to the six-decimal values and the tables follow
    x_file = np.array([float("%.6f" % v) for v in om.x])
    em.x = rng.normal(size=om.Ndof)
    E.read_phonons_(em, str(f))
    assert np.array_equal(em.x, x_file)
    om.x[:] = x_file
    om.update_model()
    v = rng.normal(size=om.Ndim)
    yo, ye = np.zeros(om.Ndim), np.zeros(om.Ndim)
    om.mulM(yo, v)
    E.mulM_(ye, em, v)
    assert relerr(ye, yo) <= 1e-12
    # a partial file only overwrites the entries it names (the reference assigns line by line)
    lines = f.read_text().splitlines(keepends=True)
    (tmp_path / "part.out").write_text("".join(lines[:1 + om.L]))
    x_before = em.x
    em.x = x_before + 1.0
    E.read_phonons_(em, str(tmp_path / "part.out"))
    got = em.x
    first_site = int(lines[1].split(" ")[3]) - 1       # the first block is cell (0,0,0), orbit 1 = site 0
    assert first_site == 0
    assert np.array_equal(got[:om.L], x_before[:om.L]) and np.array_equal(got[om.L:], x_before[om.L:] + 1.0)
    em.close()


@pytest.mark.parametrize("case", [dict(Lside=4, beta=0.5, dtau=0.05), dict(geometry="two_site", beta=0.5, dtau=0.1)])
def test_ssh_write_read(tmp_path, case):
    import elphdynamics_b200 as E
    om, rng = oracle_ssh(**case)
    em = engine_ssh_like(om)
    f = tmp_path / "phonons.out"
    E.write_phonons_(em, str(f))
    n, L = em.nph, om.L
    N = om.Nph // n
    X = om.x.reshape(n, N, L)                          # Julia: reshaped(x, (L, N, n)), column-major
    lines = ["type loc tau x\n"]
    for ph in range(1, n + 1):
        for i in range(1, N + 1):
            for tau in range(1, L + 1):
                lines.append("%d %d %d %.6f\n" % (ph, i, tau, X[ph - 1, i - 1, tau - 1]))
    assert f.read_text() == "".join(lines)
    x_file = np.array([float("%.6f" % v) for v in om.x])
    em.x = np.zeros(om.Ndof)
    E.read_phonons_(em, str(f))
    assert np.array_equal(em.x, x_file)
    om.x[:] = x_file
    om.update_model()
    v = rng.normal(size=om.Ndim)
    yo, ye = np.zeros(om.Ndim), np.zeros(om.Ndim)
    om.mulM(yo, v)
    E.mulM_(ye, em, v)
    assert relerr(ye, yo) <= 1e-12
    em.close()
