"""Independent replicas per launch (the reference's own scale-out: independent runs, src/ElPhDynamics.jl:90-95): every replica
has its own phonon field, hence its own operator, and its own vector.  These are the entry points behind bench.py's `value`
(Holstein, 24 B per lattice point) and the SSH HBM-regime figure (48 B per point); each replica must reproduce the oracle's
mulMTM of a model holding that replica's field."""
import numpy as np
import pytest

from helpers import engine_holstein_like, oracle_holstein, relerr
from helpers_ssh import engine_ssh_like, oracle_ssh

pytestmark = pytest.mark.gpu


def _eng(v, ncols, L):      # host layout [col][tau] -> engine layout [tau][col]
    return np.ascontiguousarray(v.reshape(ncols, L).T).reshape(-1)


def _host(v, ncols, L):
    return np.ascontiguousarray(v.reshape(L, ncols).T).reshape(-1)


@pytest.mark.parametrize("Lside,beta", [(32, 0.6), (64, 0.3), (4, 1.0)])
def test_holstein_replicas_with_own_tables(Lside, beta):
    import torch
    om, rng = oracle_holstein("square", Lside, beta, 0.1, mu=-0.7)
    em = engine_holstein_like(om)
    nrep, n = 5, om.Ndim
    D = np.zeros((nrep, n)); V = np.zeros((nrep, n)); want = np.zeros((nrep, n))
    for r in range(nrep):
        om.x[:] = rng.normal(size=om.Ndof) * (0.5 + r)
        om.update_model()
        D[r] = _eng(om.expnV, om.N, om.L)
        v = rng.normal(size=n)
        y = np.zeros(n)
        om.mulMTM(y, v)
        V[r], want[r] = _eng(v, om.N, om.L), y
    Dd, Vd = torch.from_numpy(D).cuda(), torch.from_numpy(V).cuda()
    Yd = torch.zeros_like(Vd)
    em._call("elph_dev_mulMTM_replicas", nrep, Dd.data_ptr(), n, Vd.data_ptr(), Yd.data_ptr(), n)
    em.synchronize()
    Y = Yd.cpu().numpy()
    for r in range(nrep):
        assert relerr(_host(Y[r], om.N, om.L), want[r]) <= 1e-12
    em.close()


@pytest.mark.parametrize("Lside,beta,dtau", [(32, 0.4, 0.05), (32, 0.35, 0.05), (64, 0.2, 0.05)])
def test_ssh_replicas_with_own_tables(Lside, beta, dtau):
    import torch
    om, rng = oracle_ssh(Lside=Lside, beta=beta, dtau=dtau, mu=0.1)
    em = engine_ssh_like(om)
    nrep, n, N, L, Nph = 4, om.Ndim, om.N, om.L, om.Nph
    X = np.zeros((nrep, Nph * L)); V = np.zeros((nrep, n)); want = np.zeros((nrep, n))
    x_keep = om.x.copy()
    for r in range(nrep):
        om.x[:] = x_keep * (1.0 + 0.3 * r) + 0.1 * rng.normal(size=om.Ndof)
        om.update_model()
        X[r] = _eng(om.x, Nph, L)
        v = rng.normal(size=n)
        y = np.zeros(n)
        om.mulMTM(y, v)
        V[r], want[r] = _eng(v, N, L), y
    Xd, Vd = torch.from_numpy(X).cuda(), torch.from_numpy(V).cuda()
    Yd = torch.zeros_like(Vd)
    tab_stride = 4 * L * N
    Td = torch.zeros(nrep * tab_stride, dtype=torch.float64, device="cuda")
    em._call("elph_dev_ssh_replica_tables", nrep, Xd.data_ptr(), Nph * L, Td.data_ptr(), tab_stride)
    em._call("elph_dev_mulMTM_replicas_ssh", nrep, Td.data_ptr(), tab_stride, Vd.data_ptr(), Yd.data_ptr(), n)
    em.synchronize()
    Y = Yd.cpu().numpy()
    for r in range(nrep):
        assert relerr(_host(Y[r], N, L), want[r]) <= 1e-12
    # the handle's own field is untouched by the replica calls
    om.x[:] = x_keep
    om.update_model()
    v = rng.normal(size=n)
    yo, ye = np.zeros(n), np.zeros(n)
    om.mulMTM(yo, v)
    import elphdynamics_b200 as E
    E.mulMTM_(ye, em, v)
    assert relerr(ye, yo) <= 1e-12
    em.close()


def test_ssh_replicas_reject_bad_arguments():
    import torch
    om, rng = oracle_ssh(Lside=4, beta=0.4, dtau=0.05)      # 4x4: no register-tile kernel
    em = engine_ssh_like(om)
    t = torch.zeros(4 * om.L * om.N, dtype=torch.float64, device="cuda")
    v = torch.zeros(om.Ndim, dtype=torch.float64, device="cuda")
    with pytest.raises(Exception):
        em._call("elph_dev_mulMTM_replicas_ssh", 1, t.data_ptr(), 4 * om.L * om.N, v.data_ptr(), v.data_ptr(), om.Ndim)
    em.close()
    hm, _ = oracle_holstein("square", 32, 0.4, 0.1)
    eh = engine_holstein_like(hm)
    with pytest.raises(Exception):
        eh._call("elph_dev_mulMTM_replicas_ssh", 1, t.data_ptr(), 4 * hm.L * hm.N, v.data_ptr(), v.data_ptr(), hm.Ndim)
    eh.close()
