timeout 600 python -m pytest tests/test_gpu_speculative_setup.py tests/test_gpu_holstein.py tests/test_gpu_ssh.py tests/test_gpu_observables.py tests/test_gpu_pcg_fused.py -m gpu -x -q 2>&1 | tail -15
python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
import elphdynamics_b200 as E
from elphdynamics_b200 import workloads
for spec in (1, 0):
    m, rng = workloads.config("B")
    m._call("elph_set_tuning", 25, spec)
    fa = E.FourierAccelerator(m); E.update_Q_(fa, m, 0.0, 10.0, 1.0)
    P = E.SymmetricKPMPreconditioner(m); dyn = E.RungeKuttaDynamics(m, 1e-3)
    its = []
    ts = []
    for step in range(30):
        eta, g1, g2 = rng.normal(size=m.Ndof), rng.normal(size=m.Ndim), rng.normal(size=m.Ndim)
        a1, a2 = rng.normal(size=2 * m.Nsites), rng.normal(size=2 * m.Nsites)
        t0 = time.perf_counter()
        its.append(E.evolve_(m, dyn, fa, P, eta=eta, g1=g1, g2=g2, arnoldi1=a1, arnoldi2=a2))
        ts.append(time.perf_counter() - t0)
    print("speculate", spec, "ms/step (median of last 20)", np.median(ts[10:]) * 1e3, "steps/s", 1 / np.median(ts[10:]), "iters", its[-5:])
    m.close()
PY
