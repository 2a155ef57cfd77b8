"""setup!(estimator, n1, n2) at config B (32x32, Ltau = 200): device convolutions (csrc/greens.cu) per pair, against the
NumPy restatement (numpy.fft = pocketfft, one host thread) on the same vectors.  Development aid; the numbers are quoted
in DESIGN.md.
    python scripts/time_greens.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import elphdynamics_b200 as E
from elphdynamics_b200 import greens as eg
from elphdynamics_b200 import workloads

m, rng = workloads.config("B")
nv = 4
Gr = eg.EstimateGreensFunction(m, nv)
Gr.R[:] = rng.normal(size=Gr.R.shape)
Gr.MinvR[:] = rng.normal(size=Gr.R.shape)
m._call("elph_greens_load", nv, E._lib.ptr(Gr.R), E._lib.ptr(Gr.MinvR))
eg.setup_pair_(Gr, 0, 1)
t0 = time.perf_counter()
npairs = 0
for i in range(nv - 1):
    for j in range(i + 1, nv):
        eg.setup_pair_(Gr, i, j)
        npairs += 1
dt = (time.perf_counter() - t0) / npairs
print(f"device: {dt * 1e3:.3f} ms per pair (4 convolutions, 12 transforms, 4 x 6.5 MB back to the host)")
try:
    from oracle import greens as og
    from helpers import oracle_holstein
    om, _ = oracle_holstein("square", 32, 20.0, 0.1, mu=-1.0)
    Go = og.EstimateGreensFunction(om, nv)
    Go.R[:], Go.MinvR[:] = Gr.R, Gr.MinvR
    t0 = time.perf_counter()
    ref = og.setup(Go, nv - 2, nv - 1)
    dtc = time.perf_counter() - t0
    err = max(float(np.linalg.norm(a - b) / np.linalg.norm(b)) for a, b in zip((Gr.G_D0, Gr.G_D0_G_D0, Gr.G_DD_G_00, Gr.G_D0_G_0D), ref))
    print(f"numpy (1 thread): {dtc * 1e3:.1f} ms per pair; max relative difference {err:.2e}")
except ImportError:
    pass
m.close()
