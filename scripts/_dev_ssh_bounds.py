import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import test_sharded as T
from elphdynamics_b200.sharded import *
for Ls, beta in ((32, 1.0), (32, 1.05)):
    om, noise, Pref, r, z_ref, b, x_ref, it_ref = T._pcg_problem(Ls, beta, kind="ssh")
    be = T._cuda_ssh_backend(om, 0, om.L)
    be.kpm_init(T._engine_ssh_global(om), n=20)
    op = ShardedOperator(be, RingComm(0, 1), tol=1e-8, maxiter=5000)
    op.update_model()
    P = ShardedKPM(op, om.N, om.L)
    P.setup(noise)
    lo, hi, e_min, e_max = be.kpm_window()
    print("ssh", Ls, beta, "rel dev e_min", abs(e_min - Pref.e_min) / Pref.e_min, "e_max", abs(e_max - Pref.e_max) / Pref.e_max)
    # the single-GPU engine's own set-up on the same field for comparison
    import elphdynamics_b200 as E
    em = T._engine_ssh_global(om); E.update_model_(em)
    Pe = E.SymmetricKPMPreconditioner(em, 20); info = E.setup_(Pe, noise)
    print("   single-GPU engine: rel dev e_min", abs(info.e_min - Pref.e_min) / Pref.e_min, "e_max", abs(info.e_max - Pref.e_max) / Pref.e_max)
