import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import test_sharded as T
from elphdynamics_b200.sharded import *
from helpers import relerr
import elphdynamics_b200 as E
for Ls, beta in ((32, 2.0), (8, 2.0)):
    om, noise, Pref, r, z_ref, b, x_ref, it_ref = T._pcg_problem(Ls, beta)
    be = T._cuda_backend(om, 0, om.L)
    aux = T._engine_global(om)
    be.kpm_init(aux, n=20)
    op = ShardedOperator(be, RingComm(0, 1), tol=1e-8, maxiter=5000)
    op.update_model()
    P = ShardedKPM(op, om.N, om.L)
    P.setup(noise)
    print(Ls, "active", P.active, "info", be._kpm_P.info.e_min, be._kpm_P.info.e_max, "ref", Pref.e_min, Pref.e_max)
    eV = np.zeros(om.N)
    rt = be.empty(); rt[1:om.L+1] = torch.from_numpy(r).cuda()
    z = be.empty()
    P.ldiv(z, rt)
    print("  z relerr", relerr(z[1:om.L+1].cpu().numpy(), z_ref))
    # stage checks
    cols = rt[1:om.L+1].contiguous()
    nu = be.tau_to_omega_cols(cols).cpu().numpy()
    theta = np.exp(-1j*np.pi*np.arange(om.L)/om.L)
    nu_ref = np.fft.fft(theta[:, None]*r, axis=0)
    print("  fft relerr", relerr(nu, nu_ref))
    a2 = np.zeros_like(nu_ref)
    for w in range(Pref.Lo2):
        a2[w] = Pref.mul_block(w, nu_ref[w]); a2[om.L-1-w] = np.conj(a2[w])
    got = P.nu_out.cpu().numpy()
    for w in (0, 1, Pref.Lo2-1):
        print("  chain w", w, relerr(got[w], a2[w]), relerr(got[om.L-1-w], a2[om.L-1-w]))
    # the engine's own apply on the aux model with the field set
    aux2 = T._engine_global(om)
    P2 = E.SymmetricKPMPreconditioner(aux2, 20)
    E.setup_(P2, noise)
    zz = np.zeros(om.Ndim)
    E.kpm_ldiv_(zz, P2, r.T.reshape(-1).copy())
    print("  engine apply relerr", relerr(zz.reshape(om.N, om.L).T, z_ref))
