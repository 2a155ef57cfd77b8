"""elphdynamics_b200 -- B200 (sm_100a) engine for the ElPhDynamics hot path.

The product is ``libelph_b200.so`` (hand-written CUDA behind the C ABI of
``include/elph_b200.h``).  This package is the thin host-side mirror of the
reference's Julia model/solver API that drives it; see DESIGN.md and
INTEGRATION.md.  There is no CPU implementation of any operator.
"""
from . import _lib
from .lattices import (Lattice, UnitCell, assemble_checkerboard, calc_neighbor_table, checkerboard_groups,
                       checkerboard_order, sorted_neighbor_table_perm)
from .models import (ConjugateGradient, HolsteinModel, I, Identity, SSHModel, SymmetricKPMPreconditioner, kpm_ldiv_, ldiv_, ldiv_batch_, mul_,
                     mulM_, mulMT_, mulMTM_, muldMdx_, setup_, solve_, update_Gr_, update_model_)
from .dynamics import (EulerDynamics, FourierAccelerator, HeunsDynamics, RungeKuttaDynamics, TimeFreqFFT, calc_dSbdx_,
                       calc_dSdx_, calc_Sb, evolve_, fourier_accelerate_, omega_to_tau_, tau_to_omega_, update_M_, update_Q_)
from . import greens, hmc
from .phonon_io import read_phonons_, write_phonons_
from . import workloads

__version__ = "0.1.0"
