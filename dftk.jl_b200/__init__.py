"""dftk_b200: B200-native plane-wave Kohn-Sham SCF hot path behind DFTK.jl's operator API.

The package directory is `dftk.jl_b200/`; import it as `dftk_b200` (see dftk_b200.py at the repo root).
Names follow the reference (Model, PlaneWaveBasis, self_consistent_field, HamiltonianBlock, ...).
"""
from . import _lib
from ._lib import DftkB200Error, LIB_PATH
from .device import Context, FFTGrid, KBlock
from .architecture import B200, CPU
from .pseudo import PspHgh, ElementPsp, load_psp, parse_hgh
from .model import Model, model_DFT, model_atomic, LDA, PBE, SymOp, symmetry_operations
from .parallel import KpointComm, split_evenly
from .basis import PlaneWaveBasis, MonkhorstPack, ExplicitKpoints, Kpoint, compute_fft_size
from .terms import guess_density
from .hamiltonian import Hamiltonian, DftHamiltonianBlock, energy_hamiltonian, energy, Energies
from .eigen import lobpcg_hyper, diagonalize_all_kblocks, random_orbitals
from .occupation import compute_occupation
from .densities import compute_density, symmetrize_rho
from .scf import (self_consistent_field, next_density, AdaptiveBands, FixedBands, AdaptiveDiagtol,
                  ScfConvergenceDensity, ScfConvergenceEnergy, SimpleMixing, KerkerMixing,
                  AndersonAcceleration, ScfDefaultCallback)
