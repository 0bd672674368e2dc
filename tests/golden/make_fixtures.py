"""Generates tests/golden/si_block_fixture.npz: seeded inputs and the ORACLE's outputs for one silicon k-block
(Hψ, LOBPCG eigenvalues, density).  The reference itself (Julia) cannot run here, so these vectors come from
the oracle after it has been pinned to the reference's known-answer tests (tests/test_oracle_golden.py); they
freeze the oracle against regressions and let the GPU parity tests run from committed data.
Run from the repo root:  python tests/golden/make_fixtures.py"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.basis import Element, Model, PlaneWaveBasis
from oracle.terms import Terms, energy_hamiltonian, guess_density
from oracle import lobpcg as olob, scf as oscf
from silicon import LATTICE, POSITIONS

m = Model(LATTICE, [Element("Si")] * 2, POSITIONS, functionals=("lda_x", "lda_c_vwn"), symmetries=False)
b = PlaneWaveBasis(m, 10, fft_size=(24, 24, 24), kcoords=[[0.1, -0.2, 0.3]], kweights=[1.0])
_, ham = energy_hamiltonian(b, Terms(b), None, None, guess_density(b))
blk = ham[0]
rng = np.random.default_rng(2026)
psi = rng.standard_normal((6, blk.kpt.n_G)) + 1j * rng.standard_normal((6, blk.kpt.n_G))
hpsi = blk.matmul(psi.T).T
res = olob.lobpcg(blk, psi.T.copy(), olob.PreconditionerTPA(blk.kin), tol=1e-10, maxiter=200)
occ = np.array([2.0, 2.0, 2.0, 2.0, 0.0, 0.0])
rho = oscf.compute_density(b, [res["X"]], [occ])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "si_block_fixture.npz"),
                    fft_size=np.array(b.fft_size), mapping=blk.kpt.mapping, kin=blk.kin, V=blk.Vtot,
                    P=blk.PD[0], D=blk.PD[1], psi=psi, hpsi=hpsi, eigenvalues=res["λ"], rho=rho[0], occ=occ,
                    volume=m.unit_cell_volume)
print("written", hpsi.shape, res["λ"])
