"""Pins the oracle's Hellmann-Feynman forces the way the reference pins its own (test/forces.jl:17-57,
`test_term_forces`): term by term, the analytic force must equal the central finite difference of that term's
energy under a displacement of one atom with ψ, occupation and ρ held fixed."""
import numpy as np
import pytest

from oracle.psp_hgh import PspHgh
from oracle.basis import Element, Model, PlaneWaveBasis
from oracle.terms import Terms, energy_hamiltonian
from oracle import scf, forces
from silicon import LATTICE, POSITIONS


def _si(positions, symmetries):
    si = Element("Si", PspHgh.from_table("Si", "lda"))
    return Model(LATTICE, [si, si], positions, functionals=("lda_x", "lda_c_vwn"), symmetries=symmetries)


@pytest.fixture(scope="module")
def rattled():
    pos = [POSITIONS[0] + np.array([0.011, -0.007, 0.004]), POSITIONS[1] + np.array([-0.003, 0.009, 0.006])]
    model = _si(pos, False)
    basis = PlaneWaveBasis(model, Ecut=7, kgrid=(2, 2, 2))
    res = scf.self_consistent_field(basis, tol=1e-10, maxiter=60)
    assert res["converged"]
    return model, basis, res


def test_term_forces_match_finite_differences(rattled):
    model, basis, res = rattled
    total, parts = forces.compute_forces(basis, res["psi"], res["occupation"], res["rho"])
    rng = np.random.default_rng(3)
    iatom = 1
    direction = rng.standard_normal(3)
    direction /= np.linalg.norm(direction)
    eps = 1e-5

    def term_energies(e):
        pos = [p.copy() for p in model.positions]
        pos[iatom] = pos[iatom] + e * direction
        mb = PlaneWaveBasis(_si(pos, False), Ecut=7, fft_size=basis.fft_size, kcoords=basis.kcoords_global,
                            kweights=basis.kweights_global)
        E, _ = energy_hamiltonian(mb, Terms(mb), res["psi"], res["occupation"], res["rho"], res["eigenvalues"],
                                  res["eF"], only_energy=True)
        return E
    Ep, Em = term_energies(eps), term_energies(-eps)
    for term in ["Kinetic", "AtomicLocal", "AtomicNonlocal", "Ewald", "PspCorrection", "Hartree", "Xc"]:
        fd = -(Ep[term] - Em[term]) / (2 * eps)
        hf = float(direction @ parts[term][iatom]) if term in parts else 0.0
        assert abs(hf - fd) < 1e-7, (term, hf, fd)
    # the forces on this rattled cell are far from zero, so the comparison is not vacuous
    assert np.linalg.norm(total[iatom]) > 1e-3
    # translation invariance: the total force on the cell vanishes (up to the egg-box effect of the xc grid)
    assert np.linalg.norm(total[0] + total[1]) < 5e-4


def test_total_force_is_energy_derivative(rattled):
    """test/forces.jl:59-88 (`test_forces`): -dE_total/dx from two re-converged SCFs."""
    model, basis, res = rattled
    total, _ = forces.compute_forces(basis, res["psi"], res["occupation"], res["rho"])
    direction = np.array([0.6, -0.64, 0.48])
    eps = 1e-4

    def etot(e):
        pos = [p.copy() for p in model.positions]
        pos[0] = pos[0] + e * direction
        mb = PlaneWaveBasis(_si(pos, False), Ecut=7, fft_size=basis.fft_size, kcoords=basis.kcoords_global,
                            kweights=basis.kweights_global)
        return scf.self_consistent_field(mb, rho=res["rho"], tol=1e-10, maxiter=60)["energies"]["total"]
    fd = -(etot(eps) - etot(-eps)) / (2 * eps)
    assert abs(float(direction @ total[0]) - fd) < 2e-6


def test_symmetrized_forces_match_unfolded_bz():
    """Forces from the irreducible BZ (symmetrised, symmetry.jl:399-413) equal those of the full k-grid."""
    shift = 0.003 * np.ones(3)                      # displacement along [111] keeps a C3v subgroup
    pos = [POSITIONS[0] + shift, POSITIONS[1]]
    out = []
    for sym in (True, False):
        model = _si(pos, sym)
        basis = PlaneWaveBasis(model, Ecut=7, kgrid=(2, 2, 2), fft_size=(20, 20, 20))
        res = scf.self_consistent_field(basis, tol=1e-9, maxiter=60)
        total, _ = forces.compute_forces(basis, res["psi"], res["occupation"], res["rho"])
        out.append((len(basis.symmetries), len(basis.kpoints), np.array(total)))
    assert out[0][0] > 1 and out[0][1] < out[1][1]
    # the inversion centre (bond midpoint) is not a point of the 20^3 grid, so the unsymmetrised run carries the xc
    # egg-box asymmetry (~2e-6) that the symmetrised one projects out
    np.testing.assert_allclose(out[0][2], out[1][2], atol=5e-6)
    np.testing.assert_allclose(out[0][2], (out[1][2] - out[1][2][::-1]) / 2, atol=2e-7)
    # reduced forces along [111] have three equal components
    np.testing.assert_allclose(out[0][2][0], out[0][2][0][0] * np.ones(3), atol=1e-8)
    cart = forces.forces_cart(_si(pos, True), list(out[0][2]))
    assert np.linalg.norm(cart[0]) > 1e-3


def test_ewald_forces_match_energy_derivative():
    pos = [POSITIONS[0] + np.array([0.01, 0.02, -0.015]), POSITIONS[1]]
    e0, f = forces.energy_forces_ewald(LATTICE, [4, 4], pos)
    from oracle.terms import energy_ewald
    assert e0 == pytest.approx(energy_ewald(LATTICE, [4, 4], pos), abs=1e-12)
    d = np.array([0.3, -0.5, 0.81])
    eps = 1e-5
    ep = energy_ewald(LATTICE, [4, 4], [pos[0] + eps * d, pos[1]])
    em = energy_ewald(LATTICE, [4, 4], [pos[0] - eps * d, pos[1]])
    assert float(d @ f[0]) == pytest.approx(-(ep - em) / (2 * eps), abs=1e-8)
