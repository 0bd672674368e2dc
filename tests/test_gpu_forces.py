"""GPU parity of the Hellmann-Feynman forces (dftk_b200_local_forces / dftk_b200_nonlocal_force_rows + host Ewald)
against the oracle, whose forces are pinned by finite differences in tests/test_oracle_forces.py (the reference's own
strategy, test/forces.jl)."""
import numpy as np
import pytest
import torch

from silicon import LATTICE, POSITIONS

pytestmark = pytest.mark.gpu

RATTLED = [POSITIONS[0] + np.array([0.011, -0.007, 0.004]), POSITIONS[1] + np.array([-0.003, 0.009, 0.006])]


def _oracle(positions, symmetries, Ecut, kgrid, fft_size=None, tol=1e-10):
    from oracle.psp_hgh import PspHgh
    from oracle.basis import Element, Model, PlaneWaveBasis
    from oracle import scf, forces
    si = Element("Si", PspHgh.from_table("Si", "lda"))
    m = Model(LATTICE, [si, si], positions, functionals=("lda_x", "lda_c_vwn"), symmetries=symmetries)
    b = PlaneWaveBasis(m, Ecut, kgrid=kgrid, fft_size=fft_size)
    res = scf.self_consistent_field(b, tol=tol, maxiter=60)
    assert res["converged"]
    total, parts = forces.compute_forces(b, res["psi"], res["occupation"], res["rho"])
    return b, res, total, parts


def _product_basis(positions, symmetries, Ecut, kgrid, fft_size=None):
    import dftk_b200 as dftk
    Si = dftk.ElementPsp("Si", functional="lda")
    model = dftk.model_DFT(LATTICE, [Si, Si], positions, functionals=["lda_x", "lda_c_vwn"], symmetries=symmetries)
    return dftk, dftk.PlaneWaveBasis(model, Ecut=Ecut, kgrid=kgrid, fft_size=fft_size)


def test_force_kernels_match_oracle_on_the_same_state():
    """Same ψ, occupation and ρ on both sides: every force term agrees to round-off."""
    ob, ores, ototal, oparts = _oracle(RATTLED, False, 7, (2, 2, 2))
    dftk, basis = _product_basis(RATTLED, False, 7, (2, 2, 2), fft_size=ob.fft_size)
    dev = basis.architecture.device
    psi, occ = [], []
    for kpt in basis.kpoints:
        jk = [j for j, ok in enumerate(ob.kpoints) if np.allclose(ok.coordinate, kpt.coordinate)][0]
        assert np.array_equal(kpt.mapping.cpu().numpy(), ob.kpoints[jk].mapping)
        psi.append(torch.from_numpy(np.ascontiguousarray(ores["psi"][jk].T)).to(dev))
        occ.append(ores["occupation"][jk])
    rho = torch.from_numpy(ores["rho"]).to(dev)
    total, parts = dftk.compute_forces(basis, psi, occ, rho=rho, per_term=True)
    assert set(parts) == {"AtomicLocal", "AtomicNonlocal", "Ewald"}
    for name in parts:
        np.testing.assert_allclose(np.array(parts[name]), np.array(oparts[name]), atol=1e-10, err_msg=name)
    np.testing.assert_allclose(np.array(total), np.array(ototal), atol=1e-10)
    assert np.linalg.norm(np.array(total)) > 1e-2
    cart = dftk.compute_forces_cart(basis, psi, occ, rho=rho)
    np.testing.assert_allclose(np.array(cart), np.array([np.linalg.inv(LATTICE).T @ f for f in ototal]), atol=1e-10)


def test_scf_forces_match_oracle():
    """End to end: forces of the product's own SCF solution against the oracle's (both converged to 1e-10)."""
    ob, ores, ototal, _ = _oracle(RATTLED, False, 7, (2, 2, 2))
    dftk, basis = _product_basis(RATTLED, False, 7, (2, 2, 2), fft_size=ob.fft_size)
    res = dftk.self_consistent_field(basis, tol=1e-10)
    assert res["converged"]
    assert abs(res["energies"].total - ores["energies"]["total"]) < 2e-8
    np.testing.assert_allclose(np.array(dftk.compute_forces(res)), np.array(ototal), atol=1e-7)


def test_symmetrised_forces_match_oracle():
    """Irreducible BZ + symmetrize_forces (symmetry.jl:399-413): one atom moved along [111], 12 operations."""
    pos = [POSITIONS[0] + 0.003 * np.ones(3), POSITIONS[1]]
    ob, ores, ototal, oparts = _oracle(pos, True, 7, (2, 2, 2), fft_size=(20, 20, 20), tol=1e-9)
    dftk, basis = _product_basis(pos, True, 7, (2, 2, 2), fft_size=(20, 20, 20))
    assert len(basis.symmetries) == len(ob.symmetries) == 12
    assert len(basis.kpoints) == len(ob.kpoints) < 8
    res = dftk.self_consistent_field(basis, tol=1e-9)
    total, parts = dftk.compute_forces(res["basis"], res["psi"], res["occupation"], rho=res["rho"], per_term=True)
    for name in parts:
        np.testing.assert_allclose(np.array(parts[name]), np.array(oparts[name]), atol=2e-7, err_msg=name)
    np.testing.assert_allclose(np.array(total), np.array(ototal), atol=2e-7)
    np.testing.assert_allclose(total[0], total[0][0] * np.ones(3), atol=1e-9)


def test_force_entry_points_reject_bad_arguments():
    import ctypes
    import dftk_b200 as dftk
    from dftk_b200.device import _ptr
    _, basis = _product_basis(RATTLED, False, 5, (1, 1, 1))
    ctx = basis.architecture.ctx
    out = np.zeros(3)
    pos = np.zeros(3)
    assert ctx.L.dftk_b200_local_forces(basis.fft_grid.h, None, 1, _ptr(pos), _ptr(out)) == -1
    host_w = np.zeros(basis.N, dtype=complex)
    assert ctx.L.dftk_b200_local_forces(basis.fft_grid.h, _ptr(host_w), 1, _ptr(pos), _ptr(out)) == -1
    assert b"device memory" in ctx.L.dftk_b200_last_error(ctx.h)
    kb = basis.kblocks[0]
    rows = np.zeros((3, kb.n_proj))
    assert ctx.L.dftk_b200_nonlocal_force_rows(kb.h, None, _ptr(out), 1, None, _ptr(rows)) == -1
    # zero bands: a no-op that returns zeros
    psi = torch.zeros((1, kb.n_pw), dtype=torch.complex128, device=ctx.device)
    gpk = basis.Gplusk_vectors(basis.kpoints[0]).T.contiguous()
    rows[:] = 1.0
    assert ctx.L.dftk_b200_nonlocal_force_rows(kb.h, _ptr(psi), _ptr(out), 0, _ptr(gpk), _ptr(rows)) == 0
    assert not rows.any()


def test_ewald_kernels_match_oracle():
    """dftk_b200_ewald (real-space and reciprocal-space lattice sums as one kernel each) against the oracle's
    energy_forces_ewald and the reference's golden Ewald energies (test/ewald.jl:1-52)."""
    import dftk_b200 as dftk
    from gpu_common import ctx
    from oracle import forces as oforces
    from silicon import LATTICE, POSITIONS
    c = ctx()
    for lat_g, ch, pos_g, ref, tol in [(16 * np.eye(3), [1], [[0, 0, 0]], -0.088665545, 1e-8),
                                        (LATTICE, [14, 14], POSITIONS, -102.8741963352893, 1e-8),
                                        (16 * np.eye(3), [5, 5], [[0, 0, 0], [0.14763485355139283, 0, 0]], 1.790634595, 1e-7)]:
        pg = [np.array(q, dtype=float) for q in pos_g]
        e, f = dftk.energy_forces_ewald_device(c, lat_g, ch, pg)
        assert e == pytest.approx(ref, abs=tol)
        eo, fo = oforces.energy_forces_ewald(np.asarray(lat_g, dtype=float), ch, pg)
        assert e == pytest.approx(eo, abs=1e-11 * max(1, abs(eo)))
        np.testing.assert_allclose(np.array(f), np.array(fo), atol=1e-11 * max(1.0, np.abs(np.array(fo)).max()))
    # a low-symmetry 5-atom, two-species cell with a skewed lattice
    rng = np.random.default_rng(3)
    lat = np.array([[7.0, 0.4, -0.3], [0.2, 8.0, 0.5], [-0.1, 0.3, 9.5]])
    pos = [rng.random(3) for _ in range(5)]
    ch = [4, 3, 4, 1, 3]
    e, f = dftk.energy_forces_ewald_device(c, lat, ch, pos)
    eo, fo = oforces.energy_forces_ewald(lat, ch, pos)
    assert e == pytest.approx(eo, abs=1e-11 * abs(eo))
    np.testing.assert_allclose(np.array(f), np.array(fo), atol=1e-10)
    assert np.abs(np.sum(np.array(f), axis=0)).max() < 1e-9            # translation invariance
