"""Hellmann-Feynman forces (oracle; test infrastructure only).

Restates src/postprocess/forces.jl:24-47 and the per-term `compute_forces` methods:
local.jl:147-181 (forces_local), nonlocal.jl:49-100, ewald.jl:64-168 (energy_forces_ewald),
symmetry.jl:379-423 (find_symmetry_preimage, symmetrize_forces).  Forces are returned in reduced
(fractional) coordinates like the reference; `forces_cart` converts (covector_red_to_cart).
Kinetic, Hartree, Xc (no non-linear core correction in the HGH tables), PspCorrection and Entropy have no force.
"""
import math
import numpy as np
from scipy.special import erfc
from .basis import estimate_integer_lattice_bounds, compute_recip_lattice, SYMMETRY_TOLERANCE
from .terms import build_projection_coefficients, build_projector_form_factors


def forces_local(basis, rho):
    """local.jl:152-181 with q = 0: F_a,α = -Re( Σ_G -2πi G_α e^{-2πi G·r_a} conj(ρ_G) v_loc(|G|) ) / sqrt(Ω)."""
    model = basis.model
    rho_f = basis.fft_cube(rho.sum(axis=0))
    pnorm = np.sqrt(np.sum(basis.G_cart ** 2, axis=1))
    G = basis.G_all.astype(float)
    F = [np.zeros(3) for _ in model.positions]
    for group in model.atom_groups:
        ff = model.atoms[group[0]].psp.eval_local_fourier(pnorm)
        for ia in group:
            r = model.positions[ia]
            rho_pot = np.exp(-2j * math.pi * (G @ r)) * np.conj(rho_f) * ff
            for a in range(3):
                tmp = np.sum(-2j * math.pi * G[:, a] * rho_pot)
                F[ia][a] += -np.real(tmp / math.sqrt(model.unit_cell_volume))
    return F


def forces_nonlocal(basis, psi, occupation):
    """nonlocal.jl:49-100: for each atom and direction, δHψ = P D (dP/dR_α)†ψ with dP/dR_α = -2πi (G+k)_α P,
    F = -w_k Σ_n f_n 2 Re<ψ_n|δHψ_n>; then symmetrised."""
    model = basis.model
    F = [np.zeros(3) for _ in model.positions]
    for group in model.atom_groups:
        psp = model.atoms[group[0]].psp
        if psp.n_proj() == 0:
            continue
        D = build_projection_coefficients(psp)
        for ik, kpt in enumerate(basis.kpoints):
            Gpk = kpt.G_vectors + kpt.coordinate
            ff = build_projector_form_factors(psp, basis.Gplusk_cart(kpt))
            for ia in group:
                r = model.positions[ia]
                sf = np.exp(-2j * math.pi * (Gpk @ r))
                P = sf[:, None] * ff / math.sqrt(model.unit_cell_volume)
                for a in range(3):
                    dPdR = (-2j * math.pi * Gpk[:, a])[:, None] * P
                    dH = P @ (D @ (dPdR.conj().T @ psi[ik]))
                    dots = np.sum(np.conj(psi[ik]) * dH, axis=0)
                    F[ia][a] += -basis.kweights[ik] * np.sum(occupation[ik] * 2 * np.real(dots))
    return symmetrize_forces(basis, F)


def energy_forces_ewald(lattice, charges, positions, eta=None):
    """ewald.jl:64-168 (q = 0, no phonon displacement)."""
    lattice = np.asarray(lattice, dtype=float)
    charges = np.asarray(charges, dtype=float)
    pos = np.array([np.asarray(p, dtype=float) for p in positions])
    recip = compute_recip_lattice(lattice)
    if eta is None:
        eta = math.sqrt(math.sqrt(1.69 * np.linalg.norm(recip / (2 * math.pi)) / np.linalg.norm(lattice))) / 2
    max_exp_arg = -math.log(np.finfo(float).eps) + 5
    max_erfc_arg = math.sqrt(max_exp_arg)
    Glims = estimate_integer_lattice_bounds(recip, math.sqrt(max_exp_arg) * 2 * eta)
    poslims = [float(np.max(pos[:, i][:, None] - pos[:, i][None, :])) for i in range(3)]
    Rlims = estimate_integer_lattice_bounds(lattice, max_erfc_arg / eta, poslims)
    vol = abs(np.linalg.det(lattice))
    n = len(pos)

    G = np.stack(np.meshgrid(*[np.arange(-g, g + 1) for g in Glims], indexing="ij"), axis=-1).reshape(-1, 3)
    G = G[np.any(G != 0, axis=1)].astype(float)
    Gsq = np.sum((G @ recip.T) ** 2, axis=1)
    ph = 2 * math.pi * (G @ pos.T)                           # (n_G, n_atoms)
    cosf, sinf = np.cos(ph), np.sin(ph)
    cs, sn = (cosf * charges).sum(axis=1), (sinf * charges).sum(axis=1)
    damp = np.exp(-Gsq / (4 * eta ** 2)) / Gsq
    sum_recip = -(charges.sum() ** 2 / (4 * eta ** 2)) + np.sum((cs ** 2 + sn ** 2) * damp)
    F_recip = np.zeros((n, 3))
    for i in range(n):
        # dsum = cs * dc + sn * ds with dc = -Z 2π G sin, ds = +Z 2π G cos
        coeff = charges[i] * 2 * math.pi * (-cs * sinf[:, i] + sn * cosf[:, i]) * damp
        F_recip[i] = -(coeff[:, None] * G).sum(axis=0)
    sum_recip *= 4 * math.pi / vol
    F_recip *= 4 * math.pi / vol

    sum_real = -2 * eta / math.sqrt(math.pi) * np.sum(charges ** 2)
    F_real = np.zeros((n, 3))
    R = np.stack(np.meshgrid(*[np.arange(-g, g + 1) for g in Rlims], indexing="ij"), axis=-1).reshape(-1, 3).astype(float)
    nonzero_R = np.any(R != 0, axis=1)
    for i in range(n):
        for j in range(n):
            d = (pos[i] - pos[j] - R) @ lattice.T            # Δr (n_R, 3)
            if i == j:
                d = d[nonzero_R]
            dist = np.sqrt(np.sum(d * d, axis=1))
            zz = charges[i] * charges[j]
            e = zz * erfc(eta * dist) / dist
            sum_real += np.sum(e)
            dE_ddist = (zz * eta * (-2 * np.exp(-(eta * dist) ** 2) / math.sqrt(math.pi)) - e) / dist
            F_real[i] -= lattice.T @ np.sum((dE_ddist / dist)[:, None] * d, axis=0)
    return (sum_recip + sum_real) / 2, [F_recip[i] + F_real[i] for i in range(n)]


def find_symmetry_preimage(positions_group, position, symop, tol=SYMMETRY_TOLERANCE):
    other = np.linalg.solve(symop.W, position - symop.w)
    dev = [np.max(np.abs((at - other) - np.round(at - other))) for at in positions_group]
    i = int(np.argmin(dev))
    assert dev[i] < 10 * tol
    return i


def symmetrize_forces(basis, forces, symmetries=None):
    """symmetry.jl:399-413: F_sym[a] = 1/n_sym Σ_s W_s^{-T} F[preimage_s(a)]."""
    model = basis.model
    symmetries = basis.symmetries if symmetries is None else symmetries
    out = [np.zeros(3) for _ in forces]
    for group in model.atom_groups:
        pg = [model.positions[i] for i in group]
        for s in symmetries:
            WinvT = np.linalg.inv(s.W.T)
            for idx, p in enumerate(pg):
                j = find_symmetry_preimage(pg, p, s)
                out[group[idx]] += WinvT @ forces[group[j]]
    return [f / len(symmetries) for f in out]


def compute_forces(basis, psi, occupation, rho):
    """forces.jl:24-30: sum of the per-term forces (reduced coordinates)."""
    m = basis.model
    total = [np.zeros(3) for _ in m.positions]
    parts = {}
    if "AtomicLocal" in m.terms:
        parts["AtomicLocal"] = forces_local(basis, rho)
    if "AtomicNonlocal" in m.terms:
        parts["AtomicNonlocal"] = forces_nonlocal(basis, psi, occupation)
    if "Ewald" in m.terms:
        parts["Ewald"] = energy_forces_ewald(m.lattice, [a.charge_ionic for a in m.atoms], m.positions)[1]
    for f in parts.values():
        for i in range(len(total)):
            total[i] = total[i] + f[i]
    return total, parts


def forces_cart(model, forces_reduced):
    """covector_red_to_cart: F_cart = inv(lattice)^T F_red."""
    inv_lat_T = np.linalg.inv(model.lattice).T
    return [inv_lat_T @ f for f in forces_reduced]
