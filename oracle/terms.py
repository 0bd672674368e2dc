"""Energy terms, Hamiltonian blocks and the Hψ apply (oracle; test infrastructure only)."""
import math
import numpy as np
from scipy.special import erfc
from .basis import estimate_integer_lattice_bounds, compute_recip_lattice, index_G_vectors
from .psp_hgh import solid_harmonic_real, atom_decay_length
from . import xc as xcmod


# ------------------------------------------------------------------ kinetic (kinetic.jl:24-35)
def kinetic_energies(basis, kpt):
    p = basis.Gplusk_cart(kpt)
    return np.sum(p * p, axis=1) / 2


# ------------------------------------------------------------------ local (local.jl:108-138)
def structure_factor_cube(basis, r):
    """exp(-2πi G·r) for every G of the cube (Julia linear order), built as an outer product of the three
    per-axis phase vectors (identical values to cis2pi(-dot(G, r)), local.jl:127)."""
    from .basis import G_axis
    nx, ny, nz = basis.fft_size
    ex = np.exp(-2j * math.pi * G_axis(nx) * r[0])
    ey = np.exp(-2j * math.pi * G_axis(ny) * r[1])
    ez = np.exp(-2j * math.pi * G_axis(nz) * r[2])
    return (ez[:, None, None] * ey[None, :, None] * ex[None, None, :]).reshape(-1)


def compute_local_potential(basis):
    model = basis.model
    pnorm = np.sqrt(np.sum(basis.G_cart ** 2, axis=1))
    pot = np.zeros(basis.N, dtype=complex)
    for group in model.atom_groups:
        ff = model.atoms[group[0]].psp.eval_local_fourier(pnorm)
        for ia in group:
            r = model.positions[ia]
            pot += structure_factor_cube(basis, r) * ff / math.sqrt(model.unit_cell_volume)
    pot = basis.enforce_real(pot)
    return basis.irfft_cube(pot)


# ------------------------------------------------------------------ nonlocal (nonlocal.jl:107-244)
def build_projection_coefficients(psp):
    n = psp.n_proj()
    D = np.zeros((n, n))
    count = 0
    for l in range(psp.lmax + 1):
        for _m in range(-l, l + 1):
            nl = psp.n_proj_radial(l)
            D[count:count + nl, count:count + nl] = psp.h[l]
            count += nl
    return D


def build_projector_form_factors(psp, Gpk_cart):
    """(n_G, n_proj) complex; ordering (l, m, i) per nonlocal.jl:140-141,205-244."""
    pn = np.sqrt(np.sum(Gpk_cart ** 2, axis=1))
    cols = []
    for l in range(psp.lmax + 1):
        nl = psp.n_proj_radial(l)
        for m in range(-l, l + 1):
            ang = ((-1j) ** l) * solid_harmonic_real(l, m, Gpk_cart)
            for i in range(1, nl + 1):
                cols.append(psp.eval_projector_fourier(i, l, pn) * ang)
    if not cols:
        return np.zeros((len(pn), 0), dtype=complex)
    return np.stack(cols, axis=1).astype(complex)


def build_projection_vectors(basis, kpt):
    model = basis.model
    Gpk = kpt.G_vectors + kpt.coordinate
    Gpk_cart = basis.Gplusk_cart(kpt)
    blocks, Ds = [], []
    for group in model.atom_groups:
        psp = model.atoms[group[0]].psp
        ff = build_projector_form_factors(psp, Gpk_cart)
        Dat = build_projection_coefficients(psp)
        for ia in group:
            sf = np.exp(-2j * math.pi * (Gpk @ model.positions[ia]))
            blocks.append(sf[:, None] * ff / math.sqrt(model.unit_cell_volume))
            Ds.append(Dat)
    P = np.concatenate(blocks, axis=1)
    n = P.shape[1]
    D = np.zeros((n, n))
    o = 0
    for Dat in Ds:
        k = Dat.shape[0]
        D[o:o + k, o:o + k] = Dat
        o += k
    return P, D


# ------------------------------------------------------------------ Ewald (ewald.jl:40-168)
def energy_ewald(lattice, charges, positions, eta=None):
    charges = np.asarray(charges, dtype=float)
    positions = [np.asarray(p, dtype=float) for p in positions]
    recip = compute_recip_lattice(lattice)
    if eta is None:
        eta = math.sqrt(math.sqrt(1.69 * np.linalg.norm(recip / (2 * math.pi)) / np.linalg.norm(lattice))) / 2
    eps = np.finfo(float).eps
    max_exp_arg = -math.log(eps) + 5
    max_erfc_arg = math.sqrt(max_exp_arg)
    Glims = estimate_integer_lattice_bounds(recip, math.sqrt(max_exp_arg) * 2 * eta)
    poslims = [max(rj[i] - rk[i] for rj in positions for rk in positions) for i in range(3)]
    Rlims = estimate_integer_lattice_bounds(lattice, max_erfc_arg / eta, poslims)
    vol = abs(np.linalg.det(lattice))
    pos = np.array(positions)

    sum_recip = -(charges.sum() ** 2 / (4 * eta ** 2))
    gr = [np.arange(-g, g + 1) for g in Glims]
    G = np.stack(np.meshgrid(*gr, indexing="ij"), axis=-1).reshape(-1, 3)
    G = G[np.any(G != 0, axis=1)]
    Gsq = np.sum((G @ recip.T) ** 2, axis=1)
    ph = 2 * math.pi * (G @ pos.T)
    cs = (np.cos(ph) * charges).sum(axis=1)
    sn = (np.sin(ph) * charges).sum(axis=1)
    sum_recip += np.sum((cs ** 2 + sn ** 2) * np.exp(-Gsq / (4 * eta ** 2)) / Gsq)
    sum_recip *= 4 * math.pi / vol

    sum_real = -2 * eta / math.sqrt(math.pi) * np.sum(charges ** 2)
    rr = [np.arange(-g, g + 1) for g in Rlims]
    R = np.stack(np.meshgrid(*rr, indexing="ij"), axis=-1).reshape(-1, 3).astype(float)
    for i in range(len(pos)):
        for j in range(len(pos)):
            d = (pos[i] - pos[j] - R) @ lattice.T
            dist = np.sqrt(np.sum(d * d, axis=1))
            if i == j:
                dist = dist[np.any(R != 0, axis=1)]
            sum_real += np.sum(charges[i] * charges[j] * erfc(eta * dist) / dist)
    return (sum_recip + sum_real) / 2


def energy_psp_correction(model):
    """psp_correction.jl:26-35."""
    corr = sum(len(g) * model.atoms[g[0]].psp.energy_correction() for g in model.atom_groups)
    return corr * sum(a.n_elec_valence for a in model.atoms) / model.unit_cell_volume


# ------------------------------------------------------------------ guess density
def guess_density(basis, magnetic_moments=None):
    """density_methods.jl:103-181,237-244 with ValenceDensityGaussian (HGH has no valence density)."""
    model = basis.model
    pn = np.sqrt(np.sum(basis.G_cart ** 2, axis=1))

    def superposition(coeffs):
        rho = np.zeros(basis.N, dtype=complex)
        for ia, atom in enumerate(model.atoms):
            L = atom_decay_length(atom.n_elec_core, atom.n_elec_valence)
            ff = atom.charge_ionic * np.exp(-(pn * L) ** 2)
            rho += (structure_factor_cube(basis, model.positions[ia]) * ff
                    * (coeffs[ia] / math.sqrt(model.unit_cell_volume)))
        return basis.irfft_cube(basis.enforce_real(rho))

    rtot = superposition(np.ones(len(model.atoms)))
    if model.n_spin_components == 1:
        rho = rtot[None, :]
    else:
        mm = magnetic_moments if magnetic_moments is not None else model.magnetic_moments
        coeffs = [m / a.n_elec_valence for m, a in zip(mm, model.atoms)]
        rspin = superposition(coeffs) if any(c != 0 for c in coeffs) else np.zeros(basis.N)
        rho = np.stack([(rtot + rspin) / 2, (rtot - rspin) / 2])
    Nel = rho.sum() * basis.dvol
    if Nel > 0:
        rho = rho * (model.n_electrons / Nel)
    return rho


# ------------------------------------------------------------------ Hartree / XC
def poisson_green_coeffs(basis):
    G2 = np.sum(basis.G_cart ** 2, axis=1)
    with np.errstate(divide="ignore"):
        g = 4 * math.pi / G2
    g[0] = 0.0
    return basis.enforce_real(g)


def hartree(basis, green, rho):
    """hartree.jl:50-59."""
    rf = basis.fft_cube(rho.sum(axis=0))
    pf = green * rf
    return float(np.real(np.vdot(pf, rf)) / 2), basis.irfft_cube(pf)


def xc_potential(basis, rho):
    """xc.jl:84-160 + LibxcDensities xc.jl:356-409 + divergence_real :576-584."""
    model = basis.model
    fun = model.functionals
    n_spin = model.n_spin_components
    is_gga = any(f.startswith("gga") for f in fun)
    sigma = grad = None
    if is_gga:
        grad = np.zeros((n_spin, 3, basis.N))
        for s in range(n_spin):
            rf = basis.fft_cube(rho[s])
            for a in range(3):
                grad[s, a] = basis.irfft_cube(1j * basis.G_cart[:, a] * rf)
        if n_spin == 1:
            sigma = np.sum(grad[0] * grad[0], axis=0)[None, :]
        else:
            sigma = np.stack([np.sum(grad[0] * grad[0], axis=0), np.sum(grad[0] * grad[1], axis=0),
                              np.sum(grad[1] * grad[1], axis=0)])
    res = xcmod.evaluate(fun, rho, sigma)
    E = float(res["e"].sum() * basis.dvol)
    pot = res["Vrho"].copy()
    if is_gga:
        ts = (lambda s, t: 0) if n_spin == 1 else (lambda s, t: {(0, 0): 0, (0, 1): 1, (1, 0): 1, (1, 1): 2}[(s, t)])
        for s in range(n_spin):
            gsum = np.zeros(basis.N, dtype=complex)
            for a in range(3):
                op = np.zeros(basis.N)
                for t in range(n_spin):
                    op += (1.0 if s == t else 0.5) * res["Vsigma"][ts(s, t)] * grad[t, a]
                gsum += 1j * basis.G_cart[:, a] * basis.fft_cube(op)
            pot[s] += -2 * basis.irfft_cube(gsum)
    return E, pot


# ------------------------------------------------------------------ smearing (Smearing.jl:71-138)
def smearing_occupation(kind, x):
    x = np.asarray(x, dtype=float)
    if kind == "None":
        return np.where(x > 0, 0.0, 1.0)
    if kind == "FermiDirac":
        with np.errstate(over="ignore"):
            return np.where(x > 0, np.exp(-np.abs(x)) / (1 + np.exp(-np.abs(x))), 1 / (1 + np.exp(np.minimum(x, 0))))
    if kind == "Gaussian":
        return erfc(x) / 2
    raise NotImplementedError(kind)


def smearing_entropy(kind, x):
    x = np.asarray(x, dtype=float)
    if kind == "None":
        return np.zeros_like(x)
    if kind == "FermiDirac":
        f = smearing_occupation(kind, x)
        eps = np.finfo(float).eps
        out = np.zeros_like(x)
        ok = (np.abs(f) >= eps) & (np.abs(1 - f) >= eps)
        fo = f[ok]
        out[ok] = -(fo * np.log(fo) + (1 - fo) * np.log(1 - fo))
        return out
    if kind == "Gaussian":
        return 1 / (2 * math.sqrt(math.pi)) * np.exp(-x ** 2)
    raise NotImplementedError(kind)


# ------------------------------------------------------------------ terms container
class Terms:
    """Instantiated terms for a basis (PlaneWaveBasis.jl:256-259 loop over model.term_types)."""

    def __init__(self, basis):
        self.basis = basis
        m = basis.model
        t = m.terms
        self.kin = [kinetic_energies(basis, k) for k in basis.kpoints] if "Kinetic" in t else None
        self.Vloc = compute_local_potential(basis) if "AtomicLocal" in t else None
        self.PD = None
        if "AtomicNonlocal" in t:
            nk = len(basis.kcoords_global)
            first = [build_projection_vectors(basis, k) for k in basis.kpoints[:nk]]
            self.PD = [first[i % nk] for i in range(len(basis.kpoints))]
        self.E_ewald = energy_ewald(m.lattice, [a.charge_ionic for a in m.atoms], m.positions) if "Ewald" in t else None
        self.E_pspcorr = energy_psp_correction(m) if "PspCorrection" in t else None
        self.green = poisson_green_coeffs(basis) if "Hartree" in t else None


class HamiltonianBlock:
    """DftHamiltonianBlock + mul! (Hamiltonian.jl:22-57,137-192)."""

    def __init__(self, basis, ik, kin, Vtot, PD):
        self.basis, self.ik, self.kpt = basis, ik, basis.kpoints[ik]
        self.kin, self.Vtot, self.PD = kin, Vtot, PD
        self.shape = (self.kpt.n_G, self.kpt.n_G)

    def local_apply(self, psi):
        """Band-at-a-time ifft -> ·V -> fft (Hamiltonian.jl:152-163)."""
        b, kpt = self.basis, self.kpt
        out = np.empty_like(psi)
        pot = self.Vtot * (b.fft_normalization * b.ifft_normalization)

        def one(n):
            pr = b.ifft_kpt(kpt, psi[:, n], normalize=False)
            pr *= pot
            out[:, n] = b.fft_kpt(kpt, pr, normalize=False)
        workers = getattr(self, "workers", 1)
        if workers > 1 and psi.shape[1] > 1:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(workers) as ex:
                list(ex.map(one, range(psi.shape[1])))
        else:
            for n in range(psi.shape[1]):
                one(n)
        return out

    def matmul(self, psi):
        if psi.shape[1] == 0:
            return np.zeros_like(psi)
        if self.Vtot is not None:
            H = self.local_apply(psi)
        else:
            H = np.zeros_like(psi)
        if self.kin is not None:
            H += self.kin[:, None] * psi                        # :180
        if self.PD is not None:
            P, D = self.PD
            H += P @ (D @ (P.conj().T @ psi))                    # operators.jl:126-128
        return H

    __matmul__ = matmul


def energy_hamiltonian(basis, terms, psi, occupation, rho, eigenvalues=None, eF=None, only_energy=False):
    """Hamiltonian.jl:200-236.  rho: (n_spin, N).  Returns (energies dict, [HamiltonianBlock])."""
    m = basis.model
    E = {}
    have = psi is not None and occupation is not None
    Vtot = np.zeros((m.n_spin_components, basis.N)) if (terms.Vloc is not None or terms.green is not None
                                                         or "Xc" in m.terms) else None
    if terms.kin is not None:
        if have:
            E["Kinetic"] = sum(basis.kweights[ik] * np.sum(occupation[ik] * np.real(
                np.sum(np.conj(psi[ik]) * (terms.kin[ik][:, None] * psi[ik]), axis=0)))
                for ik in range(len(basis.kpoints)))
        else:
            E["Kinetic"] = math.inf
    if terms.Vloc is not None:
        E["AtomicLocal"] = float(np.sum(rho.sum(axis=0) * terms.Vloc) * basis.dvol)
        Vtot += terms.Vloc[None, :]
    if terms.PD is not None:
        if have:
            e = 0.0
            for ik in range(len(basis.kpoints)):
                P, D = terms.PD[ik]
                Ppsi = P.conj().T @ psi[ik]
                be = np.sum(np.real(np.conj(Ppsi) * (D @ Ppsi)), axis=0)
                e += basis.kweights[ik] * np.sum(be * occupation[ik])
            E["AtomicNonlocal"] = float(e)
        else:
            E["AtomicNonlocal"] = math.inf
    if terms.E_ewald is not None:
        E["Ewald"] = terms.E_ewald
    if terms.E_pspcorr is not None:
        E["PspCorrection"] = terms.E_pspcorr
    if terms.green is not None:
        eh, vh = hartree(basis, terms.green, rho)
        E["Hartree"] = eh
        Vtot += vh[None, :]
    if "Xc" in m.terms:
        exc, vxc = xc_potential(basis, rho)
        E["Xc"] = exc
        Vtot += vxc
    if "Entropy" in m.terms:
        if m.temperature == 0:
            E["Entropy"] = 0.0
        elif have and eigenvalues is not None and eF is not None:
            e = 0.0
            for ik in range(len(basis.kpoints)):
                e -= (m.temperature * basis.kweights[ik] * m.filled_occupation
                      * np.sum(smearing_entropy(m.smearing, (eigenvalues[ik][:psi[ik].shape[1]] - eF) / m.temperature)))
            E["Entropy"] = float(e)
        else:
            E["Entropy"] = math.inf
    E["total"] = sum(v for k, v in E.items())
    if only_energy:
        return E, None
    blocks = [HamiltonianBlock(basis, ik,
                               terms.kin[ik] if terms.kin is not None else None,
                               Vtot[kpt.spin] if Vtot is not None else None,
                               terms.PD[ik] if terms.PD is not None else None)
              for ik, kpt in enumerate(basis.kpoints)]
    return E, blocks
