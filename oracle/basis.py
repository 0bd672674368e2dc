"""Model / PlaneWaveBasis / Kpoint / FFT conventions (oracle; test infrastructure only).

Array convention: a cube of Julia shape (nx,ny,nz) (column-major) is stored as a NumPy array of
shape (nz,ny,nx) in C order, so that `cube.reshape(-1)[i]` is Julia's linear index i+1.
"""
import itertools
import math
import numpy as np
from .psp_hgh import PspHgh

SYMMETRY_TOLERANCE = 1e-5  # src/SymOp.jl:21 default


# ---------------------------------------------------------------- structure.jl:24-61
def compute_recip_lattice(lattice):
    return 2 * math.pi * np.linalg.inv(lattice.T)


def estimate_integer_lattice_bounds(M, delta, shift=(0, 0, 0), tol=math.sqrt(np.finfo(float).eps)):
    inv_lattice_t = np.linalg.inv(M.T)
    xlims = [np.linalg.norm(inv_lattice_t[:, i]) * delta + shift[i] for i in range(3)]
    return [0 if x == 0 else int(math.ceil(x - tol)) for x in xlims]


# ---------------------------------------------------------------- fft.jl:231-337
def next_compatible_fft_size(size, smallprimes=(2, 3, 5), factors=(1,)):
    def is_product_of_primes(n):
        for p in smallprimes:
            while n % p == 0:
                n //= p
        return n == 1 or not smallprimes
    fac = int(np.prod(factors))
    while not (size % fac == 0 and is_product_of_primes(size)):
        size += 1
    return size


def compute_fft_size(lattice, Ecut, supersampling=2.0, factors=(1,)):
    Gmax = supersampling * math.sqrt(2 * Ecut)
    Glims = estimate_integer_lattice_bounds(compute_recip_lattice(lattice), Gmax)
    return tuple(next_compatible_fft_size(2 * g + 1, factors=factors) for g in Glims)


def G_axis(n):
    """fft.jl:24-31: [0..floor((n-1)/2), -ceil((n-1)/2)..-1]"""
    stop = (n - 1) // 2
    start = -((n - 1 + 1) // 2)
    return np.array(list(range(0, stop + 1)) + list(range(start, 0)), dtype=np.int64)


def G_vectors(fft_size):
    """All G of the cube in Julia linear-index order: (N,3) int array."""
    nx, ny, nz = fft_size
    gx, gy, gz = G_axis(nx), G_axis(ny), G_axis(nz)
    Z, Y, X = np.meshgrid(gz, gy, gx, indexing="ij")
    return np.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], axis=1)


def index_G_vectors(fft_size, G):
    """PlaneWaveBasis.jl:465-480: linear index (0-based) of integer G (…,3) or -1."""
    G = np.asarray(G)
    n = np.array(fft_size)
    start = -((n - 1 + 1) // 2)
    stop = (n - 1) // 2
    ok = np.all((G >= start) & (G <= stop), axis=-1)
    idx3 = np.mod(G, n)
    lin = idx3[..., 0] + n[0] * (idx3[..., 1] + n[1] * idx3[..., 2])
    return np.where(ok, lin, -1)


# ---------------------------------------------------------------- symmetry (brute force)
class SymOp:
    """src/SymOp.jl: real-space op (W,w): x -> W x + w (reduced coords);
    reciprocal S = W', τ = -W^-1 w."""

    def __init__(self, W, w):
        self.W = np.array(W, dtype=np.int64)
        self.w = np.array(w, dtype=float)
        self.S = self.W.T.copy()
        self.tau = -np.linalg.solve(self.W.astype(float), self.w)

    def isone(self):
        return np.array_equal(self.W, np.eye(3, dtype=np.int64)) and np.allclose(self.w, 0)


def symmetry_operations(lattice, species, positions, magmoms=None, tol=SYMMETRY_TOLERANCE):
    """Brute-force replacement for spglib (src/symmetry.jl:91-120): all (W,w) with integer W
    preserving the metric and mapping the decorated atom set onto itself."""
    metric = lattice.T @ lattice
    positions = [np.asarray(p, dtype=float) for p in positions]
    labels = list(species)
    if magmoms is not None and len(magmoms):
        labels = [(s, round(float(m), 6)) for s, m in zip(species, magmoms)]
    cand = []
    rng = (-1, 0, 1)
    # entries beyond ±1 are not needed for reduced cells; extend the range for odd cells
    for e in itertools.product(rng, repeat=9):
        W = np.array(e, dtype=np.int64).reshape(3, 3)
        if abs(round(np.linalg.det(W))) != 1:
            continue
        if np.allclose(W.T @ metric @ W, metric, atol=tol * np.max(np.abs(metric))):
            cand.append(W)
    ops = []
    ref = 0  # map atom 0 to each atom of the same label
    for W in cand:
        for j, pj in enumerate(positions):
            if labels[j] != labels[ref]:
                continue
            w = pj - W @ positions[ref]
            w = w - np.round(w)
            ok = True
            for a, pa in enumerate(positions):
                img = W @ pa + w
                found = False
                for b, pb in enumerate(positions):
                    if labels[a] != labels[b]:
                        continue
                    d = img - pb
                    d = d - np.round(d)
                    if np.max(np.abs(d)) < tol:
                        found = True
                        break
                if not found:
                    ok = False
                    break
            if ok:
                w = np.where(np.abs(w) < tol, 0.0, w)
                if not any(np.array_equal(W, o.W) and np.allclose((w - o.w) - np.round(w - o.w), 0, atol=tol)
                           for o in ops):
                    ops.append(SymOp(W, w))
    # identity first
    ops.sort(key=lambda o: (not o.isone(),))
    return ops


def normalize_kpoint_coordinate(k):
    """bzmesh.jl:4-10: into [-0.5, 0.5)."""
    k = np.asarray(k, dtype=float)
    k = k - np.floor(k + 0.5)
    k = np.where(k >= 0.5, k - 1.0, k)
    return k


def reducible_kcoords(kgrid_size, kshift=(0, 0, 0)):
    """bzmesh.jl:41-48."""
    ks = np.array(kgrid_size)
    start = -np.floor((ks - 1) / 2).astype(int)
    stop = np.ceil((ks - 1) / 2).astype(int)
    out = []
    for k in range(start[2], stop[2] + 1):
        for j in range(start[1], stop[1] + 1):
            for i in range(start[0], stop[0] + 1):
                out.append(normalize_kpoint_coordinate((np.array(kshift, dtype=float)
                                                        + np.array([i, j, k])) / ks))
    return out


def irreducible_kcoords(kgrid_size, kshift, symmetries):
    """bzmesh.jl:55-95 without spglib: orbit reduction of the MP grid under S (no time reversal,
    as the reference passes is_time_reversal=false)."""
    kall = reducible_kcoords(kgrid_size, kshift)
    n = len(kall)
    key = lambda k: tuple(np.round(normalize_kpoint_coordinate(k) * 1e6).astype(np.int64) % 1000000)
    index = {key(k): i for i, k in enumerate(kall)}
    rep = -np.ones(n, dtype=int)
    kirr, counts = [], []
    for i, k in enumerate(kall):
        if rep[i] >= 0:
            continue
        rep[i] = len(kirr)
        cnt = 1
        for op in symmetries:
            j = index.get(key(op.S @ k))
            if j is not None and rep[j] < 0:
                rep[j] = len(kirr)
                cnt += 1
        kirr.append(k)
        counts.append(cnt)
    assert sum(counts) == n
    return kirr, [c / n for c in counts]


# ---------------------------------------------------------------- Model
class Element:
    def __init__(self, symbol, psp=None, functional="lda"):
        self.symbol = symbol
        self.psp = psp if psp is not None else PspHgh.from_table(symbol, functional)
        self.Z = self.psp.Z

    @property
    def charge_ionic(self):
        return self.psp.Zion

    @property
    def n_elec_valence(self):
        return self.psp.Zion

    @property
    def n_elec_core(self):
        return self.Z - self.psp.Zion


class Model:
    """src/Model.jl:128-219 + standard_models.jl:45-60 (model_atomic/model_DFT term lists)."""

    def __init__(self, lattice, atoms, positions, terms=("Kinetic", "AtomicLocal", "AtomicNonlocal",
                                                        "Ewald", "PspCorrection", "Hartree", "Xc"),
                 functionals=("lda_x", "lda_c_pw"), temperature=0.0, smearing=None,
                 magnetic_moments=(), symmetries=True, n_electrons=None):
        self.lattice = np.array(lattice, dtype=float)
        self.atoms = list(atoms)
        self.positions = [np.array(p, dtype=float) for p in positions]
        self.recip_lattice = compute_recip_lattice(self.lattice)
        self.unit_cell_volume = abs(np.linalg.det(self.lattice))
        self.n_electrons = (sum(a.n_elec_valence for a in self.atoms)
                            if n_electrons is None else n_electrons)
        self.temperature = float(temperature)
        self.smearing = smearing if smearing is not None else (
            "FermiDirac" if temperature > 0 else "None")
        self.magnetic_moments = list(magnetic_moments)
        self.spin_polarization = "collinear" if len(self.magnetic_moments) else "none"
        self.n_spin_components = 2 if self.spin_polarization == "collinear" else 1
        self.terms = list(terms)
        if self.temperature > 0 and "Entropy" not in self.terms:
            self.terms.append("Entropy")
        self.functionals = list(functionals)
        # atom groups: identical elements (Model.jl:170)
        self.atom_groups = []
        seen = {}
        for i, a in enumerate(self.atoms):
            k = (a.symbol, id(a.psp) if False else a.psp.description)
            seen.setdefault(k, []).append(i)
        self.atom_groups = list(seen.values())
        if symmetries is True:
            self.symmetries = symmetry_operations(self.lattice, [a.symbol for a in self.atoms],
                                                  self.positions, self.magnetic_moments)
        elif symmetries is False:
            self.symmetries = [SymOp(np.eye(3), np.zeros(3))]
        else:
            self.symmetries = symmetries

    @property
    def filled_occupation(self):
        return 2 if self.spin_polarization == "none" else 1


# ---------------------------------------------------------------- Kpoint / basis
class Kpoint:
    """src/Kpoint.jl:6-41."""

    def __init__(self, spin, coordinate, recip_lattice, fft_size, Ecut, Gall=None):
        self.spin = spin
        self.coordinate = np.array(coordinate, dtype=float)
        Gall = G_vectors(fft_size) if Gall is None else Gall
        p = (Gall + self.coordinate) @ recip_lattice.T
        keep = np.sum(p * p, axis=1) / 2 <= Ecut
        self.mapping = np.nonzero(keep)[0].astype(np.int64)  # 0-based, ascending
        self.G_vectors = Gall[self.mapping]

    @property
    def n_G(self):
        return len(self.mapping)


class PlaneWaveBasis:
    """src/PlaneWaveBasis.jl:129-369 (single process; k-sharding handled by callers)."""

    def __init__(self, model, Ecut, kgrid=(1, 1, 1), kshift=(0, 0, 0), fft_size=None,
                 kcoords=None, kweights=None, use_symmetries_for_kpoint_reduction=True,
                 supersampling=2.0):
        self.model = model
        self.Ecut = float(Ecut)
        symmetries_respect_rgrid = fft_size is None
        if fft_size is None:
            factors = (1,)
            if symmetries_respect_rgrid:
                dens = set()
                for s in model.symmetries:
                    for wi in s.w:
                        from fractions import Fraction
                        dens.add(Fraction(wi).limit_denominator(12).denominator)
                fs = sorted(set((2, 3, 4, 6)) & dens)
                factors = tuple(fs) if fs else (1,)
            fft_size = compute_fft_size(model.lattice, Ecut, supersampling, factors)
        self.fft_size = tuple(int(n) for n in fft_size)
        self.N = int(np.prod(self.fft_size))
        self.dvol = model.unit_cell_volume / self.N
        self.ifft_normalization = 1 / math.sqrt(model.unit_cell_volume)
        self.fft_normalization = math.sqrt(model.unit_cell_volume) / self.N
        self.G_all = G_vectors(self.fft_size)
        self.G_cart = self.G_all @ model.recip_lattice.T
        # symmetries preserving grids (symmetry.jl symmetries_preserving_*)
        syms = model.symmetries
        if symmetries_respect_rgrid:
            n = np.array(self.fft_size)
            syms = [s for s in syms if np.allclose(s.w * n, np.round(s.w * n), atol=SYMMETRY_TOLERANCE)]
        if kcoords is None:
            kall = reducible_kcoords(kgrid, kshift)
            keyset = {tuple(np.round(k * 1e6).astype(np.int64) % 1000000) for k in kall}
            def preserves(s):
                return all(tuple(np.round(normalize_kpoint_coordinate(s.S @ k) * 1e6).astype(np.int64) % 1000000)
                           in keyset for k in kall)
            syms = [s for s in syms if preserves(s)]
            if use_symmetries_for_kpoint_reduction:
                kcoords, kweights = irreducible_kcoords(kgrid, kshift, syms)
            else:
                kcoords, kweights = kall, [1.0 / len(kall)] * len(kall)
        self.symmetries = syms
        self.kcoords_global = [np.array(k, dtype=float) for k in kcoords]
        self.kweights_global = list(kweights)
        self.kpoints, self.kweights = [], []
        for spin in range(model.n_spin_components):
            for k, w in zip(self.kcoords_global, self.kweights_global):
                self.kpoints.append(Kpoint(spin, k, model.recip_lattice, self.fft_size, Ecut, self.G_all))
                self.kweights.append(w)
        self.terms = None

    # ---- FFTs (fft.jl:106-172) ----
    def cube(self, flat):
        nx, ny, nz = self.fft_size
        return flat.reshape(nz, ny, nx)

    def ifft_cube(self, f_fourier_flat):
        """fft.jl:106-109: f_real = opBFFT*f * ifft_normalization."""
        return (np.fft.ifftn(self.cube(f_fourier_flat)) * self.N * self.ifft_normalization).reshape(-1)

    def irfft_cube(self, f):
        return np.real(self.ifft_cube(f))

    def fft_cube(self, f_real_flat):
        """fft.jl:155-161."""
        return (np.fft.fftn(self.cube(f_real_flat.astype(complex))) * self.fft_normalization).reshape(-1)

    def ifft_kpt(self, kpt, f_fourier, normalize=True):
        """fft.jl:110-122: zero-pad the sphere into the cube, unnormalised backward FFT."""
        c = np.zeros(self.N, dtype=complex)
        c[kpt.mapping] = f_fourier
        out = (np.fft.ifftn(self.cube(c)) * self.N).reshape(-1)
        return out * self.ifft_normalization if normalize else out

    def fft_kpt(self, kpt, f_real, normalize=True):
        """fft.jl:162-172."""
        c = np.fft.fftn(self.cube(f_real)).reshape(-1)[kpt.mapping]
        return c * self.fft_normalization if normalize else c

    def enforce_real(self, coeffs):
        """symmetry.jl:550-552 via lowpass_for_symmetry! with S=-I."""
        idx = index_G_vectors(self.fft_size, -self.G_all)
        out = coeffs.copy()
        out[idx < 0] = 0
        return out

    def Gplusk_cart(self, kpt):
        return (kpt.G_vectors + kpt.coordinate) @ self.model.recip_lattice.T
