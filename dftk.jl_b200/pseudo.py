"""Analytic GTH/HGH pseudopotentials for the host-side setup (mirror of src/pseudo/PspHgh.jl and
src/elements.jl ElementPsp).  Setup code: runs once per basis, vectorised with torch on the device."""
import math
import re
import numpy as np
import torch

# published GTH parameters (same numbers as the reference's data/psp/hgh/{lda,pbe}/*.hgh files)
_TABLE = {
    ("Si", "lda"): (14, [2, 2], 0.44, [-7.33610297],
                    [(0.42273813, [[5.90692831, -1.26189397], [3.25819622]]), (0.48427842, [[2.72701346]])]),
    ("Si", "pbe"): (14, [2, 2], 0.44, [-6.26928833],
                    [(0.43563383, [[8.95174150, -2.70627082], [3.49378060]]), (0.49794218, [[2.43127673]])]),
    ("Al", "lda"): (13, [2, 1], 0.45, [-8.49135116],
                    [(0.46010427, [[5.08833953, -1.03784325], [2.67969975]]), (0.53674439, [[2.19343827]])]),
    ("Al", "pbe"): (13, [2, 1], 0.45, [-7.55476126],
                    [(0.48743529, [[6.95993832, -1.88883584], [2.43847659]]), (0.56218949, [[1.86529857]])]),
    ("Fe", "pbe"): (26, [4, 6, 6], 0.36, [6.75678916, -0.22883251],
                    [(0.27826303, [[0.62950570, 7.91313242], [-10.21581002]]),
                     (0.25138338, [[-7.93213293, 7.69707888], [-9.10730654]]),
                     (0.22285578, [[-12.38579937]])]),
}
_ATOMIC_NUMBER = {"H": 1, "He": 2, "Li": 3, "C": 6, "N": 7, "O": 8, "Na": 11, "Mg": 12, "Al": 13, "Si": 14,
                  "Fe": 26, "Cu": 29}


class PspHgh:
    def __init__(self, Zion, rloc, cloc, rp, h, identifier=""):
        if len(rp) != len(h):
            raise ValueError("Length of rp and h do not agree.")
        if len(cloc) > 4:
            raise ValueError("length(cloc) > 4 not supported.")
        self.Zion, self.rloc = int(Zion), float(rloc)
        self.cloc = list(cloc) + [0.0] * (4 - len(cloc))
        self.lmax = len(h) - 1
        self.rp = [float(r) for r in rp]
        self.h = [np.array(x, dtype=float) for x in h]
        self.identifier = identifier

    def count_n_proj_radial(self, l):
        return self.h[l].shape[0]

    def count_n_proj(self):
        return sum((2 * l + 1) * self.h[l].shape[0] for l in range(self.lmax + 1))

    def eval_psp_local_fourier(self, p):
        """p: torch tensor of |G| values.  PspHgh.jl:110-124."""
        t = p * self.rloc
        t2 = t * t
        c = self.cloc
        poly = c[0] + c[1] * (3 - t2) + c[2] * (15 - 10 * t2 + t2 ** 2) + c[3] * (105 - 105 * t2 + 21 * t2 ** 2 - t2 ** 3)
        safe = torch.where(t2 == 0, torch.ones_like(t2), t2)
        val = (4 * math.pi * self.rloc ** 2 * (-self.Zion + math.sqrt(math.pi / 2) * self.rloc * t2 * poly)
               * torch.exp(-t2 / 2) / safe)
        return torch.where(p == 0, torch.zeros_like(val), val)

    def eval_psp_projector_fourier(self, i, l, p):
        """PspHgh.jl:140-164 (divided by p^l)."""
        rp = self.rp[l]
        t2 = (p * rp) ** 2
        common = 4 * math.pi ** 1.25 * math.sqrt(2.0 ** (l + 1) * rp ** 3) * torch.exp(-t2 / 2)
        key = (l, i)
        if key == (0, 1): return common
        if key == (0, 2): return common * (2 / math.sqrt(15)) * (3 - t2)
        if key == (0, 3): return common * (4 / (3 * math.sqrt(105))) * (15 - 10 * t2 + t2 ** 2)
        if key == (1, 1): return common * (rp / math.sqrt(3))
        if key == (1, 2): return common * (2 * rp / math.sqrt(105)) * (5 - t2)
        if key == (1, 3): return common * (4 * rp / (3 * math.sqrt(1155))) * (35 - 14 * t2 + t2 ** 2)
        if key == (2, 1): return common * (rp ** 2 / math.sqrt(15))
        if key == (2, 2): return common * (2 * rp ** 2 / (3 * math.sqrt(105))) * (7 - t2)
        if key == (3, 1): return common * (rp ** 3 / math.sqrt(105))
        raise NotImplementedError(f"Not implemented for l={l} and i={i}")

    def eval_psp_energy_correction(self):
        cc = [1.0, 3.0, 15.0, 105.0]
        dc = self.Zion * self.rloc ** 2 / 2 + math.sqrt(math.pi / 2) * self.rloc ** 3 * sum(a * b for a, b in zip(cc, self.cloc))
        return 4 * math.pi * dc


def parse_hgh(text, identifier=""):
    """The ABINIT/CP2K text format read by PspHgh(path), PspHgh.jl:25-93."""
    lines = text.splitlines()
    n_elec = [int(x) for x in re.match(r"^ *(([0-9]+ *)+)", lines[1]).group(1).split()]
    m = re.match(r"^ *([-.0-9]+) +([0-9]+)( +([-.0-9]+ *)+)? *", lines[2])
    rloc, nloc = float(m.group(1)), int(m.group(2))
    cloc = [float(x) for x in m.group(3).split()] if m.group(3) else []
    if len(cloc) != nloc:
        raise ValueError("inconsistent local coefficients")
    lmax = int(re.match(r"^ *([0-9]+)", lines[3]).group(1)) - 1
    cur, rp, h = 4, [], []
    for _ in range(lmax + 1):
        m = re.match(r"^ *([-.0-9]+) +([0-9]+)( +([-.0-9]+ *)+)? *", lines[cur])
        rp.append(float(m.group(1)))
        nproj = int(m.group(2))
        hm = np.zeros((nproj, nproj))
        if nproj == 0:
            h.append(hm)
            cur += 1
            continue
        coeff = [float(x) for x in m.group(3).split()]
        for i in range(nproj):
            for j in range(i, nproj):
                hm[i, j] = hm[j, i] = coeff[j - i]
            cur += 1
            if cur >= len(lines):
                break
            mm = re.match(r"^ *(([-.0-9]+ *)+)", lines[cur])
            coeff = [float(x) for x in mm.group(1).split()] if mm else []
        h.append(hm)
    return PspHgh(sum(n_elec), rloc, cloc, rp, h, identifier)


def load_psp(symbol, functional="lda"):
    Z, n_elec, rloc, cloc, proj = _TABLE[(symbol, functional)]
    rp, h = [], []
    for r, rows in proj:
        n = len(rows)
        hm = np.zeros((n, n))
        for i, row in enumerate(rows):
            for j, v in enumerate(row):
                hm[i, i + j] = hm[i + j, i] = v
        rp.append(r)
        h.append(hm)
    return PspHgh(sum(n_elec), rloc, cloc, rp, h, identifier=f"hgh/{functional}/{symbol.lower()}-q{sum(n_elec)}")


class ElementPsp:
    """src/elements.jl ElementPsp: species + pseudopotential."""

    def __init__(self, symbol, psp=None, functional="lda"):
        self.symbol = symbol
        self.Z = _ATOMIC_NUMBER[symbol]
        self.psp = psp if psp is not None else load_psp(symbol, functional)

    def charge_ionic(self):
        return self.psp.Zion

    def n_elec_valence(self):
        return self.psp.Zion

    def n_elec_core(self):
        return self.Z - self.psp.Zion

    def __eq__(self, o):
        return isinstance(o, ElementPsp) and o.symbol == self.symbol and o.psp.identifier == self.psp.identifier

    def __hash__(self):
        return hash((self.symbol, self.psp.identifier))


def atom_decay_length(n_elec_core, n_elec_valence):
    """density_methods.jl:286-323 (ABINIT table)."""
    nv = int(round(n_elec_valence))
    if nv == 0:
        return 0.0
    tables = [(0.5, [0.6, 0.4, 0.3, 0.25, 0.2]),
              (2.5, [1.8, 1.4, 1.0, 0.7, 0.6, 0.5, 0.4, 0.35, 0.3]),
              (10.5, [2.0, 1.6, 1.25, 1.1, 1.0, 0.9, 0.8, 0.7, 0.7, 0.7, 0.6]),
              (12.5, [1.9, 1.5, 1.15, 1.0, 0.9, 0.8, 0.7, 0.6, 0.6, 0.6, 0.5]),
              (18.5, [2.0, 1.8, 1.5, 1.2, 1.0, 0.9, 0.85, 0.8, 0.75, 0.7, 0.65, 0.65, 0.6]),
              (28.5, [1.5, 1.25, 1.15, 1.05, 1.00, 0.95, 0.95, 0.9, 0.9, 0.85, 0.85, 0.80, 0.8, 0.75, 0.7]),
              (36.5, [2.0, 2.00, 1.60, 1.40, 1.25, 1.10, 1.00, 0.95, 0.90, 0.85, 0.80, 0.75, 0.7])]
    data = [2.0, 2.00, 1.55, 1.25, 1.15, 1.10, 1.05, 1.0, 0.95, 0.9, 0.85, 0.85, 0.8]
    for lim, d in tables:
        if n_elec_core < lim:
            data = d
            break
    return data[min(nv, len(data)) - 1]


def solid_harmonic_real(l, m, v):
    """Real solid harmonics on a (n,3) torch tensor (src/common/spherical_harmonics.jl:31-66)."""
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    pi = math.pi
    if l == 0:
        return torch.full_like(x, math.sqrt(1 / (4 * pi)))
    if l == 1:
        return math.sqrt(3 / (4 * pi)) * {-1: y, 0: z, 1: x}[m]
    if l == 2:
        return {-2: math.sqrt(15 / (4 * pi)) * x * y, -1: math.sqrt(15 / (4 * pi)) * y * z,
                0: math.sqrt(5 / (16 * pi)) * (2 * z * z - x * x - y * y),
                1: math.sqrt(15 / (4 * pi)) * x * z, 2: math.sqrt(15 / (16 * pi)) * (x * x - y * y)}[m]
    if l == 3:
        return {-3: math.sqrt(35 / (32 * pi)) * (3 * x * x - y * y) * y,
                -2: math.sqrt(105 / (4 * pi)) * x * y * z,
                -1: math.sqrt(21 / (32 * pi)) * y * (4 * z * z - x * x - y * y),
                0: math.sqrt(7 / (16 * pi)) * z * (2 * z * z - 3 * x * x - 3 * y * y),
                1: math.sqrt(21 / (32 * pi)) * x * (4 * z * z - x * x - y * y),
                2: math.sqrt(105 / (16 * pi)) * (x * x - y * y) * z,
                3: math.sqrt(35 / (32 * pi)) * (x * x - 3 * y * y) * x}[m]
    raise IndexError((l, m))
