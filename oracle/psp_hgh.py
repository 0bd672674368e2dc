"""HGH/GTH pseudopotentials (oracle; test infrastructure only).

Follows src/pseudo/PspHgh.jl:25-184 of the reference.  Parameter tables are the published
GTH-LDA / GTH-PBE values (Goedecker-Teter-Hutter 1996, Hartwigsen-Goedecker-Hutter 1998,
Krack 2005), identical to data/psp/hgh/{lda,pbe}/*.hgh of the reference.
"""
import math
import re
import numpy as np

# (n_elec per channel), rloc, cloc, [(rp, h-upper-triangle rows)]
PSP_TABLE = {
    ("Si", "lda"): dict(Z=14, n_elec=[2, 2], rloc=0.44, cloc=[-7.33610297],
                        proj=[(0.42273813, [[5.90692831, -1.26189397], [3.25819622]]),
                              (0.48427842, [[2.72701346]])]),
    ("Si", "pbe"): dict(Z=14, n_elec=[2, 2], rloc=0.44, cloc=[-6.26928833],
                        proj=[(0.43563383, [[8.95174150, -2.70627082], [3.49378060]]),
                              (0.49794218, [[2.43127673]])]),
    # Fe GTH-PADE-q8 (large core; the pseudopotential of the reference's iron tests, test/testcases.jl:129)
    ("Fe", "lda-q8"): dict(Z=26, n_elec=[2, 0, 6], rloc=0.61, cloc=[],
                           proj=[(0.45448200, [[3.01664046, -1.00040646, 0.79478164], [2.58303836, -2.05211737],
                                               [3.25763534]]),
                                 (0.63890282, [[1.49964199, -0.13812935], [0.32687369]]),
                                 (0.30873177, [[-9.14535371]])]),
    ("Al", "lda"): dict(Z=13, n_elec=[2, 1], rloc=0.45, cloc=[-8.49135116],
                        proj=[(0.46010427, [[5.08833953, -1.03784325], [2.67969975]]),
                              (0.53674439, [[2.19343827]])]),
    ("Al", "pbe"): dict(Z=13, n_elec=[2, 1], rloc=0.45, cloc=[-7.55476126],
                        proj=[(0.48743529, [[6.95993832, -1.88883584], [2.43847659]]),
                              (0.56218949, [[1.86529857]])]),
    ("Fe", "pbe"): dict(Z=26, n_elec=[4, 6, 6], rloc=0.36, cloc=[6.75678916, -0.22883251],
                        proj=[(0.27826303, [[0.62950570, 7.91313242], [-10.21581002]]),
                              (0.25138338, [[-7.93213293, 7.69707888], [-9.10730654]]),
                              (0.22285578, [[-12.38579937]])]),
}


class PspHgh:
    """struct PspHgh, src/pseudo/PspHgh.jl:4-13,95-107."""

    def __init__(self, Zion, rloc, cloc, rp, h, Z=None, description=""):
        assert len(rp) == len(h)
        assert len(cloc) <= 4
        self.Zion = int(Zion)
        self.rloc = float(rloc)
        self.cloc = np.zeros(4)
        self.cloc[:len(cloc)] = cloc
        self.lmax = len(h) - 1
        self.rp = [float(r) for r in rp]
        self.h = [np.array(hl, dtype=float).reshape(len(hl), len(hl)) if len(hl) else
                  np.zeros((0, 0)) for hl in h]
        self.Z = Z
        self.description = description

    # -- counting helpers (src/pseudo/NormConservingPsp.jl count_n_proj*) --
    def n_proj_radial(self, l):
        return self.h[l].shape[0]

    def n_proj(self):
        return sum((2 * l + 1) * self.n_proj_radial(l) for l in range(self.lmax + 1))

    @staticmethod
    def from_table(symbol, functional="lda"):
        e = PSP_TABLE[(symbol, functional)]
        rp, h = [], []
        for (r, rows) in e["proj"]:
            n = len(rows)
            hm = np.zeros((n, n))
            for i, row in enumerate(rows):
                for j, v in enumerate(row):
                    hm[i, i + j] = hm[i + j, i] = v
            rp.append(r)
            h.append(hm)
        return PspHgh(sum(e["n_elec"]), e["rloc"], e["cloc"], rp, h, Z=e["Z"],
                      description=f"{symbol} GTH-{functional.upper()}")

    @staticmethod
    def parse(text):
        """Parser for the ABINIT/CP2K .hgh text format, PspHgh.jl:25-93."""
        lines = text.splitlines()
        description = lines[0]
        n_elec = [int(p) for p in re.match(r"^ *(([0-9]+ *)+)", lines[1]).group(1).split()]
        m = re.match(r"^ *([-.0-9]+) +([0-9]+)( +([-.0-9]+ *)+)? *", lines[2])
        rloc, nloc = float(m.group(1)), int(m.group(2))
        cloc = [float(p) for p in m.group(3).split()] if m.group(3) else []
        assert len(cloc) == nloc
        lmax = int(re.match(r"^ *([0-9]+)", lines[3]).group(1)) - 1
        rp, h = [], []
        cur = 4
        for _l in range(lmax + 1):
            m = re.match(r"^ *([-.0-9]+) +([0-9]+)( +([-.0-9]+ *)+)? *", lines[cur])
            rp.append(float(m.group(1)))
            nproj = int(m.group(2))
            hm = np.zeros((nproj, nproj))
            if nproj == 0:
                h.append(hm)
                cur += 1
                continue
            hcoeff = [float(p) for p in m.group(3).split()]
            for i in range(nproj):
                for j in range(i, nproj):
                    hm[j, i] = hm[i, j] = hcoeff[j - i]
                cur += 1
                if cur >= len(lines):
                    break
                mm = re.match(r"^ *(([-.0-9]+ *)+)", lines[cur])
                hcoeff = [float(p) for p in mm.group(1).split()] if mm else []
            h.append(hm)
        return PspHgh(sum(n_elec), rloc, cloc, rp, h, description=description)

    # -- PspHgh.jl:110-124 --
    def eval_local_fourier(self, p):
        p = np.asarray(p, dtype=float)
        out = np.zeros_like(p)
        nz = p != 0
        t = p[nz] * self.rloc
        c = self.cloc
        P = (c[0] + c[1] * (3 - t**2) + c[2] * (15 - 10 * t**2 + t**4)
             + c[3] * (105 - 105 * t**2 + 21 * t**4 - t**6))
        out[nz] = (4 * math.pi * self.rloc**2
                   * (-self.Zion + math.sqrt(math.pi / 2) * self.rloc * t**2 * P)
                   * np.exp(-t**2 / 2) / t**2)
        return out

    # -- PspHgh.jl:140-164 (already divided by p^l) --
    def eval_projector_fourier(self, i, l, p):
        p = np.asarray(p, dtype=float)
        rp = self.rp[l]
        t = p * rp
        common = 4 * math.pi**1.25 * math.sqrt(2.0**(l + 1) * rp**3) * np.exp(-t**2 / 2)
        if l == 0 and i == 1:
            return common
        if l == 0 and i == 2:
            return common * 2 / math.sqrt(15) * (3 - t**2)
        if l == 0 and i == 3:
            return common * 4 / (3 * math.sqrt(105)) * (15 - 10 * t**2 + t**4)
        if l == 1 and i == 1:
            return common * 1 / math.sqrt(3) * rp
        if l == 1 and i == 2:
            return common * 2 / math.sqrt(105) * rp * (5 - t**2)
        if l == 1 and i == 3:
            return common * 4 / (3 * math.sqrt(1155)) * rp * (35 - 14 * t**2 + t**4)
        if l == 2 and i == 1:
            return common * 1 / math.sqrt(15) * rp**2
        if l == 2 and i == 2:
            return common * 2 / (3 * math.sqrt(105)) * rp**2 * (7 - t**2)
        if l == 3 and i == 1:
            return common * 1 / math.sqrt(105) * rp**3
        raise NotImplementedError((l, i))

    # -- PspHgh.jl:173-184 --
    def energy_correction(self):
        cc = np.array([1.0, 3.0, 15.0, 105.0])
        dc = (self.Zion * self.rloc**2 / 2
              + math.sqrt(math.pi / 2) * self.rloc**3 * float(np.sum(cc * self.cloc)))
        return 4 * math.pi * dc


def solid_harmonic_real(l, m, v):
    """Real solid harmonics R_lm = r^l Y_lm, src/common/spherical_harmonics.jl:31-66.
    v: (..., 3) array."""
    x, y, z = v[..., 0], v[..., 1], v[..., 2]
    pi = math.pi
    if l == 0:
        return np.full(x.shape, math.sqrt(1 / (4 * pi)))
    if l == 1:
        c = math.sqrt(3 / (4 * pi))
        return {-1: c * y, 0: c * z, 1: c * x}[m]
    if l == 2:
        if m == -2: return math.sqrt(15 / (4 * pi)) * x * y
        if m == -1: return math.sqrt(15 / (4 * pi)) * y * z
        if m == 0: return math.sqrt(5 / (16 * pi)) * (2 * z**2 - x**2 - y**2)
        if m == 1: return math.sqrt(15 / (4 * pi)) * x * z
        if m == 2: return math.sqrt(15 / (16 * pi)) * (x**2 - y**2)
    if l == 3:
        if m == -3: return math.sqrt(35 / (32 * pi)) * (3 * x**2 - y**2) * y
        if m == -2: return math.sqrt(105 / (4 * pi)) * x * y * z
        if m == -1: return math.sqrt(21 / (32 * pi)) * y * (4 * z**2 - x**2 - y**2)
        if m == 0: return math.sqrt(7 / (16 * pi)) * z * (2 * z**2 - 3 * x**2 - 3 * y**2)
        if m == 1: return math.sqrt(21 / (32 * pi)) * x * (4 * z**2 - x**2 - y**2)
        if m == 2: return math.sqrt(105 / (16 * pi)) * (x**2 - y**2) * z
        if m == 3: return math.sqrt(35 / (32 * pi)) * (x**2 - 3 * y**2) * x
    raise IndexError((l, m))


def atom_decay_length(n_elec_core, n_elec_valence):
    """src/density_methods.jl:286-323 (ABINIT table)."""
    nv = int(round(n_elec_valence))
    if nv == 0:
        return 0.0
    if n_elec_core < 0.5:
        data = [0.6, 0.4, 0.3, 0.25, 0.2]
    elif n_elec_core < 2.5:
        data = [1.8, 1.4, 1.0, 0.7, 0.6, 0.5, 0.4, 0.35, 0.3]
    elif n_elec_core < 10.5:
        data = [2.0, 1.6, 1.25, 1.1, 1.0, 0.9, 0.8, 0.7, 0.7, 0.7, 0.6]
    elif n_elec_core < 12.5:
        data = [1.9, 1.5, 1.15, 1.0, 0.9, 0.8, 0.7, 0.6, 0.6, 0.6, 0.5]
    elif n_elec_core < 18.5:
        data = [2.0, 1.8, 1.5, 1.2, 1.0, 0.9, 0.85, 0.8, 0.75, 0.7, 0.65, 0.65, 0.6]
    elif n_elec_core < 28.5:
        data = [1.5, 1.25, 1.15, 1.05, 1.00, 0.95, 0.95, 0.9, 0.9, 0.85, 0.85, 0.80,
                0.8, 0.75, 0.7]
    elif n_elec_core < 36.5:
        data = [2.0, 2.00, 1.60, 1.40, 1.25, 1.10, 1.00, 0.95, 0.90, 0.85, 0.80, 0.75, 0.7]
    else:
        data = [2.0, 2.00, 1.55, 1.25, 1.15, 1.10, 1.05, 1.0, 0.95, 0.9, 0.85, 0.85, 0.8]
    return data[min(nv, len(data)) - 1]
