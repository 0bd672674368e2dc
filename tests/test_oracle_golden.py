"""Pins the CPU oracle to the reference's own known-answer tests."""
import math
import numpy as np
import pytest

from oracle.psp_hgh import PspHgh
from oracle.basis import (Element, Model, PlaneWaveBasis, compute_fft_size, reducible_kcoords,
                          index_G_vectors, Kpoint)
from oracle.terms import (Terms, energy_hamiltonian, energy_ewald, energy_psp_correction,
                          guess_density)
from oracle import scf
from silicon import LATTICE, POSITIONS, KCOORDS, KWEIGHTS


def test_psp_hgh_values():
    # reference: test/PspHgh.jl:41-84
    psp = PspHgh.from_table("Si", "lda")
    nrm = lambda v: np.array([np.linalg.norm(v)])
    for v, ref in [([0.1, 0, 0], -400.395448865164), ([0.1, 0.2, 0], -80.39317320182417),
                   ([0.1, 0.2, -0.3], -28.95951714682582), ([1.0, -2.0, 3.0], -0.275673388844235),
                   ([10.0, 0, 0], -5.1468909215285576e-5)]:
        assert psp.eval_local_fourier(nrm(v))[0] == pytest.approx(ref * 4 * math.pi, rel=1e-10)
    pn = np.sqrt([0, 0.01, 0.1, 0.3, 1, 10])
    np.testing.assert_allclose(psp.eval_projector_fourier(1, 0, pn),
                               [6.503085484692629, 6.497277328372439, 6.445236803354619,
                                6.331078654802208, 5.947214691896995, 2.661098803299718], rtol=1e-10)
    np.testing.assert_allclose(psp.eval_projector_fourier(2, 0, pn),
                               [10.074536712471094, 10.059542796942894, 9.925438587886482,
                                9.632787375976731, 8.664551612201326, 1.666783598475508], rtol=1e-10)
    np.testing.assert_allclose(psp.eval_projector_fourier(1, 1, pn) * pn,
                               [0.0, 0.3149163627204332, 0.9853983576555614,
                                1.667197861646941, 2.8039993470553535, 3.0863036233824626], rtol=1e-10)
    np.testing.assert_allclose(psp.eval_projector_fourier(3, 1, pn) * pn,
                               [0.0, 0.7482799478933317, 2.321676914155303,
                                3.8541542745249706, 6.053770711942623, 1.6078748819430986], rtol=1e-10)


def test_psp_parser_roundtrip():
    # reference format: data/psp/hgh/lda/si-q4.hgh, parser src/pseudo/PspHgh.jl:25-93
    text = ("Si GTH-PADE-q4 GTH-LDA-q4\n    2    2\n     0.44000000    1    -7.33610297\n    2\n"
            "     0.42273813    2     5.90692831    -1.26189397\n"
            "                                        3.25819622\n     0.48427842    1     2.72701346\n")
    p, q = PspHgh.parse(text), PspHgh.from_table("Si", "lda")
    assert p.Zion == q.Zion == 4 and p.rloc == q.rloc and p.lmax == 1
    np.testing.assert_array_equal(p.cloc, q.cloc)
    for a, b in zip(p.h, q.h):
        np.testing.assert_array_equal(a, b)
    assert p.rp == q.rp


def test_compute_fft_size():
    # reference: test/compute_fft_size.jl:6-12
    for E, n in [(3, 15), (4, 15), (5, 18), (15, 27), (25, 36), (30, 40)]:
        assert compute_fft_size(LATTICE, E) == (n, n, n)
    assert compute_fft_size(LATTICE, 30, supersampling=1.8) == (36, 36, 36)


def test_energy_nuclear():
    # reference: test/energy_nuclear.jl:31,48 (ABINIT)
    assert energy_ewald(LATTICE, [4, 4], POSITIONS) == pytest.approx(-8.39789357839024, abs=1e-10)
    m = Model(LATTICE, [Element("Si")] * 2, POSITIONS, symmetries=False)
    assert energy_psp_correction(m) == pytest.approx(-0.294622067023269, abs=1e-10)


def test_ewald_golden():
    # reference: test/ewald.jl:1-52 (hydrogen atom, silicon diamond with Z = 14, boron and hydrogen molecules in a 16 bohr box)
    from oracle.forces import energy_forces_ewald
    A = 5.131570667152971
    cases = [(16 * np.eye(3), [1], [[0, 0, 0]], -0.088665545, 1e-8),
             (np.array([[0, A, A], [A, 0, A], [A, A, 0]]), [14, 14], [np.ones(3) / 8, -np.ones(3) / 8], -102.8741963352893, 1e-8),
             (16 * np.eye(3), [5, 5], [[0, 0, 0], [0.14763485355139283, 0, 0]], 1.790634595, 1e-7),
             (16 * np.eye(3), [1, 1], [[0.45312500031210007, 0.5, 0.5], [0.5468749996028622, 0.5, 0.5]], 0.31316999, 1e-7)]
    for lat, charges, pos, ref, tol in cases:
        pos = [np.array(p, dtype=float) for p in pos]
        assert energy_ewald(lat, charges, pos) == pytest.approx(ref, abs=tol)
        assert energy_forces_ewald(lat, charges, pos)[0] == pytest.approx(ref, abs=tol)


def test_G_index():
    # reference: test/PlaneWaveBasis.jl:92-96 -- index of G=[-2,-3,-1] at k=[1/3,1/3,0], fft (7,9,11), Ecut 3
    m = Model(LATTICE, [Element("Si")] * 2, POSITIONS, symmetries=False)
    kpt = Kpoint(0, [1 / 3, 1 / 3, 0], m.recip_lattice, (7, 9, 11), 3)
    lin = index_G_vectors((7, 9, 11), np.array([-2, -3, -1]))
    pos = int(np.nonzero(kpt.mapping == lin)[0][0]) + 1
    assert pos == 62


def test_fft_roundtrip_and_dft_matrix():
    # reference: test/fourier_transforms.jl:1-47
    m = Model(LATTICE, [Element("Si")] * 2, POSITIONS, symmetries=False)
    b = PlaneWaveBasis(m, 4, fft_size=(8, 9, 10), kcoords=[[0.1, 0.2, 0.3]], kweights=[1.0])
    rng = np.random.default_rng(0)
    f = rng.standard_normal(b.N) + 1j * rng.standard_normal(b.N)
    np.testing.assert_allclose(b.fft_cube(b.ifft_cube(f)), f, atol=1e-12)
    kpt = b.kpoints[0]
    c = rng.standard_normal(kpt.n_G) + 1j * rng.standard_normal(kpt.n_G)
    np.testing.assert_allclose(b.fft_kpt(kpt, b.ifft_kpt(kpt, c)), c, atol=1e-12)
    # explicit DFT: f(r) = sum_G c_G e^{i2π G·r}/sqrt(Ω)
    nx, ny, nz = b.fft_size
    r = np.stack(np.meshgrid(np.arange(nz) / nz, np.arange(ny) / ny, np.arange(nx) / nx, indexing="ij"), -1)
    r = r.reshape(-1, 3)[:, ::-1]
    ph = np.exp(2j * math.pi * (r @ kpt.G_vectors.T))
    np.testing.assert_allclose(b.ifft_kpt(kpt, c), ph @ c / math.sqrt(m.unit_cell_volume), atol=1e-12)


def test_energies_guess_density():
    # reference: test/energies_guess_density.jl:8-36 -- every energy term to 5e-8
    m = Model(LATTICE, [Element("Si")] * 2, POSITIONS, functionals=("lda_x", "lda_c_vwn"), symmetries=False)
    kc = reducible_kcoords((1, 2, 3), (0, 0.5, 0))
    b = PlaneWaveBasis(m, 15, fft_size=(27, 27, 27), kcoords=kc, kweights=[1 / 6] * 6)
    terms = Terms(b)
    rho0 = guess_density(b)
    E, H = energy_hamiltonian(b, terms, None, None, rho0)
    assert E["Hartree"] == pytest.approx(0.3527293727197568, abs=5e-8)
    assert E["Xc"] == pytest.approx(-2.3033165870558165, abs=5e-8)
    res = scf.diagonalize_all_kblocks(H, 8, tol=1e-9)
    # (the reference does not assert convergence of all 8 bands either; the energies need the lowest 4)
    assert all(np.max(r[:4]) < 1e-8 for r in res["residual_norms"])
    occ = [np.array([2., 2, 2, 2, 0, 0, 0, 0]) for _ in b.kpoints]
    rho = scf.compute_density(b, res["X"], occ)
    E, _ = energy_hamiltonian(b, terms, res["X"], occ, rho)
    ref = dict(Kinetic=3.3824289861522194, AtomicLocal=-2.4178712046759157,
               AtomicNonlocal=1.664289455206788, Hartree=0.6712993199211524,
               Xc=-2.4489960475309056, Ewald=-8.397893578467201, PspCorrection=-0.294622067031369)
    for k, v in ref.items():
        assert E[k] == pytest.approx(v, abs=5e-8), k


def test_lobpcg_free_electron():
    # reference: test/lobpcg.jl:1-48
    m = Model(LATTICE, [Element("Si")] * 2, POSITIONS, terms=("Kinetic",), symmetries=False)
    b = PlaneWaveBasis(m, 5, fft_size=(15, 15, 15), kcoords=KCOORDS, kweights=KWEIGHTS)
    _, H = energy_hamiltonian(b, Terms(b), None, None, np.zeros((1, b.N)))
    ref = [[0.00000000000, 0.56219939834, 0.56219939834, 0.56219939834, 0.56219939834,
            0.56219939834, 0.56219939834, 0.56219939834, 0.56219939834, 0.74959919778],
           [0.06246659981, 0.24986639926, 0.49973279852, 0.49973279852, 0.49973279852,
            0.56219939834, 0.56219939834, 0.56219939834, 0.74959919778, 0.74959919778],
           [0.08328879975, 0.33315519901, 0.39562179883, 0.39562179883, 0.39562179883,
            0.39562179883, 0.83288799753, 0.83288799754, 0.83288799754, 0.83288799754],
           [0.16657759951, 0.22904419932, 0.22904419932, 0.41644399877, 0.41644399877,
            0.66631039803, 0.72877699784, 0.72877699784, 0.72877699784, 0.72877699784]]
    res = scf.diagonalize_all_kblocks(H, 10, tol=1e-8)
    assert res["converged"]
    for ik in range(4):
        np.testing.assert_allclose(res["λ"][ik], ref[ik], atol=1e-8)
        assert res["n_iter"][ik] < 50
    res = scf.diagonalize_all_kblocks(H, 10, tol=1e-4, prec=False, maxiter=200)
    for ik in range(4):
        np.testing.assert_allclose(res["λ"][ik], ref[ik], atol=1e-4)


@pytest.mark.slow
def test_lobpcg_kinetic_local():
    # reference: test/lobpcg.jl:50-75 (atol 5e-7)
    m = Model(LATTICE, [Element("Si")] * 2, POSITIONS, terms=("Kinetic", "AtomicLocal"), symmetries=False)
    b = PlaneWaveBasis(m, 25, fft_size=(33, 33, 33), kcoords=KCOORDS, kweights=KWEIGHTS)
    _, H = energy_hamiltonian(b, Terms(b), None, None, np.zeros((1, b.N)))
    res = scf.diagonalize_all_kblocks(H, 6, tol=1e-8)
    ref = [[-4.087198659513310, -4.085326314828677, -0.506869382308294, -0.506869382280876, -0.506869381798614],
           [-4.085824585443292, -4.085418874576503, -0.509716820984169, -0.509716820267449, -0.508545832298541],
           [-4.086645155119840, -4.085209948598607, -0.514320642233337, -0.514320641863231, -0.499373272772206],
           [-4.085991608422304, -4.085039856878318, -0.517299903754010, -0.513805498246478, -0.497036479690380]]
    for ik in range(4):
        np.testing.assert_allclose(res["λ"][ik][:5], ref[ik], atol=5e-7)


def test_silicon_lda_scf_vs_abinit():
    # reference: test/silicon_lda.jl:10-20,47-51 (Ecut 25, fft 33, eigenvalues 1e-5, Etot 1e-5)
    m = Model(LATTICE, [Element("Si")] * 2, POSITIONS, functionals=("lda_x", "lda_c_vwn"))
    assert len(m.symmetries) == 48
    b = PlaneWaveBasis(m, 25, fft_size=(33, 33, 33), kcoords=KCOORDS, kweights=KWEIGHTS)
    ref = [[-0.178566465714968, 0.261882541175914, 0.261882541178847, 0.261882541181782,
            0.354070367072414, 0.354070367076363, 0.354070367080310, 0.376871160884678],
           [-0.127794342370963, 0.064395861472044, 0.224958824747686, 0.224958824750934,
            0.321313617512188, 0.388442495007398, 0.388442495010722, 0.542078732298094],
           [-0.108449612789883, 0.077125812982728, 0.172380374761464, 0.172380374766260,
            0.283802499666810, 0.329872296009131, 0.525606867582028, 0.525606867585921],
           [-0.058089253154566, 0.012364292440522, 0.097350168867990, 0.183765652148129,
            0.314593174568090, 0.470869435132365, 0.496966579772700, 0.517009645871194]]

    def conv(info):
        h = info["history_Etot"]
        return len(h) > 1 and abs(h[-1] - h[-2]) < 1e-7
    res = scf.self_consistent_field(b, nbandsalg=scf.AdaptiveBands(m, n_bands_converge=8), is_converged=conv)
    assert res["energies"]["total"] == pytest.approx(-7.911817522631488, abs=1e-5)
    for ik in range(4):
        np.testing.assert_allclose(res["eigenvalues"][ik][:8], ref[ik], atol=1e-5)


def test_silicon_pbe_scf_vs_abinit():
    # reference: test/silicon_pbe.jl:6-41,57-61 ("Silicon PBE (large, Float64)": Ecut 25, fft 33, ABINIT eigenvalues and
    # E_tot, test_tol 1e-5) -- the absolute pin of gga_x_pbe + gga_c_pbe and of the GGA potential term
    m = Model(LATTICE, [Element("Si", functional="pbe")] * 2, POSITIONS, functionals=("gga_x_pbe", "gga_c_pbe"))
    b = PlaneWaveBasis(m, 25, fft_size=(33, 33, 33), kcoords=KCOORDS, kweights=KWEIGHTS)
    ref = [[-0.181210259413818, 0.258840553222639, 0.258840553225549, 0.258840553228459, 0.351692348652324,
            0.351692348656259, 0.351692348660193, 0.380606400669216, 0.540705881744348, 0.540705883460555],
           [-0.130553299114991, 0.062256443775155, 0.221871391287580, 0.221871391290802, 0.322398722411882,
            0.386194327436667, 0.386194327439986, 0.546859898649217, 0.550571701390781, 0.550571701394327],
           [-0.111170738096744, 0.074494899973125, 0.169461730083372, 0.169461730088140, 0.284305392082236,
            0.330468937070505, 0.524509288492752, 0.524509288496625, 0.616964090764029, 0.619623658242765],
           [-0.061054203629684, 0.009700769243041, 0.095769985640881, 0.180784778430457, 0.315000287382235,
            0.471042322838057, 0.495281775946584, 0.517469860611792, 0.530124341745161, 0.539044739392045]]

    def conv(info):
        h = info["history_Etot"]
        return len(h) > 1 and abs(h[-1] - h[-2]) < 1e-8
    res = scf.self_consistent_field(b, nbandsalg=scf.AdaptiveBands(m, n_bands_converge=10), is_converged=conv)
    assert res["energies"]["total"] == pytest.approx(-7.854477356672080, abs=1e-5)       # observed: 2e-8
    for ik in range(4):
        np.testing.assert_allclose(res["eigenvalues"][ik][:10], ref[ik], atol=1e-5)       # observed: 7e-7


def test_iron_pbe_collinear_scf_vs_abinit():
    # reference: test/iron_pbe.jl:6-70 (bcc Fe, GTH-PADE-q8, PBE, collinear spin, Fermi-Dirac T = 0.01, Ecut 20,
    # fft 20, shifted 4x4x4 k-grid; ABINIT eigenvalues 5e-6, E_tot, magnetisation 5e-5) -- pins the spin-polarised
    # GGA path, smearing / Fermi level, Kerker mixing and the d-channel projectors
    from oracle.terms import guess_density as guess
    ref = [[0.0603597727989307, 0.1964963273638626, 0.196496327424440, 0.279192222553112, 0.2791922225741613,
            0.3415221335998876, 0.837882559419754, 0.883850560591423, 0.8838505606211768, 1.3135367355436536],
           [0.1384929268069029, 0.1847168453364975, 0.223179759800174, 0.320070899985990, 0.3500724891746176,
            0.4685757607370267, 0.541752194212558, 0.751365680734661, 0.8039132927796911, 1.3939297677405071],
           [-0.017996603976028, 0.2383855826934185, 0.238385582734711, 0.248204676138927, 0.2509395500598295,
            0.2776437400588896, 1.069915401940919, 1.088217176897224, 1.094997859335961, 1.0949978593466851],
           [0.1102557166995405, 0.2077201723056727, 0.220685303120809, 0.289884460857327, 0.3490062808992303,
            0.3571047250832524, 0.664551132243957, 0.890354172420178, 0.939822681382406, 1.2259972985258636],
           [0.1723514110126840, 0.1723514110181127, 0.189598224957126, 0.315084007273243, 0.3150840073174671,
            0.5487559496577702, 0.548755949657792, 0.571153866844390, 1.0611134432316718, 1.1887518709297569],
           [0.1360541296075938, 0.1413608406233668, 0.337616953214017, 0.337616953257584, 0.3463728840905585,
            0.4304010493995122, 0.688627292839765, 0.688627292852315, 0.885008380770321, 0.9722786718518246],
           [0.0802990962833626, 0.3488798033726516, 0.348879803416372, 0.533263624117060, 0.560354114948579,
            0.5603541149670136, 0.923281827089562, 0.967838872125574, 0.9678388721641925, 1.300215418446228],
           [0.2341496631160049, 0.2737567834221212, 0.320646675118266, 0.590600827614029, 0.6440928824646408,
            0.6458637753212415, 0.678343515679297, 0.838647690182280, 0.8763210347583158, 1.4092936521531203],
           [-0.002234753604747, 0.4096246186291687, 0.409624618662776, 0.434260327970128, 0.5068101375084778,
            0.5757957165012942, 1.137207834311533, 1.137826252874365, 1.170363096833071, 1.170363096849632],
           [0.1518900787487526, 0.3293780680641614, 0.376401550325491, 0.512562269331525, 0.5557310122303195,
            0.6261449425921871, 0.794097184155989, 0.967295197092196, 1.0000550921659532, 1.2999173820510477],
           [0.2873355363445261, 0.2873355363447599, 0.319313192152575, 0.537629072823137, 0.5376290728591641,
            0.6802062250711767, 0.704199805731151, 0.704199805731498, 1.1322730987840155, 1.255912074880981],
           [0.2512356397409882, 0.315293666807424, 0.491297439253523, 0.4912974392811193, 0.5558649368408816,
            0.556692128645629, 0.777563890322163, 0.7775638903489546, 0.9998569230219644, 1.1313796020728688]]
    lat = 2.71176 * np.array([[-1, 1, 1], [1, -1, 1], [1, 1, -1]], dtype=float)
    fe = Element("Fe", PspHgh.from_table("Fe", "lda-q8"))
    m = Model(lat, [fe], [np.zeros(3)], functionals=("gga_x_pbe", "gga_c_pbe"), temperature=0.01, magnetic_moments=[4.0])
    b = PlaneWaveBasis(m, 20, kgrid=(4, 4, 4), kshift=(0.5, 0.5, 0.5), fft_size=(20, 20, 20))
    assert len(b.kpoints) == 12 and m.n_electrons == 8        # 6 irreducible k-points x 2 spins

    def conv(info):
        h = info["history_Etot"]
        return len(h) > 1 and abs(h[-1] - h[-2]) < 1e-10
    res = scf.self_consistent_field(b, rho=guess(b, [4.0]), mixing="kerker",
                                    nbandsalg=scf.AdaptiveBands(m, n_bands_converge=10), is_converged=conv)
    assert res["energies"]["total"] == pytest.approx(-18.21465922614397, abs=5e-6)       # observed: 1.1e-7
    mag = float((res["rho"][0] - res["rho"][1]).sum() * b.dvol)
    assert mag == pytest.approx(2.98199463, abs=5e-5)                                     # observed: 2.2e-5
    # the irreducible k-points come from an orbit search, not spglib: match blocks to ABINIT's by their spectra
    used = set()
    for ik, kpt in enumerate(b.kpoints):
        d = [np.abs(np.array(res["eigenvalues"][ik][:10]) - np.array(r)).max() for r in ref]
        j = int(np.argmin(d))
        assert d[j] < 5e-6 and (j < 6) == (kpt.spin == 0)                                # observed: 2.9e-6
        used.add(j)
    assert used == set(range(12))


def test_oracle_matches_committed_fixture():
    # the oracle is frozen by tests/golden/si_block_fixture.npz (made by tests/golden/make_fixtures.py)
    import os
    from oracle.terms import HamiltonianBlock
    f = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "si_block_fixture.npz"))
    m = Model(LATTICE, [Element("Si")] * 2, POSITIONS, functionals=("lda_x", "lda_c_vwn"), symmetries=False)
    b = PlaneWaveBasis(m, 10, fft_size=tuple(int(x) for x in f["fft_size"]), kcoords=[[0.1, -0.2, 0.3]], kweights=[1.0])
    np.testing.assert_array_equal(b.kpoints[0].mapping, f["mapping"])
    _, ham = energy_hamiltonian(b, Terms(b), None, None, guess_density(b))
    np.testing.assert_allclose(ham[0].matmul(f["psi"].T).T, f["hpsi"], atol=1e-12 * np.abs(f["hpsi"]).max())
