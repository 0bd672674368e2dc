"""CPU tests of the product's host-side logic (no kernels): models, symmetries, k-grids, sharding,
band/tolerance heuristics, Anderson, occupations, and the world_size-2 gloo path of the k-point comm."""
import os
import math
import numpy as np
import pytest
import torch

from silicon import LATTICE, POSITIONS


def test_model_symmetries_and_kgrid():
    import dftk_b200 as dftk
    from dftk_b200.basis import irreducible_kcoords
    Si = dftk.ElementPsp("Si")
    m = dftk.model_DFT(LATTICE, [Si, Si], POSITIONS, functionals=dftk.LDA())
    assert len(m.symmetries) == 48 and m.symmetries[0].isone()
    assert m.n_electrons == 8 and m.filled_occupation == 2 and m.n_spin_components == 1
    assert m.term_types == ["Kinetic", "AtomicLocal", "AtomicNonlocal", "Ewald", "PspCorrection", "Hartree", "Xc"]
    # silicon 3x3x3: 4 irreducible points with weights 1,8,6,12 /27 (test/testcases.jl:24-28)
    k, w = irreducible_kcoords(dftk.MonkhorstPack((3, 3, 3)), m.symmetries)
    assert sorted(round(x * 27) for x in w) == [1, 6, 8, 12]
    k, w = irreducible_kcoords(dftk.MonkhorstPack((8, 8, 8)), m.symmetries)
    assert len(k) == 29 and abs(sum(w) - 1) < 1e-14
    m2 = dftk.model_DFT(LATTICE, [Si, Si], POSITIONS, functionals=dftk.LDA(), temperature=0.01)
    assert m2.term_types[-1] == "Entropy" and m2.smearing == "FermiDirac"
    with pytest.raises(ValueError):
        dftk.Model(LATTICE, [Si], POSITIONS)


def test_fft_size_and_reference_kgrid_order():
    import dftk_b200 as dftk
    m = dftk.Model(LATTICE)
    for E, n in [(3, 15), (5, 18), (15, 27), (25, 36), (30, 40)]:       # test/compute_fft_size.jl:6-12
        assert dftk.compute_fft_size(m, E) == (n, n, n)
    ks = dftk.MonkhorstPack((1, 2, 3), kshift=(0, 0.5, 0)).reducible_kcoords()
    assert len(ks) == 6 and np.allclose(ks[0], [0, 0.25, -1 / 3])


def test_psp_parser_and_values():
    import dftk_b200 as dftk
    text = ("Si GTH-PADE-q4 GTH-LDA-q4\n    2    2\n     0.44000000    1    -7.33610297\n    2\n"
            "     0.42273813    2     5.90692831    -1.26189397\n"
            "                                        3.25819622\n     0.48427842    1     2.72701346\n")
    p, q = dftk.parse_hgh(text), dftk.load_psp("Si", "lda")
    assert p.Zion == q.Zion == 4 and p.rp == q.rp and all(np.array_equal(a, b) for a, b in zip(p.h, q.h))
    # test/PspHgh.jl:41-84
    v = q.eval_psp_local_fourier(torch.tensor([math.sqrt(0.05), 10.0], dtype=torch.float64)) / (4 * math.pi)
    np.testing.assert_allclose(v.numpy(), [-80.39317320182417, -5.1468909215285576e-5], rtol=1e-10)
    pn = torch.tensor(np.sqrt([0, 0.01, 0.1, 0.3, 1, 10]))
    np.testing.assert_allclose(q.eval_psp_projector_fourier(2, 0, pn).numpy(),
                               [10.074536712471094, 10.059542796942894, 9.925438587886482,
                                9.632787375976731, 8.664551612201326, 1.666783598475508], rtol=1e-10)
    assert q.count_n_proj() == 5


def test_split_evenly_and_padding():
    from dftk_b200.parallel import split_evenly, pad_kpoints_for_ranks
    assert split_evenly(range(10), 3) == [[0, 1, 2, 3], [4, 5, 6], [7, 8, 9]]      # test/split_evenly.jl
    assert split_evenly(range(4), 4) == [[0], [1], [2], [3]]
    k, w = pad_kpoints_for_ranks([[0, 0, 0], [0.5, 0, 0]], [0.25, 0.75], 4)
    assert len(k) == 4 and abs(sum(w) - 1) < 1e-15 and max(w) <= 0.375 + 1e-15


def test_adaptive_bands_and_diagtol():
    import dftk_b200 as dftk
    Si = dftk.ElementPsp("Si")
    m = dftk.model_DFT(LATTICE, [Si, Si], POSITIONS, functionals=dftk.LDA())
    ab = dftk.AdaptiveBands(m)
    assert (ab.n_bands_converge, ab.n_bands_compute) == (4, 7)                  # 7 / 4
    assert ab.determine_n_bands(None, None, None) == (5, 7)                     # first step converges 5
    occ = [np.array([2, 2, 2, 2, 0, 0, 0.0])]
    eig = [np.array([-0.2, 0.1, 0.1, 0.1, 0.3, 0.3, 0.4])]
    assert ab.determine_n_bands(occ, eig, None) == (4, 7)
    dt = dftk.AdaptiveDiagtol()
    assert dt.determine_diagtol(dict(n_iter=1, history_drho=[])) == 0.025
    assert dt.determine_diagtol(dict(n_iter=3, history_drho=[0.2, 0.01])) == pytest.approx(0.002)
    assert dt.determine_diagtol(dict(n_iter=9, history_drho=[1e-20])) == 100 * np.finfo(float).eps


def test_anderson_converges_linear_fixed_point():
    # reference: test/anderson.jl -- Anderson solves a linear fixed-point problem in <= dim+1 steps
    import dftk_b200 as dftk
    torch.manual_seed(0)
    n = 6
    A = torch.randn(n, n, dtype=torch.float64) * 0.3
    b = torch.randn(n, dtype=torch.float64)
    f = lambda x: A @ x + b
    xstar = torch.linalg.solve(torch.eye(n, dtype=torch.float64) - A, b)
    acc = dftk.AndersonAcceleration(m=10)
    x = torch.zeros(n, dtype=torch.float64)
    for _ in range(n + 2):
        x = acc(x, 0.8, f(x) - x)
    assert (x - xstar).abs().max().item() < 1e-9


class _FakeBasis:
    def __init__(self, model, weights, comm, layout=None):
        self.model, self.kweights, self.comm_kpts = model, weights, comm
        if layout is not None:
            self.layout = layout


def test_occupation_insulator_and_metal():
    import dftk_b200 as dftk
    from dftk_b200.occupation import compute_occupation
    Si = dftk.ElementPsp("Si")
    m = dftk.model_DFT(LATTICE, [Si, Si], POSITIONS, functionals=dftk.LDA())
    b = _FakeBasis(m, [0.5, 0.5], dftk.KpointComm())
    eig = [np.array([-0.2, 0.0, 0.1, 0.2, 0.5, 0.6]), np.array([-0.1, 0.0, 0.1, 0.25, 0.45, 0.7])]
    occ, eF = compute_occupation(b, eig)
    assert eF == pytest.approx((0.25 + 0.45) / 2) and all(np.array_equal(o, [2, 2, 2, 2, 0, 0]) for o in occ)
    mm = dftk.model_DFT(LATTICE, [Si, Si], POSITIONS, functionals=dftk.LDA(), temperature=0.02)
    bm = _FakeBasis(mm, [0.5, 0.5], dftk.KpointComm())
    occ, eF = compute_occupation(bm, eig)
    assert abs(sum(w * o.sum() for w, o in zip(bm.kweights, occ)) - 8) < 1e-10


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import dftk_b200 as dftk
    from dftk_b200.occupation import compute_occupation
    comm = dftk.KpointComm.from_torch_distributed(with_nccl_id=False)
    assert comm.sum(rank + 1.0) == 3.0 and comm.max(rank) == 1 and comm.min(rank) == 0
    assert comm.bcast_object("x" if rank == 0 else None) == "x"
    Si = dftk.ElementPsp("Si")
    mm = dftk.model_DFT(LATTICE, [Si, Si], POSITIONS, functionals=dftk.LDA(), temperature=0.02)
    eig = [np.array([-0.2, 0.0, 0.1, 0.2, 0.5, 0.6]), np.array([-0.1, 0.0, 0.1, 0.25, 0.45, 0.7])]
    # each rank owns one of the two k-points: the Fermi level must equal the single-process result bit for bit
    from dftk_b200.parallel import BlockLayout
    layout = BlockLayout(2, 1, [0.5, 0.5], [1.0, 1.0], 2, rank)
    assert layout.mine == [rank]
    b = _FakeBasis(mm, [0.5], comm, layout)
    occ, eF = compute_occupation(b, [eig[rank]])
    # packed collectives: a fixed-size allgather with statistics behind the eigenvalues, an array allreduce
    from dftk_b200.occupation import gather_eigenvalues
    ev, w, stats = gather_eigenvalues(b, [eig[rank]], stats=[10.0 + rank, 1.0])
    assert np.array_equal(ev[0], eig[0]) and np.array_equal(ev[1], eig[1]) and w == [0.5, 0.5]
    assert stats.tolist() == [[10.0, 1.0], [11.0, 1.0]]
    assert comm.allreduce(np.array([1.0, rank]), "sum").tolist() == [2.0, 1.0]
    assert comm.all_true(True) and not comm.all_true(rank == 0)
    q.put((rank, eF, occ[0].tolist()))
    dist.destroy_process_group()


def test_kpoint_comm_gloo_world2():
    import torch.multiprocessing as mp
    import dftk_b200 as dftk
    from dftk_b200.occupation import compute_occupation
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    Si = dftk.ElementPsp("Si")
    mm = dftk.model_DFT(LATTICE, [Si, Si], POSITIONS, functionals=dftk.LDA(), temperature=0.02)
    eig = [np.array([-0.2, 0.0, 0.1, 0.2, 0.5, 0.6]), np.array([-0.1, 0.0, 0.1, 0.25, 0.45, 0.7])]
    occ, eF = compute_occupation(_FakeBasis(mm, [0.5, 0.5], dftk.KpointComm()), eig)
    assert out[0][1] == eF and out[1][1] == eF                      # bit-identical Fermi level on every rank
    assert out[0][2] == occ[0].tolist() and out[1][2] == occ[1].tolist()


def test_block_layout_flattens_spin_and_balances():
    """(k, spin) blocks are flattened and dealt longest-first; the map is deterministic, covers every block
    once and leaves no rank empty (PlaneWaveBasis.jl:190-203 forbids empty ranks)."""
    from dftk_b200.parallel import BlockLayout, lpt_assign, pad_kpoints_for_ranks
    costs = [1150.0, 1144.0, 1161.0, 1139.0, 1150.0] * 2            # 5 k-points x 2 spins
    for n in (1, 2, 4, 8):
        lay = [BlockLayout(5, 2, [0.2] * 5, costs, n, r) for r in range(n)]
        assert sorted(b for l in lay for b in l.mine) == list(range(10))
        assert all(l.owner == lay[0].owner for l in lay)
        loads = [sum(costs[b] for b in l.mine) for l in lay]
        assert max(loads) - min(loads) <= max(costs) + 1e-9
        assert all(l.mine == sorted(l.mine) for l in lay)
    assert max(len(l.mine) for l in lay) == 2 and lay[0].max_local == 2     # 10 blocks on 8 ranks
    # spin x k fills ranks that k alone could not: 3 k-points, 2 spins on 6 ranks
    lay = BlockLayout(3, 2, [1 / 3] * 3, [1.0] * 6, 6, 4)
    assert len(lay.mine) == 1
    assert lpt_assign([3.0, 1.0, 1.0, 1.0], 2) == [0, 1, 1, 1]
    kc, kw = pad_kpoints_for_ranks([[0, 0, 0]], [1.0], 4, n_spin=2)
    assert len(kc) == 2 and kw == [0.5, 0.5]
    with pytest.raises(ValueError):
        BlockLayout(1, 1, [1.0], [1.0], 2, 0)


def test_ewald_forces_and_force_symmetrisation_match_oracle():
    """Host-side pieces of compute_forces (ewald.jl:64-168, symmetry.jl:379-423) against the oracle."""
    import dftk_b200 as dftk
    from oracle import forces as oforces
    from oracle.basis import Element, Model, PlaneWaveBasis
    pos = [POSITIONS[0] + np.array([0.01, 0.02, -0.015]), POSITIONS[1] + np.array([0.0, -0.004, 0.003])]
    e, f = dftk.energy_forces_ewald(LATTICE, [4, 4], pos)
    eo, fo = oforces.energy_forces_ewald(LATTICE, [4, 4], pos)
    assert e == pytest.approx(eo, abs=1e-12)
    np.testing.assert_allclose(np.array(f), np.array(fo), atol=1e-12)
    from dftk_b200.terms import energy_ewald
    assert e == pytest.approx(energy_ewald(LATTICE, [4, 4], pos), abs=1e-12)
    # the reference's own golden Ewald energies (test/ewald.jl:1-52) through the product's host code
    for lat_g, ch, pos_g, ref, tol in [(16 * np.eye(3), [1], [[0, 0, 0]], -0.088665545, 1e-8),
                                        (LATTICE, [14, 14], POSITIONS, -102.8741963352893, 1e-8),
                                        (16 * np.eye(3), [5, 5], [[0, 0, 0], [0.14763485355139283, 0, 0]], 1.790634595, 1e-7)]:
        pg = [np.array(q, dtype=float) for q in pos_g]
        assert energy_ewald(lat_g, ch, pg) == pytest.approx(ref, abs=tol)
        assert dftk.energy_forces_ewald(lat_g, ch, pg)[0] == pytest.approx(ref, abs=tol)
    # a 3-atom, two-species cell exercises the chunked structure-factor loops
    lat = np.diag([7.0, 8.0, 9.5])
    p3 = [np.array([0.1, 0.2, 0.3]), np.array([0.55, 0.6, 0.1]), np.array([0.3, 0.9, 0.7])]
    e, f = dftk.energy_forces_ewald(lat, [4, 3, 4], p3)
    eo, fo = oforces.energy_forces_ewald(lat, [4, 3, 4], p3)
    assert e == pytest.approx(eo, abs=1e-12)
    np.testing.assert_allclose(np.array(f), np.array(fo), atol=1e-12)
    assert np.abs(np.sum(np.array(f), axis=0)).max() < 1e-10            # translation invariance
    # symmetrisation: silicon with one atom moved along [111] keeps 12 operations
    shift = 0.003 * np.ones(3)
    Si = dftk.ElementPsp("Si")
    model = dftk.model_DFT(LATTICE, [Si, Si], [POSITIONS[0] + shift, POSITIONS[1]], functionals=dftk.LDA())
    assert len(model.symmetries) == 12
    om = Model(LATTICE, [Element("Si")] * 2, [POSITIONS[0] + shift, POSITIONS[1]])
    ob = PlaneWaveBasis(om, 5, kgrid=(1, 1, 1), fft_size=(15, 15, 15))
    rng = np.random.default_rng(0)
    F = [rng.standard_normal(3), rng.standard_normal(3)]
    got = dftk.symmetrize_forces(model, F, symmetries=model.symmetries)
    want = oforces.symmetrize_forces(ob, F, symmetries=om.symmetries)
    np.testing.assert_allclose(np.array(got), np.array(want), atol=1e-14)
    np.testing.assert_allclose(got[0], got[0][0] * np.ones(3), atol=1e-14)
    np.testing.assert_allclose(got[0], -got[1], atol=1e-14)


def test_smearing_functions_satisfy_reference_identities():
    """test/occupation.jl:17-34 ("Smearing functions"): f(-inf) = 1, f(inf) = 0 and the entropy identity s'(x) = x f'(x),
    for the product's host functions and the oracle's."""
    from dftk_b200.terms import smearing_occupation as pf, smearing_entropy as ps
    from oracle.terms import smearing_occupation as of, smearing_entropy as os_
    eps = 1e-6
    for f, s in ((pf, ps), (of, os_)):
        for kind in ("FermiDirac", "Gaussian"):
            assert float(f(kind, np.array([-np.inf]))[0]) == 1.0 and float(f(kind, np.array([np.inf]))[0]) == 0.0
            for x in (0.04, -0.7, 1.3):
                sp = (s(kind, np.array([x + eps]))[0] - s(kind, np.array([x - eps]))[0]) / (2 * eps)
                fp = (f(kind, np.array([x + eps]))[0] - f(kind, np.array([x - eps]))[0]) / (2 * eps)
                assert abs(sp - x * fp) < 1e-6, (kind, x)
    x = np.linspace(-30, 30, 41)
    for kind in ("FermiDirac", "Gaussian", "None"):
        np.testing.assert_allclose(pf(kind, x), of(kind, x), atol=1e-16)
        np.testing.assert_allclose(ps(kind, x), os_(kind, x), atol=1e-16)


def test_occupation_metal_many_blocks_matches_blockwise_sum():
    """The Fermi level is bisected over ONE flat array of all eigenvalues (the per-block Python loop was a third of a metal's
    SCF step): the result must satisfy the electron-count equation of occupation.jl:30-50 evaluated block by block."""
    import dftk_b200 as dftk
    from dftk_b200.occupation import compute_occupation
    from dftk_b200.terms import smearing_occupation
    Si = dftk.ElementPsp("Si")
    mm = dftk.model_DFT(LATTICE, [Si, Si], POSITIONS, functionals=dftk.LDA(), temperature=0.01)
    rng = np.random.default_rng(7)
    nk = 84
    w = rng.random(nk)
    w /= w.sum()
    eig = [np.sort(rng.normal(0.2, 0.3, 9)) for _ in range(nk)]
    b = _FakeBasis(mm, list(w), dftk.KpointComm())
    occ, eF = compute_occupation(b, eig)
    n_el = sum(wk * (2 * smearing_occupation("FermiDirac", (e - eF) / 0.01)).sum() for wk, e in zip(w, eig))
    assert abs(n_el - 8) < 1e-9
    assert all(np.allclose(o, 2 * smearing_occupation("FermiDirac", (e - eF) / 0.01), atol=1e-14) for o, e in zip(occ, eig))


def test_slab_mode_decisions_and_band_shares():
    """Single-k multi-GPU (comm_slab): which blocks take the slab eigensolver, and the band shares of compute_density."""
    from types import SimpleNamespace
    from dftk_b200.eigen import _use_slabs
    comm = SimpleNamespace(nranks=4, rank=1)
    A = SimpleNamespace(basis=SimpleNamespace(comm_slab=comm))
    assert _use_slabs(A, torch.empty((503, 264859), device="meta"))
    assert not _use_slabs(A, torch.empty((15, 5440), device="meta"))            # batched small-block regime
    assert not _use_slabs(A, torch.empty((111, 1200), device="meta"))           # slabs of <= 3 n_bands rows: the solver refuses
    assert not _use_slabs(SimpleNamespace(basis=SimpleNamespace(comm_slab=None)), torch.empty((503, 264859), device="meta"))
    for nb in (0, 1, 7, 500, 503):
        for R in (1, 2, 3, 4, 8):
            shares = [range((nb * r) // R, (nb * (r + 1)) // R) for r in range(R)]
            assert [i for s in shares for i in s] == list(range(nb))            # contiguous, disjoint, complete
            assert max(len(s) for s in shares) - min(len(s) for s in shares) <= 1
