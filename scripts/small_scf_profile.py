"""Development probe: where the wall time of an SCF goes for the small BASELINE configs (C1/C2/C4/C5 shapes).
  CASE=si2|al4|fe  ECUT=..  KGRID=..   python scripts/small_scf_profile.py
Prints: SCF wall time / iteration, kernel launches, a host-side cProfile of one SCF (cumulative, top functions) and --
when DFTK_B200_PROFILE=2 is set -- the LOBPCG section profile summed over all solves (printed by the library at exit).
"""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dftk_b200 as dftk

case = os.environ.get("CASE", "si2")
mixing = None
if case == "si2":
    a = 5.131570667152971
    lat = np.array([[0, a, a], [a, 0, a], [a, a, 0]])
    model = dftk.model_DFT(lat, [dftk.ElementPsp("Si")] * 2, [np.ones(3) / 8, -np.ones(3) / 8], functionals=dftk.LDA())
    Ecut, kg = float(os.environ.get("ECUT", 30)), int(os.environ.get("KGRID", 8))
elif case == "al4":
    lat = 7.65339 * np.eye(3)
    Al = dftk.ElementPsp("Al", functional="pbe")
    model = dftk.model_DFT(lat, [Al] * 4, [[0, 0, 0], [0, 0.5, 0.5], [0.5, 0, 0.5], [0.5, 0.5, 0]], functionals=dftk.PBE(),
                           temperature=0.01)
    Ecut, kg = float(os.environ.get("ECUT", 40)), int(os.environ.get("KGRID", 12))
    mixing = dftk.KerkerMixing()
else:
    lat = 2.71176 * np.array([[-1, 1, 1], [1, -1, 1], [1, 1, -1]], dtype=float)
    model = dftk.model_DFT(lat, [dftk.ElementPsp("Fe", functional="pbe")], [[0, 0, 0]], functionals=dftk.PBE(),
                           temperature=0.01, magnetic_moments=[4.0])
    Ecut, kg = float(os.environ.get("ECUT", 45)), int(os.environ.get("KGRID", 8))
    mixing = dftk.KerkerMixing()
t0 = time.time()
basis = dftk.PlaneWaveBasis(model, Ecut=Ecut, kgrid=(kg, kg, kg))
print(f"[{case}] basis {time.time() - t0:.2f} s; k-blocks {len(basis.kpoints)} fft {basis.fft_size} n_G {basis.kpoints[0].n_G} "
      f"n_proj {basis.kblocks[0].n_proj}", flush=True)
ctx = basis.architecture.ctx
tol = float(os.environ.get("TOL", 1e-8))
for rep in range(2):
    ctx.launch_count(reset=True)
    torch.cuda.synchronize()
    t0 = time.time()
    pr = cProfile.Profile() if rep == 1 else None
    if pr:
        pr.enable()
    res = dftk.self_consistent_field(basis, tol=tol, mixing=mixing)
    torch.cuda.synchronize()
    if pr:
        pr.disable()
    dt = time.time() - t0
    nb = res["psi"][0].shape[0]
    print(f"[{case}] SCF run {rep}: {dt:.2f} s, {res['n_iter']} iterations, {dt / res['n_iter']:.3f} s/iter, bands {nb}, "
          f"E = {res['energies'].total:.10f}, n_matvec {res['n_matvec']}, launches {ctx.launch_count()}", flush=True)
    if pr:
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
        print(s.getvalue()[:6000], flush=True)
