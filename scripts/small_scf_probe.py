"""Development probe: wall time of a full SCF for the small BASELINE configs (C2-like) on the GPU vs the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dftk_b200 as dftk

a = 5.131570667152971
lat = np.array([[0, a, a], [a, 0, a], [a, a, 0]])
pos = [np.ones(3) / 8, -np.ones(3) / 8]
Si = dftk.ElementPsp("Si")
Ecut, kg = float(os.environ.get("ECUT", 30)), int(os.environ.get("KGRID", 8))
model = dftk.model_DFT(lat, [Si, Si], pos, functionals=dftk.LDA())
t0 = time.time()
basis = dftk.PlaneWaveBasis(model, Ecut=Ecut, kgrid=(kg, kg, kg))
print("basis", time.time() - t0, "s; kpoints", len(basis.kpoints), "fft", basis.fft_size, "n_G", basis.kpoints[0].n_G, flush=True)
ctx = basis.architecture.ctx
for rep in range(2):
    ctx.launch_count(reset=True)
    torch.cuda.synchronize(); t0 = time.time()
    res = dftk.self_consistent_field(basis, tol=1e-8, callback=dftk.ScfDefaultCallback() if rep == 1 else None)
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"GPU SCF: {dt:.2f} s, {res['n_iter']} iterations, {dt / res['n_iter']:.3f} s/iter, E = {res['energies'].total:.10f}, "
          f"n_matvec {res['n_matvec']}, launches {ctx.launch_count()}", flush=True)
if os.environ.get("ORACLE", "1") == "1":
    from oracle.basis import Element, Model, PlaneWaveBasis as OBasis
    from oracle import scf as oscf
    om = Model(lat, [Element("Si")] * 2, pos, functionals=("lda_x", "lda_c_pw"))
    ob = OBasis(om, Ecut, kgrid=(kg, kg, kg))
    t0 = time.time()
    ores = oscf.self_consistent_field(ob, tol=1e-8)
    dt = time.time() - t0
    print(f"CPU oracle SCF ({os.cpu_count()} cores, NumPy): {dt:.2f} s, {ores['n_iter']} iterations, {dt / ores['n_iter']:.3f} s/iter, "
          f"E = {ores['energies']['total']:.10f}, dE = {abs(ores['energies']['total'] - res['energies'].total):.2e}", flush=True)
