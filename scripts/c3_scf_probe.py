"""Development probe: real SCF iterations on the C3 cell (Si 250 atoms, Gamma, Ecut 30) -- SCF-iteration wall time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = ["bench.py"]
import bench
import numpy as np
import torch
import dftk_b200 as dftk

rep = int(os.environ.get("REP", 5))
lat, pos = bench.supercell(rep)
Si = dftk.ElementPsp("Si")
model = dftk.model_DFT(lat, [Si] * len(pos), pos, functionals=dftk.LDA(), symmetries=False)
t0 = time.time()
basis = dftk.PlaneWaveBasis(model, Ecut=30.0, kgrid=dftk.ExplicitKpoints([[0, 0, 0]]))
print("basis setup", round(time.time() - t0, 2), "s", basis.fft_size, basis.kpoints[0].n_G, flush=True)
steps = []


def cb(info):
    d = info["diagonalization"]
    steps.append(info["time_step"])
    print(f"SCF {info['n_iter']}: E = {info['history_Etot'][-1]:.8f}  drho = {info['history_drho'][-1]:.3e}  "
          f"lobpcg iters {d['n_iter']}  matvec {d['n_matvec']}  step {info['time_step']:.2f} s", flush=True)


res = dftk.self_consistent_field(basis, tol=1e-6, maxiter=int(os.environ.get("MAXITER", 3)), callback=cb)
n_el = float(res["rho"].sum() * basis.dvol)
print("electrons", n_el, "E/atom", res["energies"].total / len(pos), "steps", steps, flush=True)
