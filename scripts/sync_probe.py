"""Development probe: latency of the small synchronising library calls (band_energies, density_accumulate, real_gram) and of a
bare launch + stream synchronisation, before and after batched LOBPCG solves -- to tell a slow call from a slow GPU box."""
import sys, os, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dftk_b200 as dftk

print(subprocess.run(["nvidia-smi", "--query-compute-apps=pid,used_memory", "--format=csv"], capture_output=True, text=True).stdout)
a = 5.131570667152971
lat = np.array([[0, a, a], [a, 0, a], [a, a, 0]])
Si = dftk.ElementPsp("Si")
model = dftk.model_DFT(lat, [Si, Si], [np.ones(3) / 8, -np.ones(3) / 8], functionals=dftk.LDA())
basis = dftk.PlaneWaveBasis(model, Ecut=30, kgrid=(8, 8, 8))
ctx = basis.architecture.ctx
rho = dftk.guess_density(basis)
_, ham = dftk.energy_hamiltonian(basis, None, None, rho=rho)
kb = ham[0].bind()
psi = dftk.random_orbitals(basis, basis.kpoints[0], 7)


def lat_of(fn, n=200):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return 1e6 * (time.perf_counter() - t) / n


x = torch.zeros(8, device=ctx.device)
r = torch.zeros(basis.N, dtype=torch.float64, device=ctx.device)
A = torch.randn(4, 20000, dtype=torch.float64, device=ctx.device)


def report(tag):
    print(tag, "us per call: torch add+sync %.1f | band_energies %.1f | density_accumulate %.1f | real_gram %.1f | apply_h %.1f"
          % (lat_of(lambda: (x.add_(1), torch.cuda.synchronize())), lat_of(lambda: kb.band_energies(psi)),
             lat_of(lambda: kb.density_accumulate(psi, np.ones(7), r)), lat_of(lambda: ctx.real_gram(A, A)),
             lat_of(lambda: (kb.apply_h(psi), torch.cuda.synchronize()))), flush=True)


report("before any LOBPCG:")
res = dftk.diagonalize_all_kblocks(dftk.lobpcg_hyper, ham, 7, tol=1e-4)
report("after one batched solve:")
for rep in range(3):
    ctx.launch_count(reset=True); ctx.sync_count(reset=True)
    torch.cuda.synchronize(); t = time.perf_counter()
    res = dftk.diagonalize_all_kblocks(dftk.lobpcg_hyper, ham, 7, psiguess=res["X"], tol=1e-7)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("batched solve of 29 blocks: %.1f ms, launches %d, rounds %d, iterations %s" % (1e3 * dt, ctx.launch_count(), ctx.sync_count(), res["n_iter"][:6]), flush=True)
report("after four batched solves:")
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,utilization.gpu,power.draw", "--format=csv"], capture_output=True, text=True).stdout)
