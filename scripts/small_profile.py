"""Development probe: kernel-time distribution of the batched small-k-block path.  One full SCF of a BASELINE config
(CONFIG=C2|C4|C5) runs unprofiled (warm), then a second one inside cudaProfilerStart/Stop for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file ... python scripts/small_profile.py
and -- without ncu -- prints wall time, launches and host syncs per SCF step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
sys.argv = ["bench.py"]
import bench
import dftk_b200 as dftk

name = os.environ.get("CONFIG", "C4")
model, bk, desc = bench.baseline_model(dftk, name)
basis = dftk.PlaneWaveBasis(model, **bk)
mixing = dftk.KerkerMixing() if model.temperature > 0 else None
ctx = basis.architecture.ctx
ctx.set_option("batch_pipeline", int(os.environ.get("PIPE", "1")))
dftk.self_consistent_field(basis, tol=1e-8, mixing=mixing, seed=3)
torch.cuda.synchronize()
ctx.launch_count(reset=True); ctx.sync_count(reset=True)
steps = []
torch.cuda.profiler.start()
t = time.perf_counter()
res = dftk.self_consistent_field(basis, tol=1e-8, mixing=mixing, seed=3, maxiter=int(os.environ.get("MAXITER", 100)),
                                 callback=lambda info: steps.append((info["time_step"], int(np.sum(info["diagonalization"]["n_iter"])))))
torch.cuda.synchronize()
dt = time.perf_counter() - t
torch.cuda.profiler.stop()
n = res["n_iter"]
print(f"{name} batch_pipeline={os.environ.get('PIPE', '1')}: {desc}: {dt:.3f} s, {n} SCF steps, {ctx.launch_count() / n:.0f} launches and {ctx.sync_count() / n:.0f} LOBPCG host syncs per step, "
      f"blocks {len(basis.kpoints)}, bands {res['psi'][0].shape[0]}", flush=True)
print("per step (s, summed LOBPCG iterations):", [(round(a, 4), b) for a, b in steps], flush=True)
