"""Development probe: where the wall time of a small-k-block SCF goes on the host side (cProfile of the second, warm SCF of a
BASELINE config; CONFIG=C2|C4|C5, PIPE=batch_pipeline option)."""
import os, sys, time, cProfile, pstats, io
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
pr = cProfile.Profile()
t = time.perf_counter()
pr.enable()
res = dftk.self_consistent_field(basis, tol=1e-8, mixing=mixing, seed=3)
torch.cuda.synchronize()
pr.disable()
print(f"{name}: {time.perf_counter() - t:.3f} s under cProfile, {res['n_iter']} steps")
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
