"""Development probe: cuSOLVER Zheevd (1509 x 1509, the Rayleigh-Ritz size of the C3 block) under different OpenMP
settings / import orders.  Each case runs in a fresh interpreter."""
import os, subprocess, sys

CODE = r'''
import os, sys, time
ORDER = os.environ["PROBE_ORDER"]
sys.path.insert(0, os.getcwd())
if ORDER == "pkg_first":
    import dftk_b200
import torch
if ORDER == "torch_first":
    import dftk_b200
n = 1509
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.view_as_complex(torch.randn(n, n, 2, generator=g, device="cuda", dtype=torch.float64))
A = A + A.conj().T
torch.linalg.eigh(A); torch.cuda.synchronize()
ts = []
for _ in range(4):
    t = time.perf_counter(); torch.linalg.eigh(A); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
print(f"{os.environ['PROBE_NAME']:42s} eigh 1509: min {min(ts)*1e3:7.1f} ms  median {sorted(ts)[2]*1e3:7.1f} ms", flush=True)
'''
cases = [("pkg_first (env set before libgomp)", "pkg_first", {}),
         ("torch_first (runtime team cap only)", "torch_first", {}),
         ("torch only, OMP_NUM_THREADS=1", "none", {"OMP_NUM_THREADS": "1"}),
         ("torch only, defaults", "none", {})]
for name, order, extra in cases:
    env = {k: v for k, v in os.environ.items() if not k.startswith("OMP_")}
    env.update(extra, PROBE_ORDER=order, PROBE_NAME=name)
    subprocess.run([sys.executable, "-c", CODE], env=env)
