"""Short C3-shaped run for ncu captures: a few Hψ-local chunks + one nonlocal apply (see profiles/README.md)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dftk_b200

A = 10.26 / 2
lat = 5 * np.array([[0, A, A], [A, 0, A], [A, A, 0]])
recip = 2 * np.pi * np.linalg.inv(lat.T)
dev = torch.device("cuda:0")
ctx = dftk_b200.Context(0)
g1 = torch.as_tensor(np.array(list(range(0, 96)) + list(range(-96, 0))), device=dev, dtype=torch.float64)
Z, Y, X = torch.meshgrid(g1, g1, g1, indexing="ij")
G = torch.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], 1)
p = G @ torch.as_tensor(recip.T, device=dev)
kin_all = (p * p).sum(1) / 2
mapping = torch.nonzero(kin_all <= 30.0).reshape(-1)
kin = kin_all[mapping].contiguous()
npw = mapping.numel()
grid = dftk_b200.FFTGrid(ctx, (192, 192, 192), abs(np.linalg.det(lat)))
nproj, M = int(os.environ.get("NPROJ", 1250)), int(os.environ.get("M", 128))
gen = torch.Generator(device=dev).manual_seed(0)
P = torch.view_as_complex(torch.randn(nproj, npw, 2, generator=gen, device=dev, dtype=torch.float64)) / np.sqrt(npw)
kb = dftk_b200.KBlock(grid, mapping.cpu().numpy(), kin=kin, P=P, D=np.eye(nproj))
kb.set_potential(torch.cos(torch.arange(192 ** 3, device=dev, dtype=torch.float64) * 1e-3))
psi = torch.view_as_complex(torch.randn(M, npw, 2, generator=gen, device=dev, dtype=torch.float64))
out = torch.empty_like(psi)
for _ in range(int(os.environ.get("REPS", 3))):
    kb.apply_terms(psi, 3, out=out)
    kb.apply_terms(psi, 4, out=out)
rho = torch.zeros(192 ** 3, dtype=torch.float64, device=dev)
kb.density_accumulate(psi[:64], np.ones(64), rho)
torch.cuda.synchronize()
print("done", ctx.launch_count())
