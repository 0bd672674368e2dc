"""Development probe: the fused z stage of the H apply at the C3 shape, one tile per CTA (z_pipeline 0) against the persistent
software-pipelined kernel (z_pipeline 1): time of the local+kinetic apply and agreement of the results."""
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
n = 192
g1 = torch.as_tensor(np.array(list(range(0, 96)) + list(range(-96, 0))), device=dev, dtype=torch.float64)
Z, Y, X = torch.meshgrid(g1, g1, g1, indexing="ij")
G = torch.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], 1)
p = G @ torch.as_tensor(recip.T, device=dev)
kin_all = (p * p).sum(1) / 2
mapping = torch.nonzero(kin_all <= 30.0).reshape(-1)
kin = kin_all[mapping].contiguous()
npw = mapping.numel()
grid = dftk_b200.FFTGrid(ctx, (n, n, n), abs(np.linalg.det(lat)))
kb = dftk_b200.KBlock(grid, mapping.cpu().numpy(), kin=kin)
kb.set_potential(torch.cos(torch.arange(n ** 3, device=dev, dtype=torch.float64) * 0.001))
g = torch.Generator(device=dev).manual_seed(0)
nb = 153
psi = torch.view_as_complex(torch.randn(nb, npw, 2, generator=g, device=dev, dtype=torch.float64))
outs = {}
for mode in (0, 1, 0, 1):
    ctx.set_option("z_pipeline", mode)
    out = torch.empty_like(psi)
    kb.apply_terms(psi, 3, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); kb.apply_terms(psi, 3, out=out); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = min(ts)
    outs[mode] = out
    print("z_pipeline", mode, "local+kinetic %.3f ms for %d bands = %.1f us/band, %.0f GB/s algorithmic"
          % (t, nb, 1e3 * t / nb, (72 * n ** 3 + 40 * npw) * nb / t / 1e6), flush=True)
print("max |difference| between the two z kernels:", float((outs[0] - outs[1]).abs().max()), "of", float(outs[0].abs().max()))
