"""Development probe: time the hot kernels at the C3 shape (Si 5x5x5 supercell, Ecut 30, Gamma)."""
import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dftk_b200

A = 10.26 / 2
lat = 5 * np.array([[0, A, A], [A, 0, A], [A, A, 0]])
recip = 2 * np.pi * np.linalg.inv(lat.T)
fft = (192, 192, 192)
Ecut = 30.0
dev = torch.device("cuda:0")
ctx = dftk_b200.Context(0)


def gax(n):
    return np.array(list(range(0, (n - 1) // 2 + 1)) + list(range(-(n // 2), 0)))


t0 = time.time()
gx = torch.as_tensor(gax(192), device=dev, dtype=torch.float64)
Z, Y, X = torch.meshgrid(gx, gx, gx, indexing="ij")
G = torch.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], 1)
p = G @ torch.as_tensor(recip.T, device=dev)
kin_all = (p * p).sum(1) / 2
mapping = torch.nonzero(kin_all <= Ecut).reshape(-1)
kin = kin_all[mapping].contiguous()
npw = mapping.numel()
print("n_pw", npw, "setup", time.time() - t0, flush=True)
vol = abs(np.linalg.det(lat))
grid = dftk_b200.FFTGrid(ctx, fft, vol)
nproj = int(os.environ.get("NPROJ", 1250))
M = int(os.environ.get("M", 503))
g = torch.Generator(device=dev).manual_seed(0)
P = torch.view_as_complex(torch.randn(nproj, npw, 2, generator=g, device=dev, dtype=torch.float64)) / np.sqrt(npw)
D = np.diag(np.random.default_rng(0).standard_normal(nproj))
kb = dftk_b200.KBlock(grid, mapping.cpu().numpy(), kin=kin, P=P, D=D)
V = torch.cos(torch.arange(192 ** 3, device=dev, dtype=torch.float64) * 0.001)
kb.set_potential(V)
psi = torch.view_as_complex(torch.randn(M, npw, 2, generator=g, device=dev, dtype=torch.float64))
out = torch.empty_like(psi)


def timeit(fn, n=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))


res = {}
for lines in (0,):
    for chunk in (16, 51):
        ctx.set_option("band_chunk", chunk)
        nb = 153
        t, _ = timeit(lambda: kb.apply_terms(psi[:nb], 3, out=out[:nb]))
        bytes_alg = (72 * 192 ** 3 + 40 * npw) * nb
        res[f"local_kin_L{lines}_chunk{chunk}"] = dict(ms=t, us_per_band=t * 1e3 / nb, GBs_alg=bytes_alg / t / 1e6)
        print("L", lines, "chunk", chunk, res[f"local_kin_L{lines}_chunk{chunk}"], flush=True)
ctx.set_option("band_chunk", 0)
for backend in (0, 2, 1):
    ctx.set_option("gemm_backend", 1 if backend == 1 else 0)
    ctx.set_option("gemm_stages", 3 if backend == 2 else 2)
    t, _ = timeit(lambda: kb.apply_terms(psi, 4, out=out), n=2)
    fl = 16.0 * npw * nproj * M
    res[f"nonlocal_backend{backend}"] = dict(ms=t, TFLOPs=fl / t / 1e9)
    print("nonlocal", backend, res[f"nonlocal_backend{backend}"], flush=True)
    # Gram 1509 x 1509 and update
    m3 = min(3 * M, psi.shape[0] * 3)
    Yb = torch.cat([psi, psi, psi], 0)[:m3].contiguous()
    C = torch.empty(m3, m3, dtype=torch.complex128, device=dev)
    t, _ = timeit(lambda: ctx.zgemm("C", Yb, Yb, C), n=2)
    res[f"gram_backend{backend}"] = dict(ms=t, TFLOPs=8.0 * npw * m3 * m3 / t / 1e9)
    print("gram", backend, res[f"gram_backend{backend}"], flush=True)
    S = torch.view_as_complex(torch.randn(M, m3, 2, generator=g, device=dev, dtype=torch.float64))
    t, _ = timeit(lambda: ctx.zgemm("N", Yb, S, out), n=2)
    res[f"update_backend{backend}"] = dict(ms=t, TFLOPs=8.0 * npw * m3 * M / t / 1e9)
    print("update", backend, res[f"update_backend{backend}"], flush=True)
    del Yb, C, S
ctx.set_option("gemm_backend", 0)
rho = torch.zeros(192 ** 3, dtype=torch.float64, device=dev)
w = np.ones(128)
t, _ = timeit(lambda: kb.density_accumulate(psi[:128], w, rho))
res["density"] = dict(ms=t, us_per_band=t * 1e3 / 128, GBs_alg=(32 * 192 ** 3 + 16 * npw) * 128 / t / 1e6)
print("density", res["density"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/perf_probe.json", "w"), indent=1)
