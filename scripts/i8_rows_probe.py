"""Development probe (one GPU): how the INT8 tensor-core products scale with the number of plane-wave rows -- the slab sizes
of the single-k multi-GPU solve (264 859 rows over 1, 2, 4, 8 ranks).  Gram-type C = A^H B (1509 x 1509) and update-type
C = A B (rows x 503 from 1509 columns), gemm_backend 4 against the FP64 DMMA kernels (backend 0)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dftk_b200

dev = torch.device("cuda:0")
ctx = dftk_b200.Context(0)
g = torch.Generator(device=dev).manual_seed(0)


def timeit(fn, n=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


out = {}
for rows in (264859, 132430, 66215, 33108):
    A = torch.view_as_complex(torch.randn(1509, rows, 2, generator=g, device=dev, dtype=torch.float64)) / np.sqrt(rows)
    Bs = torch.view_as_complex(torch.randn(503, 1509, 2, generator=g, device=dev, dtype=torch.float64)) / np.sqrt(1509)
    G = torch.zeros((1509, 1509), dtype=torch.complex128, device=dev)
    U = torch.zeros((503, rows), dtype=torch.complex128, device=dev)
    r = {}
    for backend in (4, 0):
        ctx.set_option("gemm_backend", backend)
        r[f"gram_ms_b{backend}"] = timeit(lambda: ctx.zgemm("C", A, A, G))
        r[f"update_ms_b{backend}"] = timeit(lambda: ctx.zgemm("N", A, Bs, U))
    ctx.set_option("gemm_backend", 4)
    out[rows] = r
    print(rows, r, flush=True)
    del A, G, U
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/i8_rows_probe.json", "w"), indent=1)
