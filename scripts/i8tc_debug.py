"""Bring-up diagnosis of the tcgen05 INT8 kernel (gemm_backend = 3, dftk.jl_b200/csrc/i8tc.cu).

  DFTK_B200_I8TC_DUMP=/tmp/i8tc [DFTK_B200_I8TC_SWAP=1] [DFTK_B200_I8TC_SIMPLE=1] python scripts/i8tc_debug.py [K m n]

runs one emulated GEMM through the tensor-core backend, then reads the dump of the first output tile (raw s32 accumulators
X1..X4 for modulus 0 and the int8 operand rows that produced them) and reports, per accumulator, how the hardware result
relates to the exact integer products: equal / equal after a permutation of rows in groups of 8 / equal with the K chunks of
16 bytes reordered / equal to a product with a *different* operand pairing -- the usual signatures of a wrong LBO/SBO
convention, a wrong core-matrix order, or mixed-up descriptors.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dftk_b200 as dftk

prefix = os.environ.setdefault("DFTK_B200_I8TC_DUMP", "/tmp/i8tc")
K, m, n = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (512, 40, 24)
ctx = dftk.Context(0)
g = torch.Generator(device="cpu").manual_seed(0)
A = torch.view_as_complex(torch.randn(m, K, 2, generator=g, dtype=torch.float64)).to(ctx.device)
B = torch.view_as_complex(torch.randn(n, K, 2, generator=g, dtype=torch.float64)).to(ctx.device)
ref = torch.zeros((n, m), dtype=torch.complex128, device=ctx.device)
ctx.zgemm("C", A, B, ref)
ctx.set_option("gemm_backend", 3)
C = torch.zeros_like(ref)
try:
    ctx.zgemm("C", A, B, C)
    torch.cuda.synchronize()
    print("backend 3 vs DMMA: max rel err", float((C - ref).abs().max() / ref.abs().max()))
except Exception as e:           # the dump is written before the error is raised when the kernel itself completed
    print("backend 3 raised:", e)
finally:
    ctx.set_option("gemm_backend", 0)

x = np.fromfile(prefix + ".x", dtype=np.int32).reshape(4, 128, 128).astype(np.int64)
a = np.fromfile(prefix + ".a", dtype=np.int8)
b = np.fromfile(prefix + ".b", dtype=np.int8)
kc = a.size // (2 * 128)
a = a.reshape(2, 128, kc).astype(np.int64)
b = b.reshape(2, 128, kc).astype(np.int64)
print(f"dump: k = {kc}; rows beyond m = {m} / n = {n} repeat the last row")
pairs = {"X1 = Ar.Br": (0, 0), "X2 = Ai.Bi": (1, 1), "X3 = Ar.Bi": (0, 1), "X4 = Ai.Br": (1, 0)}
exact = {name: a[pa] @ b[pb].T for name, (pa, pb) in pairs.items()}

for idx, (name, (pa, pb)) in enumerate(pairs.items()):
    got = x[idx]
    want = exact[name]
    if np.array_equal(got, want):
        print(f"{name}: exact")
        continue
    nbad = int((got != want).sum())
    msg = [f"{name}: {nbad} of {got.size} entries differ"]
    for other, w in exact.items():
        if other != name and np.array_equal(got, w):
            msg.append(f"equals {other} (descriptors paired wrongly)")
    if np.array_equal(got, want.T):
        msg.append("equals the transpose (A / B descriptors swapped)")
    if np.array_equal(np.sort(got, axis=0), np.sort(want, axis=0)):
        rows = [int(np.where((want == got[r]).all(axis=1))[0][0]) if (want == got[r]).all(axis=1).any() else -1 for r in range(16)]
        msg.append(f"rows are a permutation of the exact rows; first 16 map to {rows}")
    if np.array_equal(np.sort(got, axis=1), np.sort(want, axis=1)):
        cols = [int(np.where((want.T == got[:, c]).all(axis=1))[0][0]) if (want.T == got[:, c]).all(axis=1).any() else -1 for c in range(16)]
        msg.append(f"columns are a permutation of the exact columns; first 16 map to {cols}")
    partial = a[pa][:, :32] @ b[pb][:, :32].T
    if np.array_equal(got, partial):
        msg.append("equals the product over the first 32 K bytes only (accumulate flag / K loop)")
    good_rows = int((got == want).all(axis=1).sum())
    good_cols = int((got == want).all(axis=0).sum())
    msg.append(f"{good_rows} rows and {good_cols} columns are entirely right; got[0,:4] = {got[0, :4].tolist()} want {want[0, :4].tolist()}")
    print("; ".join(msg))
