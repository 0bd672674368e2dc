"""Development probe (torchrun, N ranks): section timing (DFTK_B200_PROFILE=1) of the plane-wave-slab LOBPCG at the C3 shape
against the same solve on one GPU.  Every rank writes its sections to gpurun_out/slab_probe_rank<r>.log."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.makedirs("gpurun_out", exist_ok=True)
log = open(f"gpurun_out/slab_probe_rank{rank}{os.environ.get('TAG', '')}.log", "w")
os.dup2(log.fileno(), 2)
if os.environ.get("PROFILE", "1") == "1":
    os.environ["DFTK_B200_PROFILE"] = "1"
import numpy as np
import torch
import torch.distributed as dist
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
sys.argv = ["bench.py"]
import bench
import dftk_b200 as dftk

lat, pos = bench.supercell(int(os.environ.get("REP", 5)))
Si = dftk.ElementPsp("Si")
model = dftk.model_DFT(lat, [Si] * len(pos), pos, functionals=dftk.LDA(), symmetries=False)
comm = dftk.KpointComm.from_torch_distributed()
basis = dftk.PlaneWaveBasis(model, Ecut=30.0, kgrid=(1, 1, 1), comm_slab=comm)
_, ham = dftk.energy_hamiltonian(basis, None, None, rho=dftk.guess_density(basis))
kb = ham[0].bind()
M = bench.n_bands_for(len(pos))
g = torch.Generator(device=basis.architecture.device).manual_seed(4242)
X0 = torch.view_as_complex(torch.randn(M, kb.n_pw, 2, generator=g, device=basis.architecture.device, dtype=torch.float64))
maxiter = int(os.environ.get("MAXITER", 6))
for name, solve in (("slab", kb.lobpcg_slab), ("one_gpu", kb.lobpcg)):
    solve(X0.clone(), tol=1.0, maxiter=1, n_conv_check=M - 3)       # warm workspaces
    X = X0.clone()
    torch.cuda.synchronize(); dist.barrier()
    print(f"==== {name}", file=sys.stderr, flush=True)
    for rep in range(int(os.environ.get("REPEATS", 1))):
        X.copy_(X0)
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = solve(X, tol=0.025, maxiter=maxiter, n_conv_check=M - 3)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        print(f"{name} repeat {rep}: {dt:.3f} s", file=sys.stderr, flush=True)
    print(f"{name}: {dt:.3f} s, {r['n_iter']} iterations -> {dt / max(1, r['n_iter']):.4f} s/iteration (sections synchronise: slower than untimed)",
          file=sys.stderr, flush=True)
dist.barrier()
dist.destroy_process_group()
