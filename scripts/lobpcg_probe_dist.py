"""torchrun variant of lobpcg_probe.py (every rank solves its own C3-shaped block) to study host-side contention."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
if local == 0:
    os.environ["DFTK_B200_PROFILE"] = "1"
sys.argv = ["bench.py"]
import bench
import dftk_b200 as dftk
lat, pos = bench.supercell(5)
Si = dftk.ElementPsp("Si")
model = dftk.model_DFT(lat, [Si] * len(pos), pos, functionals=dftk.LDA(), symmetries=False)
arch = dftk.B200(local)
basis = dftk.PlaneWaveBasis(model, Ecut=30.0, kgrid=dftk.ExplicitKpoints([[0, 0, 0]]), architecture=arch)
_, ham = dftk.energy_hamiltonian(basis, None, None, rho=dftk.guess_density(basis))
kb = ham[0].bind()
X = dftk.random_orbitals(basis, basis.kpoints[0], 503)
torch.cuda.synchronize()
if dist.is_initialized():
    dist.barrier()
t = time.perf_counter()
res = kb.lobpcg(X, tol=0.025, maxiter=int(os.environ.get("MAXITER", 3)), n_conv_check=500)
torch.cuda.synchronize()
print(f"rank {local} OMP={os.environ.get('OMP_NUM_THREADS')} policy={os.environ.get('OMP_WAIT_POLICY')} lobpcg {time.perf_counter() - t:.2f} s", flush=True)
if dist.is_initialized():
    dist.destroy_process_group()
