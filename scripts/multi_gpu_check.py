"""torchrun worker: k-point-sharded SCF (one rank per GPU, NCCL density allreduce) must reproduce the
single-GPU SCF.  Launched by tests/test_gpu_multi.py and usable standalone:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/multi_gpu_check.py"""
import os
import sys
import json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
import dftk_b200 as dftk

a = 5.131570667152971
lat = np.array([[0, a, a], [a, 0, a], [a, a, 0]])
pos = [np.ones(3) / 8, -np.ones(3) / 8]
Si = dftk.ElementPsp("Si")
temperature = float(os.environ.get("TEMPERATURE", "0.0"))
model = dftk.model_DFT(lat, [Si, Si], pos, functionals=dftk.LDA(), temperature=temperature)
comm = dftk.KpointComm.from_torch_distributed()
arch = dftk.B200(local, comm=comm)
basis = dftk.PlaneWaveBasis(model, Ecut=10, kgrid=(3, 3, 3), architecture=arch, comm_kpts=comm)
res = dftk.self_consistent_field(basis, tol=1e-9, mixing=dftk.KerkerMixing() if temperature > 0 else None)
eig = comm.allgather_object([(basis.krange_thisproc_allspin[i], res["eigenvalues"][i].tolist())
                             for i in range(len(basis.kpoints))])
out = None
if rank == 0:
    arch1 = dftk.B200(local)
    basis1 = dftk.PlaneWaveBasis(model, Ecut=10, kgrid=(3, 3, 3), architecture=arch1)
    ref = dftk.self_consistent_field(basis1, tol=1e-9, mixing=dftk.KerkerMixing() if temperature > 0 else None)
    gathered = dict(x for part in eig for x in part)
    deig = max(np.abs(np.array(gathered[i][:4]) - ref["eigenvalues"][i][:4]).max() for i in range(len(basis1.kpoints)))
    drho = float((res["rho"] - ref["rho"]).norm()) * np.sqrt(basis.dvol)
    out = dict(world=world, dE=abs(res["energies"].total - ref["energies"].total), deig=float(deig), drho=drho,
               E=res["energies"].total, n_iter=res["n_iter"], n_iter_ref=ref["n_iter"], eF=res["eF"], eF_ref=ref["eF"],
               nk_local=len(basis.kpoints), nk_total=len(basis1.kpoints))
    print("MULTIGPU_RESULT " + json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
