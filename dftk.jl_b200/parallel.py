"""k-point sharding and the collectives of the SCF loop (mirror of src/common/mpi.jl:19-31,
src/common/split_evenly.jl:4-19 and the comm_kpts logic of src/PlaneWaveBasis.jl:184-229).

One process per GPU.  Host-side scalars travel through torch.distributed (gloo on CPU-only tests, the
default group otherwise); device arrays (the density) go through ncclAllReduce inside libdftk_b200.
"""
import numpy as np


def split_evenly(itr, N):
    """src/common/split_evenly.jl:4-19: contiguous chunks, the first `rem` chunks one element longer."""
    itr = list(itr)
    n = len(itr)
    if N > n:
        raise ValueError("cannot split into more chunks than elements")
    base, rem = divmod(n, N)
    out, start = [], 0
    for i in range(N):
        ln = base + (1 if i < rem else 0)
        out.append(itr[start:start + ln])
        start += ln
    return out


def pad_kpoints_for_ranks(kcoords, kweights, n_procs):
    """PlaneWaveBasis.jl:190-203: ranks without k-points are forbidden; duplicate the highest-weight
    k-point (halving its weight) until every rank has one."""
    kcoords, kweights = list(kcoords), list(kweights)
    while len(kcoords) < n_procs:
        idx = int(np.argmax(kweights))
        kweights[idx] *= 0.5
        kweights.append(kweights[idx])
        kcoords.append(kcoords[idx])
    return kcoords, kweights


class KpointComm:
    """Communicator over k-point shards (basis.comm_kpts)."""

    def __init__(self, rank=0, nranks=1, group=None, nccl_id=None):
        self.rank, self.nranks, self.group, self.nccl_id = rank, nranks, group, nccl_id

    @staticmethod
    def from_torch_distributed(with_nccl_id=True):
        import torch.distributed as dist
        if not dist.is_initialized():
            return KpointComm()
        rank, n = dist.get_rank(), dist.get_world_size()
        nccl_id = None
        if with_nccl_id and n > 1:
            from .device import Context
            obj = [Context.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(obj, src=0)
            nccl_id = obj[0]
        return KpointComm(rank, n, None, nccl_id)

    # --- host scalars / small arrays (mpi_sum, mpi_min, mpi_max, bcast) ---
    def _all(self, value, op):
        if self.nranks == 1:
            return value
        import torch
        import torch.distributed as dist
        t = torch.as_tensor(np.asarray(value, dtype=np.float64)).clone()
        if dist.get_backend(self.group) == "nccl":          # NCCL groups only move device tensors
            t = t.cuda()
        dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX}[op],
                        group=self.group)
        r = t.cpu().numpy()
        return float(r) if r.ndim == 0 else r

    def sum(self, v): return self._all(v, "sum")
    def min(self, v): return self._all(v, "min")
    def max(self, v): return self._all(v, "max")

    def allgather_object(self, obj):
        if self.nranks == 1:
            return [obj]
        import torch.distributed as dist
        out = [None] * self.nranks
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def bcast_object(self, obj, src=0):
        if self.nranks == 1:
            return obj
        import torch.distributed as dist
        lst = [obj]
        dist.broadcast_object_list(lst, src=src, group=self.group)
        return lst[0]
