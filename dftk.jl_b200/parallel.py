"""(k-point, spin) block sharding and the collectives of the SCF loop (mirror of src/common/mpi.jl:19-31,
src/common/split_evenly.jl:4-19 and the comm_kpts logic of src/PlaneWaveBasis.jl:184-229).

One process per GPU.  Per SCF step the sharded path needs exactly two data collectives, both inside libdftk_b200
(NCCL over NVLink): one `dftk_b200_allgather` of the eigenvalues (+ solver statistics) of every rank's blocks, and one
`dftk_b200_allreduce` of the density with the k-summed band-energy partials packed behind it.  The Fermi
level is then bisected redundantly on every rank.  torch.distributed is only the transport of the CPU-only tests
(gloo, no device context attached) and of the one-off NCCL id broadcast.
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


def pad_kpoints_for_ranks(kcoords, kweights, n_procs, n_spin=1):
    """PlaneWaveBasis.jl:190-203: ranks without work are forbidden; duplicate the highest-weight k-point (halving its
    weight) until every rank can own a (k, spin) block."""
    kcoords, kweights = list(kcoords), list(kweights)
    while len(kcoords) * n_spin < n_procs:
        idx = int(np.argmax(kweights))
        kweights[idx] *= 0.5
        kweights.append(kweights[idx])
        kcoords.append(kcoords[idx])
    return kcoords, kweights


def lpt_assign(costs, n_ranks):
    """Longest-processing-time assignment of blocks to ranks: blocks in order of decreasing cost, each to
    the least-loaded rank; ties go to the lower block index / rank, so the map is deterministic and identical on every
    rank.  Returns owner[b]."""
    order = sorted(range(len(costs)), key=lambda b: (-costs[b], b))
    load = [0.0] * n_ranks
    count = [0] * n_ranks
    owner = [0] * len(costs)
    for b in order:
        r = min(range(n_ranks), key=lambda q: (load[q], count[q], q))
        owner[b] = r
        load[r] += costs[b]
        count[r] += 1
    return owner


class BlockLayout:
    """Which rank owns which (k, spin) block.  Global block index b = ik + spin * n_kpt (the reference's
    krange_allspin numbering, PlaneWaveBasis.jl:226-229)."""

    def __init__(self, n_kpt, n_spin, kweights_global, costs, n_ranks, rank):
        self.n_kpt, self.n_spin, self.n_ranks, self.rank = n_kpt, n_spin, n_ranks, rank
        self.n_blocks = n_kpt * n_spin
        self.weights = [kweights_global[b % n_kpt] for b in range(self.n_blocks)]
        self.owner = lpt_assign(list(costs), n_ranks) if n_ranks > 1 else [0] * self.n_blocks
        self.blocks_of_rank = [[b for b in range(self.n_blocks) if self.owner[b] == r] for r in range(n_ranks)]
        if any(len(x) == 0 for x in self.blocks_of_rank):
            raise ValueError("a rank would own no (k, spin) block")
        self.mine = self.blocks_of_rank[rank]                 # ascending: spin-major, then k (local block order)
        self.max_local = max(len(x) for x in self.blocks_of_rank)


class KpointComm:
    """Communicator over (k, spin) block shards (basis.comm_kpts)."""

    def __init__(self, rank=0, nranks=1, group=None, nccl_id=None):
        self.rank, self.nranks, self.group, self.nccl_id = rank, nranks, group, nccl_id
        self.ctx = None          # device context with the NCCL communicator (attached by the architecture)
        self.n_collectives = 0   # data collectives issued (diagnostics: the bench reports them per SCF step)

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

    def attach(self, ctx):
        """Route every collective through the NCCL communicator of this device context."""
        if self.nranks > 1:
            assert ctx.nranks == self.nranks and ctx.rank == self.rank
            self.ctx = ctx

    # --- small host arrays: one collective each (callers pack what belongs together) ---
    def allreduce(self, values, op="sum"):
        """Elementwise reduction of a small float64 host array over the ranks."""
        a = np.array(values, dtype=np.float64)
        if self.nranks == 1:
            return a
        self.n_collectives += 1
        if self.ctx is not None:
            import torch
            t = torch.from_numpy(np.ascontiguousarray(a).reshape(-1)).to(self.ctx.device)
            self.ctx.allreduce(t, op)
            return t.cpu().numpy().reshape(a.shape)
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(np.ascontiguousarray(a).reshape(-1).copy())
        dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX}[op],
                        group=self.group)
        return t.numpy().reshape(a.shape)

    def allgather(self, values):
        """Fixed-size allgather of a float64 host array: returns shape (nranks, *values.shape)."""
        a = np.ascontiguousarray(np.array(values, dtype=np.float64))
        if self.nranks == 1:
            return a[None]
        self.n_collectives += 1
        import torch
        if self.ctx is not None:
            send = torch.from_numpy(a.reshape(-1)).to(self.ctx.device)
            recv = torch.empty(self.nranks * send.numel(), dtype=torch.float64, device=self.ctx.device)
            self.ctx.allgather(send, recv)
            return recv.cpu().numpy().reshape((self.nranks,) + a.shape)
        import torch.distributed as dist
        send = torch.from_numpy(a.reshape(-1).copy())
        out = [torch.empty_like(send) for _ in range(self.nranks)]
        dist.all_gather(out, send, group=self.group)
        return torch.stack(out).numpy().reshape((self.nranks,) + a.shape)

    def _scalar(self, v, op):
        r = self.allreduce(np.asarray(v, dtype=np.float64), op)
        return float(r) if r.ndim == 0 else r

    def sum(self, v): return self._scalar(v, "sum")
    def min(self, v): return self._scalar(v, "min")
    def max(self, v): return self._scalar(v, "max")

    def all_true(self, flag):
        """Logical AND over the ranks (the mpi bcast of `converged`, self_consistent_field.jl:222, made symmetric)."""
        return bool(self.min(1.0 if flag else 0.0) > 0.5)

    def bcast_object(self, obj, src=0):
        """Setup-time broadcast of a small Python object (seeds); torch.distributed object transport."""
        if self.nranks == 1:
            return obj
        import torch.distributed as dist
        lst = [obj]
        dist.broadcast_object_list(lst, src=src, group=self.group)
        return lst[0]
