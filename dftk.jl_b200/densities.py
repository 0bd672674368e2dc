"""compute_density + symmetrize_ρ (mirror of src/densities.jl:13-57 and src/symmetry.jl:282-357).
The per-band |IFFT ψ|² accumulation is one batched libdftk_b200 call per k-block; the cross-rank sum is
a single ncclAllReduce on the device-resident density (the mpi_sum!(ρ) of densities.jl:46)."""
import math
import numpy as np
import torch


def compute_density(basis, psi, occupation, *, occupation_threshold=0.0, packed_sums=None):
    """`packed_sums`: a few host scalars (this rank's partial sums of the k-summed energies) that ride behind the density
    in the SAME allreduce; the call then returns (rho, summed scalars)."""
    n_spin = basis.model.n_spin_components
    dev = basis.architecture.device
    n_tail = 0 if packed_sums is None else len(packed_sums)
    flat = torch.zeros(n_spin * basis.N + n_tail, dtype=torch.float64, device=dev)
    rho = flat[:n_spin * basis.N].view(n_spin, basis.N)
    slab = getattr(basis, "comm_slab", None)
    if n_tail:
        tail = np.asarray(packed_sums, dtype=np.float64)
        if slab is not None:
            tail = tail / slab.nranks        # replicated partial sums: the allreduce below adds them up again
        flat[n_spin * basis.N:] = torch.as_tensor(tail, device=dev)
    from .device import density_accumulate_multi
    ws, psis = [], []
    for ik in range(len(basis.kblocks)):
        occ = np.asarray(occupation[ik], dtype=float)
        w = np.where(np.abs(occ) >= occupation_threshold, occ * basis.kweights[ik], 0.0)
        nb = int(np.max(np.nonzero(w)[0]) + 1) if np.any(w != 0) else 0
        if slab is not None:                 # every rank accumulates its share of the bands of every block
            lo, hi = (nb * slab.rank) // slab.nranks, (nb * (slab.rank + 1)) // slab.nranks
            ws.append(w[lo:hi])
            psis.append(psi[ik].contiguous()[lo:hi])
            continue
        ws.append(w[:nb])
        psis.append(psi[ik].contiguous())
    # one library call for all blocks of this rank (rows psi[ik][:nb] are contiguous: a band is a row)
    density_accumulate_multi(basis.kblocks, psis, ws, rho)
    if slab is not None:
        slab.n_collectives += 1
        basis.architecture.ctx.allreduce(flat, "sum")
    elif basis.comm_kpts.nranks > 1:
        basis.comm_kpts.n_collectives += 1
        basis.architecture.ctx.allreduce(flat, "sum")          # mpi_sum!(ρ) of densities.jl:46 + the packed scalars
    sums = flat[n_spin * basis.N:].cpu().numpy() if n_tail else None
    rho = symmetrize_rho(basis, rho)
    return rho if packed_sums is None else (rho, sums)


def symmetrize_rho(basis, rho):
    """symmetry.jl:340-357 with do_lowpass=false: average over the basis symmetries in Fourier space
    (one fused gather kernel over all symmetry operations, dftk_b200_symmetrize_fourier)."""
    from ._lib import check
    from .device import _ptr
    syms = basis.symmetries
    if all(s.isone() for s in syms):
        return rho
    if not hasattr(basis, "_sym_tables"):
        invS = np.ascontiguousarray(np.stack([np.rint(np.linalg.inv(s.S)).astype(np.int32) for s in syms]))
        tau = np.ascontiguousarray(np.stack([np.where(np.abs(s.tau) > 1e-12, s.tau, 0.0) for s in syms]))
        basis._sym_tables = (invS, tau)
    invS, tau = basis._sym_tables
    rf = basis.fft(rho).contiguous()
    out = torch.empty_like(rf)
    ctx = basis.architecture.ctx
    for sp in range(rho.shape[0]):
        check(ctx.L.dftk_b200_symmetrize_fourier(basis.fft_grid.h, _ptr(rf[sp]), _ptr(out[sp]), len(syms),
                                                 _ptr(invS), _ptr(tau)), ctx.h)
    return basis.irfft(out)
