"""compute_density + symmetrize_ρ (mirror of src/densities.jl:13-57 and src/symmetry.jl:282-357).
The per-band |IFFT ψ|² accumulation is one batched libdftk_b200 call per k-block; the cross-rank sum is
a single ncclAllReduce on the device-resident density (the mpi_sum!(ρ) of densities.jl:46)."""
import math
import numpy as np
import torch


def compute_density(basis, psi, occupation, *, occupation_threshold=0.0):
    n_spin = basis.model.n_spin_components
    dev = basis.architecture.device
    rho = torch.zeros((n_spin, basis.N), dtype=torch.float64, device=dev)
    for ik, kb in enumerate(basis.kblocks):
        occ = np.asarray(occupation[ik], dtype=float)
        w = np.where(np.abs(occ) >= occupation_threshold, occ * basis.kweights[ik], 0.0)
        nb = int(np.max(np.nonzero(w)[0]) + 1) if np.any(w != 0) else 0
        if nb:
            kb.density_accumulate(psi[ik][:nb], w[:nb], rho[basis.kpoints[ik].spin])
    if basis.comm_kpts.nranks > 1:
        basis.architecture.ctx.allreduce(rho, "sum")
    rho = symmetrize_rho(basis, rho)
    return rho


def symmetrize_rho(basis, rho):
    """symmetry.jl:340-357 with do_lowpass=false: average over the basis symmetries in Fourier space."""
    syms = basis.symmetries
    if all(s.isone() for s in syms):
        return rho
    if not hasattr(basis, "_sym_tables"):
        tabs = []
        Gf = basis.G_vectors.to(torch.float64)
        for s in syms:
            invS = torch.as_tensor(np.rint(np.linalg.inv(s.S)), device=rho.device, dtype=torch.float64)
            idx = basis.index_G_vectors((Gf @ invS.T).round().to(torch.int64))   # (no int64 matmul on CUDA)
            phase = None
            if np.any(np.abs(s.tau) > 1e-12):
                ph = -2 * math.pi * (Gf @ torch.as_tensor(s.tau, device=rho.device))
                phase = torch.polar(torch.ones_like(ph), ph)
            tabs.append((idx.clamp_min(0), idx >= 0, phase))
        basis._sym_tables = tabs
    rf = basis.fft(rho)
    acc = torch.zeros_like(rf)
    for idx, ok, phase in basis._sym_tables:
        val = torch.where(ok[None, :], rf[:, idx], torch.zeros_like(rf))
        acc += val if phase is None else val * phase[None, :]
    return basis.irfft(acc / len(syms))
