"""lobpcg_hyper / diagonalize_all_kblocks (mirror of src/eigen/diag.jl:9-65 and
src/eigen/diag_lobpcg_hyper.jl:5-18).  The eigensolver itself runs inside libdftk_b200
(dftk_b200_lobpcg); this file only chooses start vectors and packs results like the reference."""
import math
import numpy as np
import torch


def _draw_seed(generator):
    """One 63-bit seed from the SCF's torch generator (deterministic per rank and call order)."""
    if generator is None:
        return int(torch.randint(0, 2 ** 62, (1,)).item())
    return int(torch.randint(0, 2 ** 62, (1,), device=generator.device, generator=generator).item())


def random_orbitals(basis, kpt, howmany, generator=None):
    """orbitals.jl:82-87 (ortho_qr(randn)): filled and orthonormalised by libdftk_b200 (counter-based normal numbers +
    the eigensolver's Cholesky-QR ortho!), stored (n_bands, n_G)."""
    from .device import random_orbitals_multi
    ik = next(i for i, k in enumerate(basis.kpoints) if k is kpt)
    return random_orbitals_multi([basis.kblocks[ik]], howmany, _draw_seed(generator))[0]


def lobpcg_hyper(A, X0, *, prec=True, tol=None, maxiter=100, miniter=1, n_conv_check=None):
    """Eigensolver with the reference's call signature `eigensolver(A, X0; prec, tol, maxiter, ...)`;
    returns (; λ, X, residual_norms, n_iter, converged, n_matvec).  X0 is consumed (updated in place)."""
    if tol is None:
        tol = 20 * A.shape[1] * np.finfo(float).eps
    kw = dict(tol=tol, miniter=miniter, maxiter=maxiter, n_conv_check=n_conv_check, prec=bool(prec))
    if _use_slabs(A, X0):
        return A.bind().lobpcg_slab(X0, **kw)
    return A.bind().lobpcg(X0, **kw)


def _use_slabs(A, X0):
    """Plane-wave-slab solve over all ranks of basis.comm_slab: for blocks of the large-solve regime whose slabs keep more
    than 3 n_bands rows each (the solver's own size requirement); small blocks stay replicated."""
    comm = getattr(A.basis, "comm_slab", None)
    if comm is None:
        return False
    nb, n_pw = X0.shape
    return nb > 32 and n_pw // comm.nranks > 3 * nb


def _lobpcg_hyper_batched(blocks, X0s, *, prec=True, tol=None, maxiter=100, miniter=1, n_conv_check=None):
    """The same solver for a list of independent Hamiltonian blocks in ONE library call (dftk_b200_lobpcg_multi): the
    k-blocks of a rank advance in lockstep and share kernel launches and host synchronisations."""
    from .device import lobpcg_multi
    if tol is None:
        tol = 20 * blocks[0].shape[1] * np.finfo(float).eps
    if blocks and _use_slabs(blocks[0], X0s[0]):
        return [lobpcg_hyper(b, x, prec=prec, tol=tol, maxiter=maxiter, miniter=miniter, n_conv_check=n_conv_check)
                for b, x in zip(blocks, X0s)]
    return lobpcg_multi([b.bind() for b in blocks], X0s, tol=tol, miniter=miniter, maxiter=maxiter,
                        n_conv_check=n_conv_check, prec=bool(prec))


lobpcg_hyper.batched = _lobpcg_hyper_batched


def _start_vectors(g, nev, n_Gk, generator):
    """Guess selection of diag.jl:22-38: truncate, take, or pad with orthogonalised random vectors."""
    if g.shape[1] != n_Gk:
        raise ValueError(f"Mismatch in dimension between guess ({g.shape[1]}) and Hamiltonian ({n_Gk})")
    if g.shape[0] > nev:
        return g[:nev].clone()
    if g.shape[0] == nev:
        return g.clone()
    # the solver starts with ortho!(X0) (lobpcg_hyper_impl.jl:370): the padded block only has to be full rank
    extra = torch.view_as_complex(torch.randn(nev - g.shape[0], n_Gk, 2, dtype=torch.float64, device=g.device,
                                              generator=generator))
    return torch.cat([g, extra / math.sqrt(n_Gk)], dim=0).contiguous()


def diagonalize_all_kblocks(eigensolver, ham, nev_per_kpoint, *, psiguess=None, prec_type="TPA",
                            interpolate_kpoints=True, tol=1e-6, miniter=1, maxiter=100, n_conv_check=None,
                            generator=None):
    """diag.jl:9-65.  The per-k eigenproblems are independent: an eigensolver that offers `.batched` (lobpcg_hyper does)
    gets all blocks of this rank in one call.  Without a guess the reference interpolates each k-point's start vectors
    from the solution of the previous one (diag.jl:39-44), which chains the solves; here the first block of every spin
    channel is solved on its own and its solution is interpolated to all other blocks of that channel, which are then
    solved together."""
    basis = ham.basis
    nk = len(basis.kpoints)
    for kpt in basis.kpoints:
        if kpt.n_G < nev_per_kpoint:
            raise ValueError(f"The size of the plane wave basis is {kpt.n_G}, and you are asking for "
                             f"{nev_per_kpoint} eigenvalues. Increase Ecut.")
    kw = dict(prec=prec_type is not None, tol=tol, miniter=miniter, maxiter=maxiter, n_conv_check=n_conv_check)
    batched = getattr(eigensolver, "batched", None)
    results = [None] * nk

    def solve(idx, guesses):
        if not idx:
            return
        if batched is not None:
            for i, r in zip(idx, batched([ham[i] for i in idx], guesses, **kw)):
                results[i] = r
        else:
            for i, g in zip(idx, guesses):
                results[i] = eigensolver(ham[i], g, **kw)

    if psiguess is not None:
        solve(list(range(nk)), [_start_vectors(psiguess[ik], nev_per_kpoint, basis.kpoints[ik].n_G, generator)
                                for ik in range(nk)])
    elif not interpolate_kpoints or batched is None:
        for ik, kpt in enumerate(basis.kpoints):      # the reference's chain, one block after the other
            if interpolate_kpoints and ik > 0 and basis.kpoints[ik - 1].spin == kpt.spin:
                X0 = interpolate_kpoint(results[ik - 1]["X"], basis, basis.kpoints[ik - 1], kpt)
            else:
                X0 = random_orbitals(basis, kpt, nev_per_kpoint, generator)
            solve([ik], [X0])
    else:
        from .device import random_orbitals_multi
        heads = [ik for ik, kpt in enumerate(basis.kpoints) if ik == 0 or basis.kpoints[ik - 1].spin != kpt.spin]
        solve(heads, random_orbitals_multi([basis.kblocks[ik] for ik in heads], nev_per_kpoint, _draw_seed(generator)))
        rest, guesses = [], []
        for ik, kpt in enumerate(basis.kpoints):
            if ik in heads:
                continue
            h = max(i for i in heads if i <= ik)
            rest.append(ik)
            guesses.append(interpolate_kpoint(results[h]["X"], basis, basis.kpoints[h], kpt))
        solve(rest, guesses)
    return dict(λ=[r["λ"] for r in results], X=[r["X"] for r in results],
                residual_norms=[r["residual_norms"] for r in results], n_iter=[r["n_iter"] for r in results],
                converged=all(r["converged"] for r in results), n_matvec=sum(r["n_matvec"] for r in results))


def interpolate_kpoint(X_in, basis, kpt_in, kpt_out):
    """interpolation.jl:96-115: copy the coefficients of common G vectors, orthonormalise (QR)."""
    if kpt_in is kpt_out:
        return X_in.clone()
    N = basis.N
    dev = X_in.device
    lookup = torch.full((N,), -1, dtype=torch.int64, device=dev)
    lookup[kpt_in.mapping] = torch.arange(kpt_in.n_G, device=dev)
    src = lookup[kpt_out.mapping]
    out = torch.zeros((X_in.shape[0], kpt_out.n_G), dtype=X_in.dtype, device=dev)
    ok = src >= 0
    out[:, ok] = X_in[:, src[ok]]
    # the reference orthonormalises here (ortho_qr); our eigensolver starts with ortho!(X0) anyway (same span)
    return out
