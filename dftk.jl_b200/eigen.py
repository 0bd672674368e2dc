"""lobpcg_hyper / diagonalize_all_kblocks (mirror of src/eigen/diag.jl:9-65 and
src/eigen/diag_lobpcg_hyper.jl:5-18).  The eigensolver itself runs inside libdftk_b200
(dftk_b200_lobpcg); this file only chooses start vectors and packs results like the reference."""
import numpy as np
import torch


def random_orbitals(basis, kpt, howmany, generator=None):
    """orbitals.jl:82-87: randn + QR (on the device)."""
    dev = kpt.mapping.device
    A = torch.view_as_complex(torch.randn(kpt.n_G, howmany, 2, dtype=torch.float64, device=dev, generator=generator))
    Q, _ = torch.linalg.qr(A)
    return Q[:, :howmany].T.contiguous()      # stored (n_bands, n_G)


def lobpcg_hyper(A, X0, *, prec=True, tol=None, maxiter=100, miniter=1, n_conv_check=None):
    """Eigensolver with the reference's call signature `eigensolver(A, X0; prec, tol, maxiter, ...)`;
    returns (; λ, X, residual_norms, n_iter, converged, n_matvec).  X0 is consumed (updated in place)."""
    if tol is None:
        tol = 20 * A.shape[1] * np.finfo(float).eps
    return A.bind().lobpcg(X0, tol=tol, miniter=miniter, maxiter=maxiter, n_conv_check=n_conv_check,
                           prec=bool(prec))


def diagonalize_all_kblocks(eigensolver, ham, nev_per_kpoint, *, psiguess=None, prec_type="TPA",
                            interpolate_kpoints=True, tol=1e-6, miniter=1, maxiter=100, n_conv_check=None,
                            generator=None):
    basis = ham.basis
    results = []
    for ik, kpt in enumerate(basis.kpoints):
        n_Gk = kpt.n_G
        if n_Gk < nev_per_kpoint:
            raise ValueError(f"The size of the plane wave basis is {n_Gk}, and you are asking for "
                             f"{nev_per_kpoint} eigenvalues. Increase Ecut.")
        if psiguess is not None:
            g = psiguess[ik]
            if g.shape[1] != n_Gk:
                raise ValueError(f"Mismatch in dimension between guess ({g.shape[1]}) and Hamiltonian ({n_Gk})")
            if g.shape[0] > nev_per_kpoint:
                X0 = g[:nev_per_kpoint].clone()
            elif g.shape[0] == nev_per_kpoint:
                X0 = g.clone()
            else:
                extra = torch.view_as_complex(torch.randn(nev_per_kpoint - g.shape[0], n_Gk, 2, dtype=torch.float64,
                                                          device=g.device, generator=generator))
                Q, _ = torch.linalg.qr(torch.cat([g, extra], dim=0).T)
                X0 = Q.T.contiguous()
        elif interpolate_kpoints and ik > 0 and basis.kpoints[ik - 1].spin == kpt.spin:
            X0 = interpolate_kpoint(results[ik - 1]["X"], basis, basis.kpoints[ik - 1], kpt)
        else:
            X0 = random_orbitals(basis, kpt, nev_per_kpoint, generator)
        results.append(eigensolver(ham[ik], X0, prec=prec_type is not None, tol=tol, miniter=miniter,
                                   maxiter=maxiter, n_conv_check=n_conv_check))
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
    Q, _ = torch.linalg.qr(out.T)
    return Q.T.contiguous()
