"""LOBPCG restatement (oracle; test infrastructure only).

Follows src/eigen/lobpcg_hyper_impl.jl (all of it, B = I as always passed by
src/eigen/diag_lobpcg_hyper.jl:11) and src/eigen/preconditioners.jl:27-78.  Julia views over the
active block are expressed as absolute column offsets `a0:` into the full N×M arrays.
"""
import numpy as np

EPS = np.finfo(float).eps


def columnwise_norms(X):
    return np.sqrt(np.sum(np.abs(X) ** 2, axis=0))


def normest(M):
    d = np.diag(M)
    return np.max(np.abs(d)) + np.linalg.norm(M - np.diag(d))


def safe_cholesky(O, nchol=0, alpha=100.0):
    """:190-210.  Returns (R, invR, nchol)."""
    if nchol >= 5:
        return None, None, 10000
    nchol += 1
    try:
        L = np.linalg.cholesky(O)           # O = L L^H, R = L^H
        R = L.conj().T
        invR = np.linalg.inv(R)
        if np.any(np.isnan(invR)):
            raise np.linalg.LinAlgError("nan")
    except np.linalg.LinAlgError:
        O = O + alpha * EPS * np.linalg.norm(O) * np.eye(O.shape[0])
        return safe_cholesky(O, nchol, alpha * 10)
    return R, invR, nchol


def ortho(X, tol=2 * EPS, stats=None):
    """ortho!(X) :216-261.  Returns (X, nchol_total, growth_factor)."""
    growth = 1.0
    nchol_total = 0
    while True:
        O = X.conj().T @ X
        O = np.triu(O) + np.triu(O, 1).conj().T   # Hermitian(upper)
        R, invR, nchol = safe_cholesky(O)
        nchol_total += nchol
        if nchol > 10:
            U, _, Vh = np.linalg.svd(X, full_matrices=False)
            return U @ Vh, 100, 1.0
        X = X @ invR
        norminvR = normest(invR)
        growth *= norminvR
        condR = normest(R) * norminvR
        est = EPS * condR ** 2
        if stats is not None:
            stats["n_chol"] = stats.get("n_chol", 0) + nchol
        if nchol == 1 and est < tol:
            break
    return X, nchol_total, growth


def drop_small(X, tol, rng):
    dropped = np.nonzero(columnwise_norms(X) <= tol)[0]
    if len(dropped):
        X[:, dropped] = (rng.standard_normal((X.shape[0], len(dropped)))
                         + 1j * rng.standard_normal((X.shape[0], len(dropped)))) / np.sqrt(2)
    return dropped


def ortho_against(X, Y, tol=2 * EPS, rng=None, stats=None):
    """ortho!(X, Y, BY) with BY = Y, :271-323."""
    rng = rng or np.random.default_rng(0)
    X = X / columnwise_norms(X)[None, :]
    niter = 1
    while True:
        BYX = Y.conj().T @ X
        X = X - Y @ BYX
        dropped = drop_small(X, tol, rng)
        if len(dropped):
            X[:, dropped] -= Y @ (Y.conj().T @ X[:, dropped])
        if np.linalg.norm(BYX) < tol and niter > 1:
            break
        X, _ninner, growth = ortho(X, tol, stats)
        if growth * EPS < tol:
            break
        if niter > 10:
            U, _, Vh = np.linalg.svd(X, full_matrices=False)
            return U @ Vh
        niter += 1
    return X


def rayleigh_ritz(Y, AY, N):
    """:141-171 (Julia >= 1.12 path: eigen(Hermitian; alg=DivideAndConquer))."""
    XAX = Y.conj().T @ AY
    assert not np.any(np.isnan(np.triu(XAX)))
    w, v = np.linalg.eigh(XAX, UPLO="U")
    return v[:, :N], w[:N]


class PreconditionerTPA:
    """preconditioners.jl:27-78."""

    def __init__(self, kin, default_shift=1.0):
        self.kin = kin
        self.mean_kin = None
        self.default_shift = default_shift

    def prep(self, X):
        self.mean_kin = np.real(np.sum(np.conj(X) * (self.kin[:, None] * X), axis=0))

    def ldiv(self, R):
        if self.mean_kin is None:
            return R / (self.kin + self.default_shift)[:, None]
        mk = self.mean_kin[None, :]
        return (mk / (mk + self.kin[:, None])) * R


def lobpcg(A, X0, precon=None, tol=1e-10, maxiter=100, miniter=1, ortho_tol=2 * EPS,
           n_conv_check=None, rng=None, stats=None):
    """LOBPCG :354-582 + lobpcg_hyper wrapper diag_lobpcg_hyper.jl:5-18.

    A: object with `matmul(X)`;  returns dict(λ, X, residual_norms, n_iter, converged, n_matvec).
    """
    rng = rng or np.random.default_rng(1234)
    N, M = X0.shape
    assert N > 3 * M, "The eigenproblem is too small"
    if n_conv_check is None:
        n_conv_check = M
    resid_history = np.zeros((M, maxiter + 1))

    X, _, _ = ortho(X0.astype(complex).copy(), ortho_tol, stats)
    n_matvec = M
    AX = A.matmul(X)
    assert not np.any(np.isnan(AX))
    P, AP, R, AR = (np.zeros_like(X) for _ in range(4))
    new_R = np.zeros_like(X)
    new_X, new_AX = X.copy(), AX.copy()
    new_P, new_AP = np.zeros_like(X), np.zeros_like(X)
    nlocked = 0
    niter = 0
    lam = np.real(np.sum(np.conj(X) * AX, axis=0) / np.sum(np.conj(X) * X, axis=0))
    a0 = 0
    cX = None
    Y = AY = None

    def result(niter_final):
        lam_out, Xo, res = lam.copy(), X, resid_history
        if not np.all(np.diff(lam_out) >= 0):
            p = np.argsort(lam_out, kind="stable")
            lam_out, Xo, res = lam_out[p], X[:, p], resid_history[p, :]
        rn = res[:, niter_final]
        return dict(λ=lam_out, X=Xo, residual_norms=rn, n_iter=niter_final, n_matvec=n_matvec,
                    converged=bool(np.max(rn[:n_conv_check]) < tol),
                    residual_history=res[:, :niter_final + 1])

    while True:
        Ma = M - a0
        if niter > 0:
            AR[:, a0:] = A.matmul(R[:, a0:])
            n_matvec += Ma
            if niter > 1:
                Y = np.concatenate([X[:, a0:], R[:, a0:], P[:, a0:]], axis=1)
                AY = np.concatenate([AX[:, a0:], AR[:, a0:], AP[:, a0:]], axis=1)
            else:
                Y = np.concatenate([X[:, a0:], R[:, a0:]], axis=1)
                AY = np.concatenate([AX[:, a0:], AR[:, a0:]], axis=1)
            cX, lamRR = rayleigh_ritz(Y, AY, M - nlocked)
            lam[a0:] = lamRR
            new_X[:, a0:] = Y @ cX
            new_AX[:, a0:] = AY @ cX

        new_R[:, a0:] = new_AX[:, a0:] - new_X[:, a0:] * lam[None, a0:]
        resid_history[a0:, niter] = columnwise_norms(new_R[:, a0:])

        if precon is not None:
            precon.prep(new_X[:, a0:])
            new_R[:, a0:] = precon.ldiv(new_R[:, a0:])

        prev_nlocked = nlocked
        if niter >= miniter:
            for i in range(nlocked, M):
                if resid_history[i, niter] < tol:
                    nlocked += 1
                else:
                    break

        if nlocked >= n_conv_check:
            X[:, a0:] = new_X[:, a0:]
            AX[:, a0:] = new_AX[:, a0:]
            return result(niter)
        newly_locked = nlocked - prev_nlocked

        if niter > 0:
            lenXn = Ma - newly_locked
            e = np.zeros((cX.shape[0], Ma))
            for c in range(lenXn):
                e[newly_locked + c, c] = 1.0
            cP = (cX - e)[:, newly_locked:Ma]
            cP = ortho_against(cP, cX, ortho_tol, rng, stats)
            new_P[:, a0 + newly_locked:] = Y @ cP
            new_AP[:, a0 + newly_locked:] = AY @ cP

        X[:, a0:] = new_X[:, a0:]
        AX[:, a0:] = new_AX[:, a0:]
        R[:, a0:] = new_R[:, a0:]

        diffs = np.abs(np.sum(np.conj(X[:, a0:]) * X[:, a0:], axis=0) - 1)
        if np.any(diffs >= np.sqrt(EPS)):
            raise RuntimeError("LOBPCG is badly failing to keep the vectors normalized")

        a0 += newly_locked
        if niter > 0:
            P[:, a0:] = new_P[:, a0:]
            AP[:, a0:] = new_AP[:, a0:]
            Z = np.concatenate([X, P[:, a0:]], axis=1)
        else:
            Z = X
        R[:, a0:] = ortho_against(R[:, a0:], Z, ortho_tol, rng, stats)

        if niter >= maxiter:
            break
        niter += 1

    return result(maxiter)
