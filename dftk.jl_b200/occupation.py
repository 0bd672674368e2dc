"""Fermi level and occupations (mirror of src/occupation.jl).  The reference performs one scalar
allreduce per bisection step (occupation.jl:23-27); here all eigenvalues are allgathered once and the
bisection runs redundantly on every rank, which yields a bit-identical Fermi level everywhere."""
import numpy as np
from .terms import smearing_occupation


def gather_eigenvalues(basis, eigenvalues, stats=()):
    """ONE fixed-size allgather per SCF step: the eigenvalues of every rank's blocks (all blocks carry the same number
    of bands) with a few per-rank solver statistics behind them.  Returns (eigenvalues of all blocks in global block
    order b = ik + spin * n_kpt, their k-weights, stats[rank, :])."""
    comm = basis.comm_kpts
    layout = getattr(basis, "layout", None)
    stats = np.asarray(stats, dtype=np.float64).reshape(-1)
    if comm.nranks == 1:
        w = list(layout.weights) if layout is not None else list(basis.kweights)
        return [np.asarray(e, dtype=np.float64) for e in eigenvalues], w, stats[None, :]
    nb = len(eigenvalues[0])
    if any(len(e) != nb for e in eigenvalues):
        raise ValueError("all blocks must carry the same number of bands")
    buf = np.zeros(layout.max_local * nb + len(stats))
    for j, e in enumerate(eigenvalues):
        buf[j * nb:(j + 1) * nb] = e
    buf[layout.max_local * nb:] = stats
    got = comm.allgather(buf)
    ev = [None] * layout.n_blocks
    for r in range(comm.nranks):
        for j, b in enumerate(layout.blocks_of_rank[r]):
            ev[b] = got[r, j * nb:(j + 1) * nb].copy()
    return ev, list(layout.weights), got[:, layout.max_local * nb:]


def _occ(model, eigs, eF):
    if model.temperature == 0:
        return [model.filled_occupation * smearing_occupation("None", e - eF) for e in eigs]
    return [model.filled_occupation * smearing_occupation(model.smearing, (e - eF) / model.temperature) for e in eigs]


def compute_occupation(basis, eigenvalues, *, tol_n_elec=1e-6, gathered=None, return_global=False):
    """Returns (occupation of the local blocks, εF).  `gathered` = (eigenvalues of all blocks, weights) when the caller
    already did the allgather (next_density packs solver statistics into the same collective)."""
    model = basis.model
    for ek in eigenvalues:
        if not np.all(np.diff(ek) >= -np.finfo(float).eps):
            raise ValueError("Eigenvalues should be monotonically increasing.")
    ev, w = gathered if gathered is not None else gather_eigenvalues(basis, eigenvalues)[:2]
    filled = model.filled_occupation
    if model.n_electrons == 0:
        eF = min(e.min() for e in ev) - 1.0
        occ = [np.zeros(len(e)) for e in eigenvalues]
        return (occ, eF, [np.zeros(len(e)) for e in ev]) if return_global else (occ, eF)

    # all eigenvalues of all blocks as ONE array with their k-weights: an evaluation of the electron count is one vectorised
    # pass (the bisection needs ~60 of them; a per-block Python loop over 84 blocks made this a third of a metal's SCF step)
    e_all = np.concatenate([np.asarray(e, dtype=np.float64) for e in ev])
    w_all = np.concatenate([np.full(len(e), float(wk)) for wk, e in zip(w, ev)])

    def excess(eF):
        return float(np.dot(w_all, _occ(model, [e_all], eF)[0])) - model.n_electrons

    if filled * sum(wk * len(e) for wk, e in zip(w, ev)) < model.n_electrons - tol_n_elec:
        raise RuntimeError("Could not obtain required number of electrons by filling every state. Increase n_bands.")
    n_fill = -(-model.n_electrons // (model.n_spin_components * filled))
    HOMO = max(e[n_fill - 1] for e in ev)
    lum = [e[n_fill:].min() for e in ev if len(e) > n_fill]
    eF = (HOMO + min(lum)) / 2 if lum else HOMO + 1
    if model.temperature == 0:
        if model.n_electrons % (model.n_spin_components * filled) != 0:
            raise RuntimeError(f"{model.n_electrons} electrons cannot be attained by filling states with "
                               f"occupation {filled}; add a temperature or use collinear spin")
        if abs(excess(eF)) > tol_n_elec:
            raise RuntimeError("Unable to find non-fractional occupations that have the correct number of "
                               "electrons. You should add a temperature.")
    else:
        ex = excess(eF)
        if abs(ex) >= tol_n_elec / 10:
            lo, hi = (eF, max(e.max() for e in ev) + 1) if ex < 0 else (min(e.min() for e in ev) - 1, eF)
            if not (excess(lo) <= 0 <= excess(hi)):      # occupation.jl:100-103 (@assert on the bracket)
                raise RuntimeError("compute_occupation: the Fermi level is not bracketed by the eigenvalue range")
            for _ in range(200):      # Roots.Bisection to atol = eps
                mid = (lo + hi) / 2
                if mid == lo or mid == hi:
                    break
                if excess(mid) < 0:
                    lo = mid
                else:
                    hi = mid
            eF = (lo + hi) / 2
            if abs(excess(eF)) > tol_n_elec:
                import warnings
                warnings.warn("Large deviation of electron count in compute_occupation.")
    occ = _occ(model, [np.asarray(e) for e in eigenvalues], eF)
    return (occ, eF, _occ(model, ev, eF)) if return_global else (occ, eF)
