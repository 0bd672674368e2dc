"""Fermi level and occupations (mirror of src/occupation.jl).  The reference performs one scalar
allreduce per bisection step (occupation.jl:23-27); here all eigenvalues are allgathered once and the
bisection runs redundantly on every rank, which yields a bit-identical Fermi level everywhere."""
import numpy as np
from .terms import smearing_occupation


def _gather(basis, eigenvalues):
    comm = basis.comm_kpts
    if comm.nranks == 1:
        return [np.asarray(e) for e in eigenvalues], list(basis.kweights)
    pieces = comm.allgather_object(([np.asarray(e) for e in eigenvalues], list(basis.kweights)))
    ev, w = [], []
    for e, ww in pieces:
        ev += e
        w += ww
    return ev, w


def _occ(model, eigs, eF):
    if model.temperature == 0:
        return [model.filled_occupation * smearing_occupation("None", e - eF) for e in eigs]
    return [model.filled_occupation * smearing_occupation(model.smearing, (e - eF) / model.temperature) for e in eigs]


def compute_occupation(basis, eigenvalues, *, tol_n_elec=1e-6):
    """Returns (occupation of the local k-points, εF)."""
    model = basis.model
    for ek in eigenvalues:
        if not np.all(np.diff(ek) >= -np.finfo(float).eps):
            raise ValueError("Eigenvalues should be monotonically increasing.")
    ev, w = _gather(basis, eigenvalues)
    filled = model.filled_occupation

    def excess(eF):
        return sum(wk * o.sum() for wk, o in zip(w, _occ(model, ev, eF))) - model.n_electrons

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
    return _occ(model, [np.asarray(e) for e in eigenvalues], eF), eF
