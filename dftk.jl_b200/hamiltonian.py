"""Hamiltonian / DftHamiltonianBlock / energy_hamiltonian (host-side mirror of
src/terms/Hamiltonian.jl:22-57,137-236).  `mul!(Hψ, H::DftHamiltonianBlock, ψ)` is one C-ABI call
(dftk_b200_apply_h) that runs the batched FFT pipeline + fused kinetic/local scalings + nonlocal GEMMs."""
import math
import torch

from .terms import (RealSpaceMultiplication, FourierMultiplication, NonlocalOperator, NoopOperator)


class Energies(dict):
    """src/Energies.jl: named energy container with a `.total`."""

    @property
    def total(self):
        return sum(self.values())


class DftHamiltonianBlock:
    def __init__(self, basis, ik, operators, pot_cache=None):
        self.basis, self.ik, self.kpoint = basis, ik, basis.kpoints[ik]
        self.operators = operators
        ops = [o for o in operators if not isinstance(o, NoopOperator)]
        fourier = [o for o in ops if isinstance(o, FourierMultiplication)]
        real = [o for o in ops if isinstance(o, RealSpaceMultiplication)]
        nonloc = [o for o in ops if isinstance(o, NonlocalOperator)]
        if len(fourier) > 1 or len(nonloc) > 1 or len(fourier) + len(real) + len(nonloc) != len(ops):
            raise NotImplementedError("only DFT Hamiltonians (one Fourier multiplication, local potentials, "
                                      "at most one nonlocal operator) are supported by the B200 back end")
        self.fourier_op = fourier[0] if fourier else None
        self.nonlocal_op = nonloc[0] if nonloc else None
        # optimize_operators (operators.jl:213-222): sum all real-space multiplications
        self.local_op = None
        if real:
            # all k-blocks of one spin share the same term potentials: sum them once (keyed by the storage they view)
            key = tuple((o.potential.data_ptr(), o.potential.numel()) for o in real)
            pot = pot_cache.get(key) if pot_cache is not None else None
            if pot is None:
                pot = real[0].potential
                for o in real[1:]:
                    pot = pot + o.potential
                pot = pot.contiguous()
                if pot_cache is not None:
                    pot_cache[key] = pot
            self.local_op = RealSpaceMultiplication(basis, self.kpoint, pot)
        self.kblock = basis.kblocks[ik]
        self.kblock.set_potential(self.local_op.potential if self.local_op is not None else None)

    @property
    def shape(self):
        return (self.kpoint.n_G, self.kpoint.n_G)

    def mul(self, psi, out=None):
        """Hψ for a block of bands; psi: (n_bands, n_G) complex128 on the device."""
        return self.kblock.apply_h(psi, out)

    __matmul__ = mul


class Hamiltonian:
    def __init__(self, basis, blocks):
        self.basis, self.blocks = basis, blocks

    def __getitem__(self, ik):
        return self.blocks[ik]

    def mul(self, psi):
        return [blk.mul(p) for blk, p in zip(self.blocks, psi)]

    __matmul__ = mul


def energy_hamiltonian(basis, psi, occupation, *, rho, eigenvalues=None, eF=None, **kw):
    """Hamiltonian.jl:200-227: energies of every term + the per-k Hamiltonian blocks."""
    energies, per_term_ops = Energies(), []
    basis._be_cache = {}
    try:
        for name, term in zip(basis.model.term_types, basis.terms):
            E, ops = term.ene_ops(basis, psi, occupation, rho=rho, eigenvalues=eigenvalues, eF=eF)
            energies[name] = E
            per_term_ops.append(ops)
    finally:
        basis._be_cache = None
    pot_cache = {}
    blocks = [DftHamiltonianBlock(basis, ik, [ops[ik] for ops in per_term_ops], pot_cache)
              for ik in range(len(basis.kpoints))]
    return energies, Hamiltonian(basis, blocks)


def energy(basis, psi, occupation, *, rho, eigenvalues=None, eF=None, **kw):
    """Hamiltonian.jl:232-236 (energies only)."""
    energies = Energies()
    basis._be_cache = {}
    try:
        for name, term in zip(basis.model.term_types, basis.terms):
            energies[name] = term.ene_ops(basis, psi, occupation, rho=rho, eigenvalues=eigenvalues, eF=eF)[0]
    finally:
        basis._be_cache = None
    return energies
