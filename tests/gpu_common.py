"""Helpers for the GPU parity tests: build device k-blocks from oracle-side basis data."""
import numpy as np
import torch

import dftk_b200
from oracle.basis import Element, Model, PlaneWaveBasis
from oracle.terms import Terms, energy_hamiltonian, guess_density
from silicon import LATTICE, POSITIONS

_CTX = None


def ctx():
    global _CTX
    if _CTX is None:
        _CTX = dftk_b200.Context(0)
    return _CTX


def silicon_setup(Ecut=15, fft_size=(27, 27, 27), kcoords=((0.1, -0.2, 0.3),), kweights=(1.0,),
                  functionals=("lda_x", "lda_c_vwn"), terms=None):
    kw = {} if terms is None else dict(terms=terms)
    m = Model(LATTICE, [Element("Si")] * 2, POSITIONS, functionals=functionals, symmetries=False, **kw)
    b = PlaneWaveBasis(m, Ecut, fft_size=fft_size, kcoords=list(kcoords), kweights=list(kweights))
    t = Terms(b)
    rho = guess_density(b)
    _, ham = energy_hamiltonian(b, t, None, None, rho)
    return m, b, t, rho, ham


def device_blocks(b, ham):
    c = ctx()
    grid = dftk_b200.FFTGrid(c, b.fft_size, b.model.unit_cell_volume)
    blocks = []
    for blk in ham:
        P = D = None
        if blk.PD is not None:
            P = torch.from_numpy(np.ascontiguousarray(blk.PD[0].T)).to(c.device)   # (n_proj, n_pw)
            D = blk.PD[1]
        kb = dftk_b200.KBlock(grid, blk.kpt.mapping, kin=blk.kin, P=P, D=D, spin=blk.kpt.spin,
                              kweight=b.kweights[blk.ik])
        if blk.Vtot is not None:
            kb.set_potential(torch.from_numpy(blk.Vtot).to(c.device))
        blocks.append(kb)
    return grid, blocks


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(ctx().device)


def rand_psi(n_G, nb, seed=0):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((nb, n_G)) + 1j * rng.standard_normal((nb, n_G))
