"""Shared silicon fixture (reference: test/testcases.jl:12-29)."""
import numpy as np

A = 5.131570667152971
LATTICE = np.array([[0, A, A], [A, 0, A], [A, A, 0]])
POSITIONS = [np.ones(3) / 8, -np.ones(3) / 8]
KCOORDS = [[0, 0, 0], [1 / 3, 0, 0], [1 / 3, 1 / 3, 0], [-1 / 3, 1 / 3, 0]]
KWEIGHTS = [1 / 27, 8 / 27, 6 / 27, 12 / 27]
