"""dftk_b200: B200-native plane-wave Kohn-Sham SCF hot path behind DFTK.jl's operator API.

The package directory is `dftk.jl_b200/`; import it as `dftk_b200` (see dftk_b200.py at the repo root).
"""
from . import _lib
from ._lib import DftkB200Error, LIB_PATH
from .device import Context, FFTGrid, KBlock

__all__ = ["Context", "FFTGrid", "KBlock", "DftkB200Error", "LIB_PATH"]
