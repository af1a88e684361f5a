"""fft-isdf-scratch_b200: B200-native FFT-ISDF build (point selection, Theta fit, V_{mu nu}(k)).

Import name: `fft_isdf_scratch_b200` (see the loader shim at the repo root; the directory
name carries a hyphen).  Public surface mirrors /root/reference/fftisdf.py.
"""
from . import pbc_tools  # noqa: F401
from .cell import SyntheticCell, TableCell, random_cubic_cell, diamond_standin, nio_afm_standin  # noqa: F401

__all__ = ["pbc_tools", "SyntheticCell", "TableCell", "random_cubic_cell", "diamond_standin", "nio_afm_standin"]
