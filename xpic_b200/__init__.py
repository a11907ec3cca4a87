"""xpic_b200 -- B200-native implementation of xpic's implicit energy-conserving PIC step.

The product is the C-ABI shared library built from xpic_b200/csrc (include/xpic_b200.h);
this package is the thin ctypes binding the tests and bench.py drive it through, plus the
build helper.  There is no CPU fallback: importing works anywhere, creating a Simulation
needs the built library and a CUDA device.
"""
from .binding import (ECCAPFIM, ECSIM, ECSIMCORR, FIELDS, SCALARS, STAGES, Simulation, XpicB200Error, build_library, coef_table,
                      comm_unique_id, library_path, load_library, owner_rank, slab_range)

__all__ = [
    "ECCAPFIM", "ECSIM", "ECSIMCORR", "FIELDS", "SCALARS", "STAGES", "Simulation", "XpicB200Error", "build_library", "coef_table",
    "comm_unique_id", "library_path", "load_library", "owner_rank", "slab_range",
]
