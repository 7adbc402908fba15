"""greb_b200 — host-side Python mirror of the GREB time-stepping core on B200.

The product is the CUDA library ``libgreb_b200.so`` (C ABI in ``include/greb_b200.h``); this
package is the thin ctypes layer over it plus the reference's host-side formats (namelist,
input/output files) and the synthetic forcing generator.  There is no CPU implementation here:
everything numerical goes through the C ABI and needs a B200.
"""
from . import bigrid, campaign, host, sharding, synth  # noqa: F401
from .lib import (Ensemble, GrebError, Physics, build_library, default_physics, library_path,  # noqa: F401
                  load_library, original_physics, pad_co2)

__all__ = ["Ensemble", "GrebError", "Physics", "build_library", "default_physics", "original_physics",
           "library_path", "load_library", "pad_co2", "synth", "host", "sharding", "campaign", "bigrid"]
