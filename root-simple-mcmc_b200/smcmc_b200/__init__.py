"""smcmc_b200 -- Python binding of libsmcmc_b200.so (include/smcmc_b200.h).

The binding is ctypes over the C ABI; it adds no computation of its own.  The
hot path lives in the CUDA library, and importing :class:`Engine` without the
built library, or creating one without a GPU, fails loudly: there is no CPU
fallback.
"""
from .binding import (Engine, EVENT_DTYPE, LLH_ASYM, LLH_DUMMY, LLH_FAKE,
                      LLH_FAKE2, LLH_HARD, LLH_HORRIFIC, LLH_UNBINNED, LLH_UNIT_GAUSS, LLH_USER, SmcmcError,
                      build_library, library_path, load_library)
from .binding import PROPOSAL_ADAPTIVE, PROPOSAL_VAAT
from . import shard, synth

__all__ = ["Engine", "PROPOSAL_ADAPTIVE", "PROPOSAL_VAAT", "EVENT_DTYPE", "SmcmcError", "build_library",
           "library_path", "load_library", "shard", "synth", "LLH_UNIT_GAUSS",
           "LLH_DUMMY", "LLH_HORRIFIC", "LLH_ASYM", "LLH_FAKE", "LLH_UNBINNED", "LLH_HARD", "LLH_FAKE2", "LLH_USER"]
