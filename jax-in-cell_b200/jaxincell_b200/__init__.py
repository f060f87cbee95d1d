"""jaxincell_b200 -- B200-native hot path (gather, Boris push, boundaries, deposition, filter, Maxwell update) behind
JAX-in-Cell's Simulation API.  The arithmetic lives in libjic_b200.so (hand-written CUDA for sm_100a, C ABI in
include/jic_b200.h); this package is the Python host side.  There is no CPU fallback."""
from ._lib import JicError, LIB_PATH, load  # noqa: F401
from ._engine import HotPath, make_params, make_species, sample_particles, simulate_host, trim_memory  # noqa: F401
from ._parallel import shard_counts, shard_particles, shard_species  # noqa: F401
from ._simulation import Simulation, diagnostics, load_parameters, simulation  # noqa: F401
from ._algorithms import Boris_step, CN_step  # noqa: F401
