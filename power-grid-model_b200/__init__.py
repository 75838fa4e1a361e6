"""pgm_b200 -- B200-native batch power-flow engine behind the power-grid-model calculation interface."""
from . import structs  # noqa: F401
from ._lib import (FLAG_RESIDENT_INPUT, FLAG_RESIDENT_OUTPUT, TAP_STRATEGIES, BatchError, PgmB200Error, lib,  # noqa: F401
                   pinned_empty)
from .engine import Engine  # noqa: F401
from .fictional_grid import BENCHMARK_OPTION, FictionalGrid  # noqa: F401
from .model import PowerGridModel  # noqa: F401
from . import distributed  # noqa: F401,E402
from . import pgm_core  # noqa: F401,E402
