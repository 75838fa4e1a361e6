"""pgm_b200 -- B200-native batch power-flow engine behind the power-grid-model calculation interface."""
from . import structs  # noqa: F401
