"""Public Python surface of the package (host-side plumbing over the C ABI in include/h1ilqr.h)."""
from .config import Config, CostWeights, MpcParams, dump_config_yaml, load_config_from_file  # noqa: F401
from .ctypes_defs import NQ, NU, NV, NX, H1Model, H1SolverOptions, H1StageTimes, H1Weights  # noqa: F401
