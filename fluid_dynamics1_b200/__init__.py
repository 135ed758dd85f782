"""cnavier hot path on B200 (sm_100a): streamfunction Poisson solve + explicit vorticity-transport
stencils behind the reference's own C signatures.  See DESIGN.md / INTEGRATION.md."""
from ._lib import Config, dropin, lib, require_gpu  # noqa: F401
from .config import config_from_dict, load_config_from_file, load_default_config, print_config  # noqa: F401
from .solver import (PoissonNotConverged, PoissonSolver, Simulation, apply_operator, continuity, diff_matrix,  # noqa: F401
                     error, euler, host_empty, num_steps, poisson_sor, pressure_rhs, sor_beta, vorticity)
