"""The host-side generator behind every draw the reference takes from R's global RNG
(proposals and acceptance in pmcmc, R/pmcmc_state.R:101-138; sampling from results,
R/pmcmc_tools.R:62-90; smc2, R/smc2.R:120-189): ``set_seed(n)`` is the counterpart of R's
``set.seed(n)``, so a default run is reproducible.  The particle streams on the device are
seeded separately (``seed=`` of the filter)."""
from __future__ import annotations

import numpy as np

_host = np.random.default_rng()


def set_seed(seed=None):
    """Reseed the host generator (``None``: from OS entropy)."""
    global _host
    _host = np.random.default_rng(seed)


def host_rng():
    return _host
