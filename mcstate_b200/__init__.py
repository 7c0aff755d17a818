"""mcstate_b200: B200-native particle-filter engine behind mcstate's model-object API.

The CUDA engine lives in ``libmcstate_b200.so`` (sources in ``csrc/``, C ABI in
``include/mcstate_b200.h``).  The Python modules mirror the reference's host
interface for the hot path only: the dust model object (``dust``), the data
helper (``particle_filter_data``) and the filter facade (``particle_filter``).  (``mcstate_b200.dust`` is the module; its
``dust(path)`` function -- user-supplied models -- is reached as ``dust.dust``.)
"""
from . import _lib
from ._rng import set_seed
from .dust import (dust_data, dust_example, dust_generator, dust_model,
                   dust_openmp_threads, dust_repair_environment, dust_rng,
                   dust_rng_distributed_state)

__all__ = ["_lib", "dust_data", "dust_example", "dust_generator", "dust_model",
           "dust_openmp_threads", "dust_repair_environment", "dust_rng",
           "dust_rng_distributed_state", "set_seed"]
