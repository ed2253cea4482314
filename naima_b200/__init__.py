# -*- coding: utf-8 -*-
"""naima_b200 -- B200-native implementation of naima's radiative-model likelihood
hot path (radiative models + Gaussian likelihood + ensemble sampling), behind the
reference's Python API.  See DESIGN.md / INTEGRATION.md.

Importing the package does not touch the GPU; the first computation loads
``libnaima_b200.so`` (built by ``python -m naima_b200.build``) and needs a CUDA
device -- there is no CPU fallback.
"""
from . import units  # noqa: F401
from . import models  # noqa: F401
from .analysis import find_ML, model_samples, read_run, save_run  # noqa: F401
from .core import (get_sampler, lnprob, lnprobmodel, log_uniform_prior, normal_prior,  # noqa: F401
                   run_sampler, uniform_prior)
from .fused import LikelihoodPlan, TraceError  # noqa: F401
from .sampler import DeviceEnsemble, EnsembleSampler, PlanSampler, State  # noqa: F401
from .utils import (DataTable, build_data_table, generate_energy_edges, read_ipac,  # noqa: F401
                    sed_conversion, trapz_loglog, validate_data_table)

__version__ = "0.1.0"
