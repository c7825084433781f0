"""B200-native Monte Carlo hot path of dfcastellanos/COVID19-Vaccination-Model.

The directory name contains hyphens, so import it with importlib (or through the
`vaxmc_b200` alias module at the repo root):

    import importlib
    vax = importlib.import_module("covid19-vaccination-model_b200")
    results, err, msg = vax.run_model([35, 45], [12, 25], [0., 0.02], [5, 7], [1, 1.2], [5, 10],
                                      95, 1_000_000, 50000,
                                      dict(start_date="2020-12-15", end_date="2022-07-1"))

Contents: csrc/ (CUDA kernels + C ABI, built in-tree to libvaxmc.so), _native (ctypes),
model (the reference's run_model / CLI interface), distributed (sample sharding + NCCL merge),
country_data (offline-safe get_country_data for the app's overlay).
"""
from . import _native, build, country_data, distributed, model  # noqa: F401
from .country_data import get_country_data  # noqa: F401
from ._native import Context, VaxmcError, default_context  # noqa: F401
from .model import (  # noqa: F401
    SERIES, cache_clear, main, results_to_jsonable, run_model, run_model_grid, run_model_quantiles, run_sampling,
    run_single_realization, sample_param_combinations, seed,
)

__all__ = [
    "run_model", "run_model_grid", "run_model_quantiles", "run_sampling", "sample_param_combinations", "run_single_realization", "seed",
    "main", "SERIES", "Context", "VaxmcError", "default_context", "cache_clear",
    "results_to_jsonable", "get_country_data",
]
