import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")

USA_BOUNDS_PCT = dict(pro=[35, 45], anti=[12, 25], pressure=[0.0, 0.02], tau=[5, 7], nv0=[1, 1.2],
                      nvmax=[5, 10])
USA_DATES = dict(start_date="2020-12-15", end_date="2022-07-1")
WIDE_BOUNDS_PCT = dict(pro=[30, 50], anti=[5.5, 31.5], pressure=[0.0, 0.1], tau=[4, 8],
                       nv0=[0.9, 1.3], nvmax=[2.5, 12.5])


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def vax():
    """The product package (directory name has hyphens -> importlib)."""
    return importlib.import_module("covid19-vaccination-model_b200")


@pytest.fixture(scope="session")
def ctx(vax):
    """A libvaxmc context on cuda:0; GPU tests only."""
    return vax.default_context(0)


def usa_bounds_frac():
    import numpy as np

    b = USA_BOUNDS_PCT
    return np.concatenate([np.array(b["pro"]) / 100, np.array(b["anti"]) / 100,
                           np.array(b["pressure"], dtype=float), np.array(b["tau"], dtype=float),
                           np.array(b["nv0"]) / 100, np.array(b["nvmax"]) / 100])


def wide_bounds_frac():
    import numpy as np

    b = WIDE_BOUNDS_PCT
    return np.concatenate([np.array(b["pro"]) / 100, np.array(b["anti"]) / 100,
                           np.array(b["pressure"], dtype=float), np.array(b["tau"], dtype=float),
                           np.array(b["nv0"]) / 100, np.array(b["nvmax"]) / 100])
