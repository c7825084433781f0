"""oracle/reference_harness.py -- TEST INFRASTRUCTURE, build-container only.

Imports the UNMODIFIED reference /root/reference/model.py.  model.py:26 does
`from plot import plot_model_results`, and plot.py imports plotly, which is not
installed in this image; a stub module named `plot` is inserted into sys.modules
first.  Nothing under /root/reference is edited or copied.

/root/reference does not exist on the GPU box, so nothing that runs there
(`-m gpu` tests, smoke(), bench.py) may import this module; it is used by
oracle/make_golden.py and by CPU-side tests that skip when the tree is absent.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types
import warnings

REFERENCE_DIR = os.environ.get("VAX_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "model.py"))


def load_reference():
    """Return the reference `model` module (imported once)."""
    if "plot" not in sys.modules or not hasattr(sys.modules["plot"], "plot_model_results"):
        stub = types.ModuleType("plot")
        stub.plot_model_results = lambda *a, **k: None
        sys.modules["plot"] = stub
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    warnings.filterwarnings("ignore", category=FutureWarning)
    warnings.filterwarnings("ignore", message=".*'d' is deprecated.*")
    import model  # noqa: E402  (the reference's model.py)

    assert os.path.abspath(model.__file__).startswith(os.path.abspath(REFERENCE_DIR))
    return model


@contextlib.contextmanager
def expected_value_mode():
    """north_star check (1): 'RNG replaced by means' -> np.random.poisson := identity."""
    import numpy as np

    orig = np.random.poisson
    np.random.poisson = lambda lam: lam
    try:
        yield
    finally:
        np.random.poisson = orig
