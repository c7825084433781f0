"""Importable alias of the package directory `covid19-vaccination-model_b200/` (hyphens are
not legal in an `import` statement).  `import vaxmc_b200 as vax; vax.run_model(...)`."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("covid19-vaccination-model_b200")
globals().update({k: getattr(_pkg, k) for k in dir(_pkg) if not k.startswith("__")})
package = _pkg
