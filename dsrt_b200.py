"""Import shim: the package directory is named `deep-space-ray-tracer_b200` (not a valid
Python identifier), so `import dsrt_b200` loads it from its path."""
import importlib.util
import sys
from pathlib import Path

_pkg = Path(__file__).resolve().parent / "deep-space-ray-tracer_b200"
_spec = importlib.util.spec_from_file_location(
    "dsrt_b200", _pkg / "__init__.py", submodule_search_locations=[str(_pkg)]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
