"""Import shim: the package directory is named ``whisper-burn_b200`` (after the reference), which
is not a valid Python identifier.  ``import wb200`` registers it as ``whisper_burn_b200``."""
import importlib.util
import sys
from pathlib import Path

_ROOT = Path(__file__).resolve().parent
_PKG = _ROOT / "whisper-burn_b200"

if "whisper_burn_b200" not in sys.modules:
    _spec = importlib.util.spec_from_file_location("whisper_burn_b200", _PKG / "__init__.py",
                                                   submodule_search_locations=[str(_PKG)])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules["whisper_burn_b200"] = _mod
    _spec.loader.exec_module(_mod)

pkg = sys.modules["whisper_burn_b200"]
