"""Import shim: the package directory is named `collaborative-zksnark_b200` (not a valid Python
identifier), so it is loaded by path and exposed as the module `czk_b200`."""
import importlib.util
import sys
from pathlib import Path

_pkg_dir = Path(__file__).resolve().parent / "collaborative-zksnark_b200"
_spec = importlib.util.spec_from_file_location("czk_b200", _pkg_dir / "__init__.py", submodule_search_locations=[str(_pkg_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["czk_b200"] = _mod
_spec.loader.exec_module(_mod)
