"""Import shim: loads the package directory ``incubator-predictionio_b200/`` as module ``pio_b200``."""
import importlib.util as _u
import sys as _sys
from pathlib import Path as _P

_dir = _P(__file__).resolve().parent / "incubator-predictionio_b200"
_spec = _u.spec_from_file_location("pio_b200", _dir / "__init__.py", submodule_search_locations=[str(_dir)])
_mod = _u.module_from_spec(_spec)
_sys.modules["pio_b200"] = _mod
_spec.loader.exec_module(_mod)
