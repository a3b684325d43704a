"""Import shim: the package directory is named `pli-slam_b200` (not a valid identifier), so `import plf` loads it."""
import importlib.util as _u
import os as _os
import sys as _sys

_d = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "pli-slam_b200")
_spec = _u.spec_from_file_location("pli_slam_b200", _os.path.join(_d, "__init__.py"), submodule_search_locations=[_d])
_m = _u.module_from_spec(_spec)
_sys.modules["pli_slam_b200"] = _m
_spec.loader.exec_module(_m)
globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
