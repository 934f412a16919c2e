"""Import alias: loads the package directory `3dy4-real-time-software-defined-radio-_b200/`
(a name Python cannot import directly) as the module ``dy4_b200``."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "3dy4-real-time-software-defined-radio-_b200")
_spec = _u.spec_from_file_location("dy4_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["dy4_b200"] = _mod
_spec.loader.exec_module(_mod)
