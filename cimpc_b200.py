"""Import shim: the package directory `contactimplicitmpc.jl_b200/` is not a valid dotted module
name, so load it by path and expose it as `cimpc_b200`."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "contactimplicitmpc.jl_b200")
_name = "contactimplicitmpc_jl_b200"
if _name not in sys.modules:
    _spec = importlib.util.spec_from_file_location(
        _name, os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_name] = _mod
    _spec.loader.exec_module(_mod)
_mod = sys.modules[_name]
globals().update({k: getattr(_mod, k) for k in dir(_mod) if not k.startswith("__")})
package = _mod
