"""Import alias for the package directory ``cmf.jl_b200/``.

The package directory carries the reference's name (CMF.jl) and therefore a dot, which Python's
import system cannot spell.  ``import cmf_jl_b200`` loads that directory as a regular package
under this module name (sub-modules ``cmf_jl_b200.model`` etc. resolve into ``cmf.jl_b200/``).
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cmf.jl_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
