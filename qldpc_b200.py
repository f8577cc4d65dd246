"""Import shim: the package directory is named ``qldpc-branched-off_b200`` (not a valid
Python identifier), so ``import qldpc_b200`` loads that directory as the package
``qldpc_b200``.  After this runs, ``sys.modules['qldpc_b200']`` is the real package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "qldpc-branched-off_b200")
_spec = importlib.util.spec_from_file_location(
    "qldpc_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["qldpc_b200"] = _mod
_spec.loader.exec_module(_mod)
