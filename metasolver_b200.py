"""Import shim: `import metasolver_b200` loads the package in ./neural-ode-metasolver_b200/.

The package directory carries the project's name (with hyphens), which Python cannot import
directly; this module replaces itself in sys.modules with that package.
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "neural-ode-metasolver_b200")
_spec = importlib.util.spec_from_file_location("metasolver_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["metasolver_b200"] = _mod
_spec.loader.exec_module(_mod)
