"""Import alias: `import cetkmc` loads the package directory
`cet-driven-simulation-for-3d-printing-am-kmc-approach_b200/` (whose name is not a Python
identifier) under the module name `cetkmc`, once."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "cet-driven-simulation-for-3d-printing-am-kmc-approach_b200")
_spec = importlib.util.spec_from_file_location(
    "cetkmc", os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["cetkmc"] = _mod
_spec.loader.exec_module(_mod)
