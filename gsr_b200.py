"""Import alias for the package directory `gnn-sparsification-research_b200/`.

The directory name is fixed by the project layout and is not a valid Python
identifier, so `import gsr_b200` loads it under this name (sub-modules resolve
as `gsr_b200.<name>`).
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gnn-sparsification-research_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR]
)
_module = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _module
_spec.loader.exec_module(_module)
