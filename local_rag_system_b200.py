"""Import alias: `import local_rag_system_b200` loads the package that lives in the
hyphenated directory `local-rag-system_b200/` (not a valid Python identifier)."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "local-rag-system_b200")
_spec = importlib.util.spec_from_file_location(
    "local_rag_system_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["local_rag_system_b200"] = _mod
_spec.loader.exec_module(_mod)
