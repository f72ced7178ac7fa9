"""Import alias: the package directory name required by the build contract
(`multi-modal-misinformation-detection-with-explanation-generation_b200/`) is not a
valid Python identifier, so `import mmf_b200` loads that directory as a package."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "multi-modal-misinformation-detection-with-explanation-generation_b200")
_spec = importlib.util.spec_from_file_location(
    "mmf_b200", os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mmf_b200"] = _mod
_spec.loader.exec_module(_mod)
