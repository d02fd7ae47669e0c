"""Import alias: the package directory name required by the repo layout contains dashes, so it cannot be
imported by name.  ``import ae_b200`` loads it under this module name."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "hybrid-autoencoder-mlp-pipeline-for-satellite-image-classification_b200")
_spec = importlib.util.spec_from_file_location("ae_b200", os.path.join(_PKG_DIR, "__init__.py"),
                                               submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ae_b200"] = _mod
_spec.loader.exec_module(_mod)
