"""Import shim: `import sshslie_b200` loads the package directory
`self-supervised-image-enhancement-network-training-with-low-light-images-only_b200/` under this name."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "self-supervised-image-enhancement-network-training-with-low-light-images-only_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
