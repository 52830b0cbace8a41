"""Import alias: `import vpe_b200` == the package in ./volumetric-particles-for-unity_b200/
(a hyphenated directory name cannot appear in an `import` statement)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("volumetric-particles-for-unity_b200")
sys.modules[__name__] = _pkg
