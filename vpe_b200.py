"""Import alias: `import vpe_b200` == the package in ./volumetric-particles-for-unity_b200/
(a hyphenated directory name cannot appear in an `import` statement)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_real = "volumetric-particles-for-unity_b200"
_pkg = importlib.import_module(_real)
for _sub in ("_abi", "engine", "scenes", "renderer", "slabs"):
    try:
        importlib.import_module(_real + "." + _sub)
    except ModuleNotFoundError:
        pass
# one module object per submodule, under both names (otherwise ctypes classes would be duplicated)
for _name, _mod in list(sys.modules.items()):
    if _name == _real or _name.startswith(_real + "."):
        sys.modules[__name__ + _name[len(_real):]] = _mod
