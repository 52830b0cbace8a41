"""vpe — B200-native sparse volumetric particle engine (Fill Volume + Ray March hot path of
rajabala/Volumetric-Particles-For-Unity behind a C-ABI).

The directory name carries a hyphen (the project's name), so import it with
`importlib.import_module("volumetric-particles-for-unity_b200")` or through the `vpe_b200` alias
module at the repository root.
"""
from . import _abi, scenes  # noqa: F401
from .engine import CUDA_LIB_PATH, Engine, VpeError, engine_for_scene, load_cuda_library  # noqa: F401
