"""Loader for the CPU oracle (oracle/libvpe_ref.so). TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import this."""
import ctypes as C
import os
import subprocess

import vpe_b200
from vpe_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "libvpe_ref.so")
_lib = None


def build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)


def load_oracle():
    global _lib
    if _lib is None:
        src = os.path.join(ORACLE_DIR, "vpe_ref.cpp")
        if not os.path.exists(ORACLE_LIB) or os.path.getmtime(ORACLE_LIB) < os.path.getmtime(src):
            build_oracle()
        lib = C.CDLL(ORACLE_LIB)
        _abi.bind(lib)
        assert lib.vpe_backend() == b"oracle"
        P = C.c_void_p
        lib.vpe_ref_f32_to_f16.restype = C.c_uint16
        lib.vpe_ref_f32_to_f16.argtypes = [C.c_float]
        lib.vpe_ref_f16_to_f32.restype = C.c_float
        lib.vpe_ref_f16_to_f32.argtypes = [C.c_uint16]
        lib.vpe_ref_num_threads.restype = C.c_int
        lib.vpe_ref_set_num_threads.restype = C.c_int
        lib.vpe_ref_set_num_threads.argtypes = [C.c_int]
        lib.vpe_ref_march_partial.restype = C.c_int
        lib.vpe_ref_march_partial.argtypes = [P, C.POINTER(_abi.VpeCamera), P, P, P]
        lib.vpe_ref_touched_metavoxels.restype = C.c_int
        lib.vpe_ref_touched_metavoxels.argtypes = [P, C.POINTER(_abi.VpeCamera), P, C.c_int, P]
        lib.vpe_ref_write_light_sheet.restype = C.c_int
        lib.vpe_ref_write_light_sheet.argtypes = [P, P]
        _lib = lib
    return _lib


def oracle_engine(scene, **overrides):
    return vpe_b200.engine_for_scene(load_oracle(), scene, **overrides)
