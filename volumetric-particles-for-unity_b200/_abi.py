"""ctypes declaration of include/vpe.h (the C-ABI both libvpe_cuda.so and the oracle export).

This module only *declares* the interface; which shared object is loaded is decided by the caller:
the product (`engine.py`) loads libvpe_cuda.so and nothing else, the tests additionally load the
oracle through `tests/oracle_lib.py`.
"""
import ctypes as C

VPE_OK = 0
VPE_E_INVALID_ARG = -1
VPE_E_NOT_READY = -2
VPE_E_CUDA = -3
VPE_E_OUT_OF_MEMORY = -4
VPE_E_UNSUPPORTED = -5

VPE_BIN_REFERENCE = 0
VPE_BIN_EXACT = 1


class VpeTransform(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("rotation", C.c_float * 4)]


class VpeConfig(C.Structure):
    _fields_ = [
        ("numMetavoxelsX", C.c_int32), ("numMetavoxelsY", C.c_int32), ("numMetavoxelsZ", C.c_int32),
        ("mvScale", C.c_float),
        ("numVoxelsInMetavoxel", C.c_int32),
        ("numBorderVoxels", C.c_int32),
        ("rayMarchSteps", C.c_int32),
        ("ambientColor", C.c_float * 3),
        ("displacementScale", C.c_float),
        ("fadeOutParticles", C.c_int32),
        ("opacityFactor", C.c_float),
        ("softParticleStepDistance", C.c_int32),
        ("lightNear", C.c_float), ("lightFar", C.c_float),
        ("lightCameraDistance", C.c_float),
        ("binMode", C.c_int32),
        ("marchEarlyOutTransmittance", C.c_float),
        ("slabZBegin", C.c_int32), ("slabZEnd", C.c_int32),
    ]


class VpeParticle(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("size", C.c_float), ("rotationDeg", C.c_float),
                ("lifetime", C.c_float), ("startLifetime", C.c_float)]


class VpeCamera(C.Structure):
    _fields_ = [("transform", VpeTransform), ("fovYDegrees", C.c_float),
                ("width", C.c_int32), ("height", C.c_int32)]


class VpeStats(C.Structure):
    _fields_ = [
        ("numParticles", C.c_int32), ("numMetavoxelsCovered", C.c_int32),
        ("numParticlePairs", C.c_int64), ("voxelsFilled", C.c_int64), ("raySamples", C.c_int64),
        ("zBoundary", C.c_int32), ("fillLaunches", C.c_int32), ("marchLaunches", C.c_int32),
        ("fillMs", C.c_float), ("marchMs", C.c_float),
        ("brickPoolBytes", C.c_int64),
        ("fillKernelMs", C.c_float), ("marchKernelMs", C.c_float),
        ("raySamplesSkipped", C.c_int64),
    ]


class VpeMarchOptions(C.Structure):
    _fields_ = [("targetFormat", C.c_int32), ("debugMode", C.c_int32), ("sceneDepth", C.c_void_p),
                ("sceneWidth", C.c_int32), ("sceneHeight", C.c_int32)]


class VpeDebugOptions(C.Structure):
    _fields_ = [("marchKernel", C.c_int32), ("noSkip", C.c_int32), ("noGray", C.c_int32), ("noRowPad", C.c_int32),
                ("marchBands", C.c_int32), ("marchTileLog2W", C.c_int32), ("linkSpinMs", C.c_int32),
                ("sweepOverlap", C.c_int32), ("noTmaSweep", C.c_int32), ("profileSlices", C.c_int32), ("noHeadFused", C.c_int32), ("reserved", C.c_int32 * 5)]


ABI_VERSION = 2   # VPE_ABI_VERSION of include/vpe.h this mirror was written against

assert C.sizeof(VpeParticle) == 28

_P = C.c_void_p

# every symbol include/vpe.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "vpe_default_config": (None, [C.POINTER(VpeConfig)]),
    "vpe_create": (C.c_int, [C.POINTER(VpeConfig), C.c_int, C.POINTER(_P)]),
    "vpe_destroy": (C.c_int, [_P]),
    "vpe_set_config": (C.c_int, [_P, C.POINTER(VpeConfig)]),
    "vpe_set_light": (C.c_int, [_P, C.POINTER(VpeTransform), C.POINTER(C.c_float)]),
    "vpe_set_displacement_cubemap": (C.c_int, [_P, _P, C.c_int]),
    "vpe_set_light_depth_map": (C.c_int, [_P, _P]),
    "vpe_render_light_depth_map": (C.c_int, [_P, _P, C.c_int]),
    "vpe_read_light_depth_map": (C.c_int, [_P, _P]),
    "vpe_set_march_options": (C.c_int, [_P, C.POINTER(VpeMarchOptions)]),
    "vpe_set_debug_options": (C.c_int, [_P, C.POINTER(VpeDebugOptions)]),
    "vpe_read_slice_profile": (C.c_int, [_P, _P, _P, _P]),
    "vpe_composite_scene": (C.c_int, [_P, _P, _P, C.c_int, C.c_int]),
    "vpe_fill": (C.c_int, [_P, _P, C.c_int, C.POINTER(VpeTransform)]),
    "vpe_march": (C.c_int, [_P, C.POINTER(VpeCamera), _P, _P]),
    "vpe_march_pixels": (C.c_int, [_P, C.POINTER(VpeCamera), _P, C.c_int, _P, _P]),
    "vpe_set_stream": (C.c_int, [_P, _P]),
    "vpe_fill_device": (C.c_int, [_P, _P, C.c_int, C.POINTER(VpeTransform)]),
    "vpe_march_device": (C.c_int, [_P, C.POINTER(VpeCamera), _P, _P]),
    "vpe_fill_prepare": (C.c_int, [_P, _P, C.c_int, C.POINTER(VpeTransform), C.c_int]),
    "vpe_fill_region": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vpe_fill_density": (C.c_int, [_P]),
    "vpe_fill_sweep_region": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vpe_light_sheet_device": (_P, [_P]),
    "vpe_sheet_link_create": (C.c_int, [_P, _P, C.POINTER(_P)]),
    "vpe_sheet_link_connect": (C.c_int, [_P, _P, _P, C.c_int]),
    "vpe_fill_sweep_linked": (C.c_int, [_P]),
    "vpe_fill_linked": (C.c_int, [_P]),
    "vpe_read_sample_bitmap": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.POINTER(C.c_int)]),
    "vpe_sheet_link_status": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "vpe_march_partial_device": (C.c_int, [_P, C.POINTER(VpeCamera), _P, _P, _P]),
    "vpe_image_link_create": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.POINTER(_P)]),
    "vpe_image_link_connect": (C.c_int, [_P, C.POINTER(_P), C.c_int]),
    "vpe_march_linked": (C.c_int, [_P, C.POINTER(VpeCamera), _P]),
    "vpe_composite_linked": (C.c_int, [_P, _P]),
    "vpe_image_link_status": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "vpe_composite_device": (C.c_int, [_P, C.POINTER(_P), C.c_int, C.c_int, _P]),
    "vpe_march_footprint": (C.c_int, [_P, C.POINTER(VpeCamera), C.POINTER(C.c_int64)]),
    "vpe_read_brick": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.POINTER(C.c_int)]),
    "vpe_read_light_sheet": (C.c_int, [_P, _P]),
    "vpe_read_particle_list": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.POINTER(C.c_int)]),
    "vpe_read_metavoxel_position": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "vpe_debug_div_rn": (C.c_int, [_P, _P, _P, _P, C.c_int]),
    "vpe_get_stats": (C.c_int, [_P, C.POINTER(VpeStats)]),
    "vpe_last_error": (C.c_char_p, [_P]),
    "vpe_abi_version": (C.c_int, []),
    "vpe_backend": (C.c_char_p, []),
}


def bind(lib):
    """Attach restype/argtypes for every declared symbol; raises AttributeError if one is missing."""
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
