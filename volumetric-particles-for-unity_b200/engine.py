"""Thin Python host over the C-ABI of include/vpe.h.

`Engine.cuda(...)` is the product path: it loads csrc/libvpe_cuda.so and fails loudly when the
library is missing or was not built — there is no CPU fallback.  `Engine(lib, ...)` with another
bound library is used by the tests to drive the oracle through the very same calls.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import VpeCamera, VpeConfig, VpeStats, VpeTransform

_HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_LIB_PATH = os.path.join(_HERE, "csrc", "libvpe_cuda.so")
_cuda_lib = None


class VpeError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("vpe error %d: %s" % (code, message))
        self.code = code


def load_cuda_library():
    """Load the CUDA engine. No fallback: a missing library is an error."""
    global _cuda_lib
    if _cuda_lib is None:
        if not os.path.exists(CUDA_LIB_PATH):
            raise RuntimeError(
                "libvpe_cuda.so is not built (%s). Run `python __graft_entry__.py build` "
                "(nvcc, sm_100a). There is no CPU fallback for the hot path." % CUDA_LIB_PATH)
        lib = C.CDLL(CUDA_LIB_PATH)
        _abi.bind(lib)
        if lib.vpe_backend() != b"cuda":
            raise RuntimeError("libvpe_cuda.so reports backend %r" % lib.vpe_backend())
        if lib.vpe_abi_version() != _abi.ABI_VERSION:
            raise RuntimeError("ABI version mismatch")
        _cuda_lib = lib
    return _cuda_lib


def make_config(lib, grid=(10, 10, 10), mvScale=3.0, numVoxels=32, border=1, rayMarchSteps=64,
                ambient=(0.2, 0.2, 0.2), displacementScale=0.7, fadeOutParticles=0,
                opacityFactor=0.04, softDistance=20, binMode=_abi.VPE_BIN_REFERENCE,
                earlyOut=0.0, slab=(0, 0), **_ignored):
    cfg = VpeConfig()
    lib.vpe_default_config(C.byref(cfg))
    cfg.numMetavoxelsX, cfg.numMetavoxelsY, cfg.numMetavoxelsZ = [int(g) for g in grid]
    cfg.mvScale = float(mvScale)
    cfg.numVoxelsInMetavoxel = int(numVoxels)
    cfg.numBorderVoxels = int(border)
    cfg.rayMarchSteps = int(rayMarchSteps)
    cfg.ambientColor[:] = [float(a) for a in ambient]
    cfg.displacementScale = float(displacementScale)
    cfg.fadeOutParticles = int(fadeOutParticles)
    cfg.opacityFactor = float(opacityFactor)
    cfg.softParticleStepDistance = int(softDistance)
    cfg.binMode = int(binMode)
    cfg.marchEarlyOutTransmittance = float(earlyOut)
    cfg.slabZBegin, cfg.slabZEnd = int(slab[0]), int(slab[1])
    return cfg


def _transform(position, rotation):
    t = VpeTransform()
    t.position[:] = [float(v) for v in position]
    t.rotation[:] = [float(v) for v in rotation]
    return t


def _camera(cam):
    c = VpeCamera()
    c.transform = _transform(cam["position"], cam["rotation"])
    c.fovYDegrees = float(cam["fovYDegrees"])
    c.width, c.height = int(cam["width"]), int(cam["height"])
    return c


def _particles(p):
    p = np.ascontiguousarray(p, dtype=np.float32)
    if p.ndim != 2 or p.shape[1] != 7:
        raise ValueError("particles must be (n, 7) float32: position xyz, size, rotationDeg, lifetime, startLifetime")
    return p


class Engine:
    """One VpeContext. Mirrors the C-ABI one to one; numpy in, numpy out."""

    def __init__(self, lib, device=0, **config):
        self.lib = lib
        self.cfg = make_config(lib, **config)
        self._ctx = C.c_void_p()
        rc = lib.vpe_create(C.byref(self.cfg), int(device), C.byref(self._ctx))
        if rc != 0:
            raise VpeError(rc, "vpe_create failed (invalid configuration or no CUDA device)")
        self.N = self.cfg.numVoxelsInMetavoxel
        self.grid = (self.cfg.numMetavoxelsX, self.cfg.numMetavoxelsY, self.cfg.numMetavoxelsZ)

    @classmethod
    def cuda(cls, device=0, **config):
        return cls(load_cuda_library(), device=device, **config)

    # -- lifetime -------------------------------------------------------------------------
    def close(self):
        if self._ctx:
            self.lib.vpe_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise VpeError(rc, (self.lib.vpe_last_error(self._ctx) or b"").decode())

    @property
    def backend(self):
        return self.lib.vpe_backend().decode()

    # -- configuration --------------------------------------------------------------------
    def set_config(self, **changes):
        """≙ the GUI setters (VPR.cs:1040-1119): keyword = VpeConfig field name."""
        for k, v in changes.items():
            if k == "ambientColor":
                self.cfg.ambientColor[:] = [float(a) for a in v]
            else:
                setattr(self.cfg, k, v)
        self._check(self.lib.vpe_set_config(self._ctx, C.byref(self.cfg)))

    def set_light(self, position, rotation, grid_center=(0.0, 0.0, 0.0)):
        t = _transform(position, rotation)
        gc = (C.c_float * 3)(*[float(v) for v in grid_center])
        self._check(self.lib.vpe_set_light(self._ctx, C.byref(t), gc))

    def set_displacement_cubemap(self, r8):
        r8 = np.ascontiguousarray(r8, dtype=np.uint8)
        assert r8.ndim == 3 and r8.shape[0] == 6 and r8.shape[1] == r8.shape[2]
        self._check(self.lib.vpe_set_displacement_cubemap(self._ctx, r8.ctypes.data, r8.shape[1]))

    def set_light_depth_map(self, depth01):
        if depth01 is None:
            self._check(self.lib.vpe_set_light_depth_map(self._ctx, None))
            return
        d = np.ascontiguousarray(depth01, dtype=np.float32)
        assert d.shape == (self.grid[1] * self.N, self.grid[0] * self.N)
        self._check(self.lib.vpe_set_light_depth_map(self._ctx, d.ctypes.data))

    # -- around the path ---------------------------------------------------------
    def render_light_depth_map(self, triangles):
        """≙ lightCamera.RenderWithShader(generateLightDepthMapShader) (VPR.cs:184): triangles is
        (n, 3, 3) world-space vertices, Unity winding (clockwise = front)."""
        t = np.ascontiguousarray(triangles, dtype=np.float32).reshape(-1, 9)
        self._check(self.lib.vpe_render_light_depth_map(self._ctx, t.ctypes.data if t.shape[0] else None, t.shape[0]))

    def read_light_depth_map(self):
        out = np.empty((self.grid[1] * self.N, self.grid[0] * self.N), dtype=np.float32)
        self._check(self.lib.vpe_read_light_depth_map(self._ctx, out.ctypes.data))
        return out

    def set_march_options(self, target_format=0, debug_mode=0, scene_depth=None):
        """target_format 1 = UNORM8 target like particlesRT (VPR.cs:228); debug_mode 1/2/3 = draw order /
        blend function / sample-count views (March.shader:170-181,283-299); scene_depth = (H, W) eye-space
        depths of the opaque scene (≙ ZTest Less against mainSceneRT.depthBuffer, VPR.cs:204)."""
        o = _abi.VpeMarchOptions()
        o.targetFormat, o.debugMode = int(target_format), int(debug_mode)
        keep = None
        if scene_depth is not None:
            keep = np.ascontiguousarray(scene_depth, dtype=np.float32)
            o.sceneDepth = keep.ctypes.data
            o.sceneHeight, o.sceneWidth = keep.shape
        self._check(self.lib.vpe_set_march_options(self._ctx, C.byref(o)))

    def set_debug_options(self, march_kernel=0, no_skip=False, no_gray=False, no_row_pad=False, march_bands=0,
                          march_tile_log2w=None, link_spin_ms=0, sweep_overlap=False, no_tma_sweep=False, profile_slices=False,
                          no_head_fused=False):
        """VpeDebugOptions: experiment switches of the CUDA library (all defaults = production). march_kernel 1 = general
        kernel, 2 = round 1's per-fragment loop; no_gray / no_row_pad change the brick layout (fill again before marching)."""
        o = _abi.VpeDebugOptions()
        o.marchKernel, o.noSkip, o.noGray, o.noRowPad = int(march_kernel), int(bool(no_skip)), int(bool(no_gray)), int(bool(no_row_pad))
        o.marchBands = int(march_bands)
        o.marchTileLog2W = 0 if march_tile_log2w is None else int(march_tile_log2w) + 1
        o.linkSpinMs = int(link_spin_ms)
        o.sweepOverlap = int(bool(sweep_overlap))
        o.noTmaSweep = int(bool(no_tma_sweep))
        o.profileSlices = int(bool(profile_slices))
        o.noHeadFused = int(bool(no_head_fused))
        self._check(self.lib.vpe_set_debug_options(self._ctx, C.byref(o)))

    def debug_div_rn(self, a, b):
        """The fill kernel's inlined division, elementwise (test hook)."""
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        q = np.empty_like(a)
        self._check(self.lib.vpe_debug_div_rn(self._ctx, a.ctypes.data, b.ctypes.data, q.ctypes.data, a.size))
        return q

    def read_slice_profile(self):
        """(pairs, covered metavoxels, ray samples) per light-axis slice of the last fill / march (needs profile_slices)."""
        nz = self.grid[2]
        out = [np.zeros(nz, dtype=np.int64) for _ in range(3)]
        self._check(self.lib.vpe_read_slice_profile(self._ctx, out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data))
        return tuple(out)

    def composite_scene(self, particles_rgba, scene_rgba, target_format=0):
        """≙ Graphics.Blit(particlesRT, mainSceneRT, matBlendParticles) (VPR.cs:210). Returns the new scene."""
        p = np.ascontiguousarray(particles_rgba, dtype=np.float32)
        sc = np.array(scene_rgba, dtype=np.float32, order="C", copy=True)
        assert p.shape == sc.shape and p.shape[-1] == 4
        self._check(self.lib.vpe_composite_scene(self._ctx, p.ctypes.data, sc.ctypes.data, p.size // 4, int(target_format)))
        return sc

    # -- hot path, host buffers -----------------------------------------------------------
    def fill(self, particles, emitter):
        p = _particles(particles)
        t = _transform(emitter["position"], emitter["rotation"])
        self._check(self.lib.vpe_fill(self._ctx, p.ctypes.data, p.shape[0], C.byref(t)))

    def fill_prepare(self, particles, emitter):
        p = _particles(particles)
        t = _transform(emitter["position"], emitter["rotation"])
        self._check(self.lib.vpe_fill_prepare(self._ctx, p.ctypes.data, p.shape[0], C.byref(t), 0))

    def fill_region(self, x0, x1, y0, y1):
        self._check(self.lib.vpe_fill_region(self._ctx, x0, x1, y0, y1))

    def fill_density(self):
        self._check(self.lib.vpe_fill_density(self._ctx))

    def fill_sweep_region(self, x0, x1, y0, y1):
        self._check(self.lib.vpe_fill_sweep_region(self._ctx, x0, x1, y0, y1))

    # -- sheet link: the multi-GPU sweep over NVLink peer memory (CUDA library only) ---------
    def sheet_link_create(self):
        """Returns (64-byte CUDA IPC handle, device pointer) of this context's link buffer."""
        h = (C.c_ubyte * 64)()
        p = C.c_void_p()
        self._check(self.lib.vpe_sheet_link_create(self._ctx, h, C.byref(p)))
        return bytes(h), p.value

    def sheet_link_connect(self, upstream, downstream):
        """upstream / downstream: None, a 64-byte IPC handle (bytes; another process) or a device
        pointer (int; a context of this process). Both must be of the same kind."""
        kinds = {type(v) for v in (upstream, downstream) if v is not None}
        if len(kinds) > 1:
            raise ValueError("upstream and downstream must both be IPC handles or both device pointers")
        ipc = bytes in kinds

        def arg(v):
            if v is None:
                return None
            if ipc:
                return C.cast(C.create_string_buffer(v, 64), C.c_void_p)
            return C.cast(C.pointer(C.c_void_p(int(v))), C.c_void_p)
        keep = [arg(upstream), arg(downstream)]
        self._check(self.lib.vpe_sheet_link_connect(self._ctx, keep[0], keep[1], 1 if ipc else 0))

    def fill_sweep_linked(self):
        self._check(self.lib.vpe_fill_sweep_linked(self._ctx))

    def fill_linked(self):
        """The linked fill after fill_prepare: fused kernel on the head rank, density + linked sweep elsewhere."""
        self._check(self.lib.vpe_fill_linked(self._ctx))

    def sheet_link_timeouts(self):
        n = C.c_int(0)
        self._check(self.lib.vpe_sheet_link_status(self._ctx, C.byref(n)))
        return n.value

    # -- image link: the slab partials go straight into the compositing ranks' memory (CUDA library only) ------
    def image_link_create(self, world, rank, width, height):
        h = (C.c_ubyte * 64)()
        p = C.c_void_p()
        self._check(self.lib.vpe_image_link_create(self._ctx, int(world), int(rank), int(width), int(height), h, C.byref(p)))
        return bytes(h), p.value

    def image_link_connect(self, peers):
        """peers[q]: rank q's 64-byte IPC handle (bytes) or device pointer (int); the own entry may be None."""
        kinds = {type(v) for v in peers if v is not None}
        if len(kinds) != 1:
            raise ValueError("peers must all be IPC handles or all device pointers")
        ipc = bytes in kinds
        keep, arr = [], (C.c_void_p * len(peers))()
        for q, v in enumerate(peers):
            if v is None:
                arr[q] = None
                continue
            buf = C.create_string_buffer(v, 64) if ipc else C.pointer(C.c_void_p(int(v)))
            keep.append(buf)
            arr[q] = C.cast(buf, C.c_void_p)
        self._check(self.lib.vpe_image_link_connect(self._ctx, arr, 1 if ipc else 0))

    def march_linked(self, camera, samples_ptr=None):
        c = _camera(camera)
        self._check(self.lib.vpe_march_linked(self._ctx, C.byref(c), C.c_void_p(int(samples_ptr)) if samples_ptr else None))

    def composite_linked(self, band_ptr):
        self._check(self.lib.vpe_composite_linked(self._ctx, C.c_void_p(int(band_ptr))))

    def image_link_timeouts(self):
        n = C.c_int(0)
        self._check(self.lib.vpe_image_link_status(self._ctx, C.byref(n)))
        return n.value

    def march(self, camera, want_samples=True, out=None, samples_out=None):
        c = _camera(camera)
        rgba = out if out is not None else np.empty((c.height, c.width, 4), dtype=np.float32)
        samples = None
        if want_samples:
            samples = samples_out if samples_out is not None else np.empty((c.height, c.width), dtype=np.int32)
        self._check(self.lib.vpe_march(self._ctx, C.byref(c), rgba.ctypes.data,
                                       samples.ctypes.data if samples is not None else None))
        return rgba, samples

    def march_pixels(self, camera, pixels):
        c = _camera(camera)
        pix = np.ascontiguousarray(pixels, dtype=np.int32)
        rgba = np.empty((pix.shape[0], 4), dtype=np.float32)
        samples = np.empty((pix.shape[0],), dtype=np.int32)
        self._check(self.lib.vpe_march_pixels(self._ctx, C.byref(c), pix.ctypes.data, pix.shape[0],
                                              rgba.ctypes.data, samples.ctypes.data))
        return rgba, samples

    # -- hot path, device buffers (raw CUDA pointers, e.g. torch tensors' data_ptr()) --------
    def set_stream(self, cuda_stream_handle):
        self._check(self.lib.vpe_set_stream(self._ctx, C.c_void_p(int(cuda_stream_handle))))

    def fill_device(self, particles_ptr, n, emitter):
        t = _transform(emitter["position"], emitter["rotation"])
        self._check(self.lib.vpe_fill_device(self._ctx, C.c_void_p(int(particles_ptr)), int(n), C.byref(t)))

    def fill_prepare_device(self, particles_ptr, n, emitter):
        t = _transform(emitter["position"], emitter["rotation"])
        self._check(self.lib.vpe_fill_prepare(self._ctx, C.c_void_p(int(particles_ptr)), int(n), C.byref(t), 1))

    def march_device(self, camera, rgba_ptr, samples_ptr=None):
        c = _camera(camera)
        self._check(self.lib.vpe_march_device(self._ctx, C.byref(c), C.c_void_p(int(rgba_ptr)),
                                              C.c_void_p(int(samples_ptr)) if samples_ptr else None))

    def light_sheet_device_ptr(self):
        return self.lib.vpe_light_sheet_device(self._ctx)

    def march_partial_device(self, camera, over_ptr, under_ptr, samples_ptr=None):
        c = _camera(camera)
        self._check(self.lib.vpe_march_partial_device(
            self._ctx, C.byref(c), C.c_void_p(int(over_ptr)), C.c_void_p(int(under_ptr)),
            C.c_void_p(int(samples_ptr)) if samples_ptr else None))

    def composite_device(self, part_ptrs, num_pixels, rgba_ptr):
        arr = (C.c_void_p * len(part_ptrs))(*[C.c_void_p(int(p)) for p in part_ptrs])
        self._check(self.lib.vpe_composite_device(self._ctx, arr, len(part_ptrs) // 2, int(num_pixels),
                                                  C.c_void_p(int(rgba_ptr))))

    def march_footprint(self, camera):
        """Distinct texels touched by the march's trilinear footprints (measurement, never timed)."""
        c = _camera(camera)
        n = C.c_int64(0)
        self._check(self.lib.vpe_march_footprint(self._ctx, C.byref(c), C.byref(n)))
        return n.value

    # -- test hooks -----------------------------------------------------------------------
    def read_sample_bitmap(self, x, y, z):
        """The march's empty-space bitmap of a metavoxel as bool [z0][y0][x0] (None when not covered)."""
        rw = (self.N + 31) // 32
        words = np.empty((self.N, self.N, rw), dtype=np.uint32)
        covered = C.c_int(0)
        self._check(self.lib.vpe_read_sample_bitmap(self._ctx, x, y, z, words.ctypes.data, C.byref(covered)))
        if not covered.value:
            return None
        bits = (words[..., None] >> np.arange(32, dtype=np.uint32)) & 1
        return bits.reshape(self.N, self.N, rw * 32)[..., :self.N].astype(bool)

    def read_brick(self, x, y, z):
        """half4 brick as uint16 [N][N][N][4] (slice, row, col, rgba) or None when not covered."""
        out = np.empty((self.N, self.N, self.N, 4), dtype=np.uint16)
        covered = C.c_int(0)
        self._check(self.lib.vpe_read_brick(self._ctx, x, y, z, out.ctypes.data, C.byref(covered)))
        return out if covered.value else None

    def read_light_sheet(self):
        out = np.empty((self.grid[1] * self.N, self.grid[0] * self.N), dtype=np.float32)
        self._check(self.lib.vpe_read_light_sheet(self._ctx, out.ctypes.data))
        return out

    def read_particle_list(self, x, y, z):
        n = C.c_int(0)
        self._check(self.lib.vpe_read_particle_list(self._ctx, x, y, z, None, 0, C.byref(n)))
        out = np.empty((n.value,), dtype=np.int32)
        if n.value:
            self._check(self.lib.vpe_read_particle_list(self._ctx, x, y, z, out.ctypes.data, n.value, C.byref(n)))
        return out

    def read_metavoxel_position(self, x, y, z):
        p = (C.c_float * 3)()
        self._check(self.lib.vpe_read_metavoxel_position(self._ctx, x, y, z, p))
        return np.array(list(p), dtype=np.float32)

    def stats(self):
        s = VpeStats()
        self._check(self.lib.vpe_get_stats(self._ctx, C.byref(s)))
        return {name: getattr(s, name) for name, _ in VpeStats._fields_}


def engine_for_scene(lib_or_none, scene, device=0, **overrides):
    """Create an Engine configured for a scenes.make_scene() dict. lib None = the CUDA product."""
    keys = ("grid", "mvScale", "numVoxels", "border", "rayMarchSteps", "ambient", "displacementScale",
            "fadeOutParticles", "opacityFactor", "softDistance")
    kw = {k: scene[k] for k in keys}
    kw.update(overrides)
    if lib_or_none is None:
        return Engine.cuda(device=device, **kw)
    return Engine(lib_or_none, device=device, **kw)
