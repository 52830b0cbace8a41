"""Python host mirror of the reference's MonoBehaviour `VolumetricParticleRenderer`
(Assets/Main Scene/VolumetricParticleRenderer.cs, "VPR.cs") over the C-ABI — same field and method
names, same frame structure (fill every `updateInterval` frames, march every frame), so that code
and tests written against it read like the reference.  The C++ twin is host/cpp/.

What Unity supplied implicitly is passed explicitly: `particles` is what
ParticleSystem.GetParticles() returned (n x 7 float32: position xyz in emitter space, size,
rotation in degrees, lifetime, startLifetime — the fields read at VPR.cs:418,425,583-586), `camera`
is Camera.main (dict: position, rotation quaternion, fovYDegrees, width, height) and the returned
array is particlesRT (VPR.cs:228) as float32 RGBA, premultiplied, before CompositeParticles.
"""
import numpy as np

from . import scenes
from .engine import Engine


class VolumetricParticleRenderer:
    def __init__(self, lib=None, device=0):
        """lib None = the CUDA product (libvpe_cuda.so); tests may pass another bound library."""
        self._lib, self._device, self._engine = lib, device, None
        # game objects that need to be set (VPR.cs:72-79), as plain transforms
        self.dirLight = {"position": scenes.LIGHT_POSITION, "rotation": scenes.LIGHT_ROTATION}
        self.particleSys = {"position": (0.0, 5.0, 11.2), "rotation": (0.0, 1.0, 0.0, 0.0)}
        self.gridCenter = (0.0, 0.0, 0.0)
        # metavoxel layout/size (VPR.cs:82-85); defaults = the demo scene (scene:9013-9026)
        self.numMetavoxelsX = self.numMetavoxelsY = self.numMetavoxelsZ = 10
        self.mvScale = (3.0, 3.0, 3.0)
        self.numVoxelsInMetavoxel = 32
        self.numBorderVoxels = 1
        # rendering vars (VPR.cs:88-101)
        self.updateInterval = 2
        self.rayMarchSteps = 64
        self.ambientColor = (0.2, 0.2, 0.2)
        self.fDisplacementScale = 0.7
        self.fadeOutParticles = False
        self.opacityFactor = 0.04
        self.softParticleStepDistance = 20
        # debug views (VPR.cs:96-99) and the render target (particlesRT is ARGB32, VPR.cs:228; float by default here)
        self.bShowRayMarchSamplesPerPixel = False
        self.bShowMetavoxelDrawOrder = False
        self.bShowRayMarchBlendFunc = False
        self.particlesRT_8bit = False
        # counters (VPR.cs:124-125)
        self.numParticlesEmitted = 0
        self.numMetavoxelsCovered = 0
        self._frameCount = 0

    # -- Unity callbacks ------------------------------------------------------------------------
    def Start(self, displacement_r8=None):
        """VPR.cs:132-149."""
        self.fadeOutParticles = False  # VPR.cs:134
        kw = self._config()
        self._engine = Engine.cuda(self._device, **kw) if self._lib is None else Engine(self._lib, self._device, **kw)
        self._engine.set_displacement_cubemap(displacement_r8 if displacement_r8 is not None else scenes.load_displacement_cubemap())
        self._engine.set_light_depth_map(None)
        self._frameCount = 0
        self.UpdateMetavoxelPositions()

    def OnPostRender(self, particles, camera, occluders=None, mainSceneRT=None, sceneDepth=None):
        """VPR.cs:181-220. Returns particlesRT for this frame, or, when mainSceneRT (H x W x 4 scene colour) is
        given, the scene with the particles blended onto it (VPR.cs:210). occluders: (n, 3, 3) world-space
        triangles of the Default layer seen by the light camera (VPR.cs:184); sceneDepth: (H, W) eye-space
        depth of mainSceneRT.depthBuffer (VPR.cs:204)."""
        if occluders is not None:
            self._engine.render_light_depth_map(occluders)   # VPR.cs:184, every frame
        if self.updateInterval < 1 or self._frameCount % self.updateInterval == 0:  # VPR.cs:186
            self._push_config()
            self.UpdateMetavoxelPositions()        # VPR.cs:188-195
            self.FillMetavoxels(particles)         # VPR.cs:197-198
        self._frameCount += 1
        particlesRT = self.RenderMetavoxels(camera, sceneDepth=sceneDepth)   # VPR.cs:204-207
        if mainSceneRT is None:
            return particlesRT
        return self._engine.composite_scene(particlesRT, mainSceneRT, 1 if self.particlesRT_8bit else 0)  # VPR.cs:210

    # -- the two dispatch entry points ----------------------------------------------------------
    def UpdateMetavoxelPositions(self):
        """VPR.cs:370-394."""
        self._engine.set_light(self.dirLight["position"], self.dirLight["rotation"], self.gridCenter)

    def SetLightDepthMap(self, depth01):
        """lightDepthMap (VPR.cs:184,274): (NY*N, NX*N) floats in [0,1], None = no occluders."""
        self._engine.set_light_depth_map(depth01)

    def FillMetavoxels(self, particles):
        """BinParticlesToMetavoxels + FillMetavoxels, VPR.cs:397-520."""
        self._engine.fill(np.asarray(particles, dtype=np.float32), self.particleSys)
        st = self._engine.stats()
        self.numParticlesEmitted, self.numMetavoxelsCovered = st["numParticles"], st["numMetavoxelsCovered"]

    def RenderMetavoxels(self, camera, show_samples=False, sceneDepth=None):
        """VPR.cs:637-713 (+ SetRaymarchPassConstants' debug flags, VPR.cs:744-761). show_samples: also return
        the per-pixel ray-sample count."""
        self._push_config()
        debug = 1 if self.bShowMetavoxelDrawOrder else 2 if self.bShowRayMarchBlendFunc else 3 if self.bShowRayMarchSamplesPerPixel else 0
        self._engine.set_march_options(target_format=1 if self.particlesRT_8bit else 0, debug_mode=debug, scene_depth=sceneDepth)
        rgba, samples = self._engine.march(camera, want_samples=show_samples)
        return (rgba, samples) if show_samples else rgba

    def DrawMetavoxelGrid(self):
        """VPR.cs:962-1033 (bShowMetavoxelGrid): the wire cube of every covered metavoxel as world-space line
        segments, shape (covered, 12, 2, 3), in the reference's vertex order (left face, right face, the four
        joining edges). Drawing them (GL.LINES over mainSceneRT) is left to the host application."""
        e = self._engine
        q = np.asarray(self.dirLight["rotation"], dtype=np.float64)
        hw, hh, hd = [0.5 * float(v) for v in self.mvScale]
        corner = {}
        for name, v in (("BLB", (-hw, -hh, hd)), ("BLF", (-hw, -hh, -hd)), ("TLB", (-hw, hh, hd)), ("TLF", (-hw, hh, -hd)),
                        ("BRB", (hw, -hh, hd)), ("BRF", (hw, -hh, -hd)), ("TRB", (hw, hh, hd)), ("TRF", (hw, hh, -hd))):
            corner[name] = scenes.quat_rotate(q, np.asarray([v]))[0]           # lightOrientation * offset (VPR.cs:989-996)
        edges = [("BLB", "BLF"), ("BLF", "TLF"), ("TLF", "TLB"), ("TLB", "BLB"),     # left   (VPR.cs:998-1006)
                 ("BRB", "BRF"), ("BRF", "TRF"), ("TRF", "TRB"), ("TRB", "BRB"),     # right  (VPR.cs:1007-1015)
                 ("TLB", "TRB"), ("TLF", "TRF"), ("BLB", "BRB"), ("BLF", "BRF")]     # joins  (VPR.cs:1016-1025)
        out = []
        for zz in range(self.numMetavoxelsZ):
            for yy in range(self.numMetavoxelsY):
                for xx in range(self.numMetavoxelsX):
                    if e.read_particle_list(xx, yy, zz).shape[0] == 0:     # mParticlesCovered.Count != 0 (VPR.cs:970)
                        continue
                    pos = e.read_metavoxel_position(xx, yy, zz).astype(np.float64)
                    out.append([[pos + corner[a], pos + corner[b]] for a, b in edges])
        return np.asarray(out, dtype=np.float32).reshape(-1, 12, 2, 3)

    # -- GUI callback setters, VPR.cs:1040-1119 -------------------------------------------------
    def SetDisplacementScale(self, ds):
        self.fDisplacementScale = float(ds)

    def SetRayMarchSteps(self, steps):
        self.rayMarchSteps = int(steps)

    def SetGridScale(self, s):
        self.mvScale = (float(s),) * 3
        self._push_config()
        self.UpdateMetavoxelPositions()

    def SetFadeOutParticles(self, fade):
        self.fadeOutParticles = bool(fade)

    SetFadeParticles = SetFadeOutParticles

    def SetParticleOpacityFactor(self, f):
        self.opacityFactor = float(f)

    def ShowRayMarchSamplesPerPixel(self, show):   # VPR.cs:1091-1094
        self.bShowRayMarchSamplesPerPixel = bool(show)

    def ShowMetavoxelDrawOrder(self, show):        # VPR.cs:1096-1099
        self.bShowMetavoxelDrawOrder = bool(show)

    def ShowRayMarchBlendFunc(self, show):
        self.bShowRayMarchBlendFunc = bool(show)

    def SetUpdateInterval(self, interval):
        self.updateInterval = int(interval)

    def SetSoftParticleDistance(self, step_distance):
        self.softParticleStepDistance = int(step_distance)

    # -- plumbing -------------------------------------------------------------------------------
    @property
    def engine(self):
        return self._engine

    def _config(self):
        return dict(grid=(self.numMetavoxelsX, self.numMetavoxelsY, self.numMetavoxelsZ), mvScale=self.mvScale[0],
                    numVoxels=self.numVoxelsInMetavoxel, border=self.numBorderVoxels, rayMarchSteps=self.rayMarchSteps,
                    ambient=self.ambientColor, displacementScale=self.fDisplacementScale,
                    fadeOutParticles=1 if self.fadeOutParticles else 0, opacityFactor=self.opacityFactor,
                    softDistance=self.softParticleStepDistance)

    def _push_config(self):
        self._engine.set_config(mvScale=float(self.mvScale[0]), rayMarchSteps=int(self.rayMarchSteps),
                                ambientColor=self.ambientColor, displacementScale=float(self.fDisplacementScale),
                                fadeOutParticles=1 if self.fadeOutParticles else 0, opacityFactor=float(self.opacityFactor),
                                softParticleStepDistance=int(self.softParticleStepDistance))
