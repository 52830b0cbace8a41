"""Seeded synthetic scenes for BASELINE.json's configs and the reference's demo
defaults (Assets/Volumetric_Particle_System.unity:9013-9026, 8965-8988, 6763-6793, 2259-2520).

Everything is numpy (`default_rng(seed)`, PCG64) so the oracle and the CUDA engine are fed the
same little-endian fp32 arrays.  A scene is a plain dict; `apply_scene` pushes it into an Engine.
"""
import math
import os

import numpy as np

ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")

# the demo scene's directional light (scene:6763-6764, 6792-6793)
LIGHT_ROTATION = (0.185594, 0.0, 0.0, 0.982627)
LIGHT_POSITION = (0.0, 0.0, -44.34)

# name -> (grid G, voxels N, particles, width, height, steps per metavoxel, seed)
CONFIGS = {
    "cfg1": (8, 8, 32, 128, 128, 64, 1001),
    "cfg2": (16, 32, 1000, 1920, 1080, 64, 1002),
    "cfg3": (32, 32, 8000, 1920, 1080, 64, 1003),
    "cfg4": (32, 64, 16000, 3840, 2160, 64, 1004),
    "cfg5": (64, 32, 65536, 3840, 2160, 8, 1005),
}


def load_displacement_cubemap():
    """The reference's displacement cubemap, R channel, [6][128][128] uint8 (see
    tools/decode_displacement_cubemap.py)."""
    path = os.path.join(ASSET_DIR, "displacement_r8.bin")
    return np.fromfile(path, dtype=np.uint8).reshape(6, 128, 128)


def quat_rotate(q, v):
    """Rotate rows of v by the unit quaternion q = (x,y,z,w) (float64 helper for scene building)."""
    x, y, z, w = [float(t) for t in q]
    r = np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])
    return np.asarray(v, dtype=np.float64) @ r.T


def _base(grid, n_vox, scale, width, height, steps):
    return {
        "grid": (grid, grid, grid), "mvScale": float(scale), "numVoxels": n_vox, "border": 1,
        "rayMarchSteps": steps, "ambient": (0.2, 0.2, 0.2), "displacementScale": 0.7,
        "fadeOutParticles": 0, "opacityFactor": 0.04, "softDistance": 20,
        "light": {"position": LIGHT_POSITION, "rotation": LIGHT_ROTATION},
        "gridCenter": (0.0, 0.0, 0.0),
        "emitter": {"position": (0.0, 0.0, 0.0), "rotation": (0.0, 0.0, 0.0, 1.0)},
        "camera": {"position": (0.0, 0.0, -0.75 * grid * scale), "rotation": (0.0, 0.0, 0.0, 1.0),
                   "fovYDegrees": 60.0, "width": width, "height": height},
    }


def make_particles_uniform(rng, count, grid, scale):
    """Displaced spheres: centres uniform in the light-space box [-(G/2-1)s, (G/2-1)s]^3 about the
    grid centre, diameter uniform in [1.2s, 2.4s]."""
    half = (grid / 2 - 1) * scale
    ls = rng.uniform(-half, half, size=(count, 3))
    ws = quat_rotate(LIGHT_ROTATION, ls)
    p = np.empty((count, 7), dtype=np.float32)
    p[:, 0:3] = ws.astype(np.float32)
    p[:, 3] = rng.uniform(1.2 * scale, 2.4 * scale, size=count).astype(np.float32)
    p[:, 4] = rng.uniform(0.0, 360.0, size=count).astype(np.float32)
    p[:, 6] = 6.0
    p[:, 5] = (6.0 * (1.0 - rng.uniform(0.0, 1.0, size=count))).astype(np.float32)  # (0, 6]
    return p


def make_particles_plume(rng, count, grid, scale):
    """Sparse variant: Gaussian cluster, sigma = G*s/8, clipped to the grid box."""
    half = (grid / 2 - 1) * scale
    ls = np.clip(rng.normal(0.0, grid * scale / 8.0, size=(count, 3)), -half, half)
    ws = quat_rotate(LIGHT_ROTATION, ls)
    p = make_particles_uniform(rng, count, grid, scale)
    p[:, 0:3] = ws.astype(np.float32)
    return p


def make_scene(name, variant="uniform", particles=None, image=None):
    """name: cfg1..cfg5 or 'ref-defaults'.  `particles`/`image` override the counts for reduced
    parity cases (the override is part of the returned dict's name)."""
    if name == "ref-defaults":
        return _ref_defaults()
    grid, n_vox, count, width, height, steps, seed = CONFIGS[name]
    if particles is not None:
        count = particles
    if image is not None:
        width, height = image
    sc = _base(grid, n_vox, 1.0, width, height, steps)
    rng = np.random.default_rng(seed)
    make = make_particles_uniform if variant == "uniform" else make_particles_plume
    sc["particles"] = make(rng, count, grid, 1.0)
    sc["name"] = "%s-%s-p%d-%dx%d" % (name, variant, count, width, height)
    sc["seed"] = seed
    return sc


def _ref_defaults():
    """The demo scene: 10^3 grid, mvScale 3, N 32, <=60 particles of size 4 from a 10-degree cone
    emitter at (0,5,11.2) rotated 180 degrees about Y; camera (-10,0,-20), 1024x768, fov 60."""
    sc = _base(10, 32, 3.0, 1024, 768, 64)
    sc["camera"]["position"] = (-10.0, 0.0, -20.0)
    sc["emitter"] = {"position": (0.0, 5.0, 11.2), "rotation": (0.0, 1.0, 0.0, 0.0)}
    rng = np.random.default_rng(1000)
    count = 60
    age = rng.uniform(0.0, 6.0, size=count)                  # seconds since emission, lifetime 6
    dist = 3.0 * age                                          # start speed 3
    ang = rng.uniform(0.0, 2 * math.pi, size=count)
    rad = (0.5 * np.sqrt(rng.uniform(0.0, 1.0, size=count))) + dist * math.tan(math.radians(10.0)) * rng.uniform(0.0, 1.0, size=count)
    p = np.empty((count, 7), dtype=np.float32)
    p[:, 0] = (rad * np.cos(ang)).astype(np.float32)
    p[:, 1] = (rad * np.sin(ang)).astype(np.float32)
    p[:, 2] = dist.astype(np.float32)
    p[:, 3] = 4.0
    p[:, 4] = rng.uniform(0.0, 180.0, size=count).astype(np.float32)
    p[:, 5] = (6.0 - age).astype(np.float32)
    p[:, 6] = 6.0
    sc["particles"] = p
    sc["name"] = "ref-defaults"
    sc["seed"] = 1000
    return sc


def apply_scene(engine, sc):
    """Push light, cubemap and (no) depth map of a scene into an Engine."""
    engine.set_light(sc["light"]["position"], sc["light"]["rotation"], sc["gridCenter"])
    engine.set_displacement_cubemap(sc.get("cubemap", None) if sc.get("cubemap", None) is not None
                                    else load_displacement_cubemap())
    engine.set_light_depth_map(sc.get("depthMap", None))
