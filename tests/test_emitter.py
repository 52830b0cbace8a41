"""The demo scene's particle emitter restated (vpe_b200.emitter) and an animated run of the
host mirror over it: fill every updateInterval frames, march every frame (VPR.cs:186,207)."""
import numpy as np

from vpe_b200.emitter import ConeEmitter
from vpe_b200.renderer import VolumetricParticleRenderer
from vpe_b200 import scenes


def test_steady_state_of_the_demo_emitter():
    e = ConeEmitter(seed=3)
    assert e.particleCount == 60              # prewarmed: 10/s x 6 s lifetime = maxNumParticles (scene:2497,2512)
    for _ in range(200):
        e.Simulate(1.0 / 30.0)
        assert 50 <= e.particleCount <= 60
    p = e.GetParticles()
    assert p.dtype == np.float32 and p.shape[1] == 7
    assert (p[:, 3] == 4.0).all() and (p[:, 6] == 6.0).all()
    assert (p[:, 5] > 0).all() and (p[:, 5] <= 6.0).all()
    age = 6.0 - p[:, 5]
    assert np.allclose(p[:, 4], np.mod(4.0 * age, 360.0), atol=1e-3)   # 0.0698 rad/s
    # inside the cone: radius 0.5 at the base, opening 10 degrees, speed 3
    rad = np.hypot(p[:, 0], p[:, 1])
    assert (rad <= 0.5 + p[:, 2] * np.tan(np.radians(10.0)) + 1e-4).all()
    assert np.allclose(np.linalg.norm(p[:, :3] - np.c_[p[:, 0], p[:, 1], np.zeros(len(p))] * 0, axis=1) >= 0, True)
    assert (p[:, 2] <= 3.0 * age + 1e-4).all() and (p[:, 2] >= 3.0 * age * np.cos(np.radians(10.0)) - 1e-4).all()


def test_deterministic_and_seeded():
    a, b, c = ConeEmitter(seed=1), ConeEmitter(seed=1), ConeEmitter(seed=2)
    for e in (a, b, c):
        for _ in range(10):
            e.Simulate(0.05)
    assert np.array_equal(a.GetParticles(), b.GetParticles())
    assert not np.array_equal(a.GetParticles(), c.GetParticles())


def test_one_shot_system_runs_dry():
    e = ConeEmitter(seed=0, looping=False, prewarm=False, max_particles=1000)
    assert e.particleCount == 30              # the burst at t = 0
    e.Simulate(3.0)
    assert e.particleCount == 60
    e.Simulate(3.0)                           # emission stops at lengthInSec; the burst particles die at 6 s
    e.Simulate(6.5)
    assert e.particleCount == 0


def test_animated_frames_through_the_host_mirror():
    """Demo defaults (10^3 grid, scale 3, emitter at (0,5,11.2) facing -Z... scene:8965-8988) at a small
    voxel count over the oracle: the volume follows the emitter at every second frame only."""
    from oracle_lib import load_oracle
    r = VolumetricParticleRenderer(load_oracle())
    r.numVoxelsInMetavoxel = 8
    r.Start()
    em = ConeEmitter(seed=5)
    cam = {"position": (0.0, 0.0, -22.0), "rotation": (0.0, 0.0, 0.0, 1.0), "fovYDegrees": 60.0, "width": 64, "height": 48}
    frames = []
    for f in range(4):
        frames.append(r.OnPostRender(em.GetParticles(), cam))
        em.Simulate(0.5)
    assert frames[0][..., 3].max() > 0.05 and r.numParticlesEmitted >= 50
    assert np.array_equal(frames[0], frames[1]) and np.array_equal(frames[2], frames[3])   # updateInterval = 2
    assert not np.array_equal(frames[1], frames[2])
