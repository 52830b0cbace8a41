"""The step before the path: a deterministic stand-in for the demo scene's Unity
ParticleSystem (Assets/Volumetric_Particle_System.unity:2259-2900), so that frames can be animated
headlessly and fed to VolumetricParticleRenderer.OnPostRender exactly like ParticleSystem.GetParticles()
feeds the reference (VPR.cs:412-413).

Unity's particle system is closed source; what is restated here are the serialized parameters of the
demo's emitter and their documented meaning, with numpy's PCG64 in place of Unity's generator:

    lengthInSec 6, looping, prewarm                      scene:2264-2269
    startLifetime 6, startSpeed 3, startSize 4           scene:2275, 2308, 2435
    startRotation: curve of constant 0                   scene:2467-2494
    maxNumParticles 60                                   scene:2497
    shape: cone (type 4), radius 0.5, angle 10 degrees   scene:2498-2511
    emission: 10 particles/s + a burst of 30 at t = 0    scene:2512-2556
    rotation over lifetime: 0.0698131695 rad/s (4 deg/s) scene:2587-2590
    gravity 0, no velocity / force modules               scene:2495, 2759-2897
    simulation space: local (moveWithTransform 1)        scene:2271

Size over lifetime (scene:2557-2586) is procedural in Unity 5 — `Particle.size` stays the start size — so
`size` is constant, which is what the reference's binning and fill read (VPR.cs:425,583).

`GetParticles()` returns the (n, 7) float32 array of include/vpe.h's VpeParticle: position (emitter
local), size, rotation in degrees, remaining lifetime, startLifetime.
"""
import math

import numpy as np


class ConeEmitter:
    def __init__(self, seed=0, duration=6.0, looping=True, prewarm=True, start_lifetime=6.0, start_speed=3.0,
                 start_size=4.0, max_particles=60, radius=0.5, angle_degrees=10.0, rate=10.0,
                 bursts=((0.0, 30),), angular_velocity_degrees=math.degrees(0.0698131695)):
        self.rng = np.random.default_rng(seed)
        self.duration, self.looping = float(duration), bool(looping)
        self.start_lifetime, self.start_speed, self.start_size = float(start_lifetime), float(start_speed), float(start_size)
        self.max_particles = int(max_particles)
        self.radius, self.angle = float(radius), math.radians(float(angle_degrees))
        self.rate, self.bursts = float(rate), tuple((float(t), int(c)) for t, c in bursts)
        self.angular_velocity = float(angular_velocity_degrees)
        self.time = 0.0              # seconds since the system started playing
        self._emit_debt = 0.0        # fractional particles owed by the rate
        self.p0 = np.zeros((0, 3))   # birth position, velocity, birth time of the live particles (float64 state)
        self.vel = np.zeros((0, 3))
        self.born = np.zeros((0,))
        if self.bursts:
            self._fire_bursts(-1e-9, 0.0)   # bursts scheduled at t = 0 fire when playback starts
        if prewarm and looping:
            for _ in range(180):            # one full cycle in 1/30 s steps, as Unity's prewarm does
                self.Simulate(self.duration / 180.0)

    # -- emission -------------------------------------------------------------------------------
    def _emit(self, count, t_birth):
        room = self.max_particles - self.born.shape[0]
        count = int(min(count, max(room, 0)))   # a full system drops the emission (maxNumParticles)
        if count <= 0:
            return
        # cone: uniform over the base disc; the direction tilts outwards in proportion to the radial position
        r = self.radius * np.sqrt(self.rng.random(count))
        phi = 2.0 * np.pi * self.rng.random(count)
        tilt = self.angle * (r / self.radius if self.radius > 0 else 0.0)
        pos = np.stack([r * np.cos(phi), r * np.sin(phi), np.zeros(count)], axis=1)
        d = np.stack([np.sin(tilt) * np.cos(phi), np.sin(tilt) * np.sin(phi), np.cos(tilt)], axis=1)
        self.p0 = np.concatenate([self.p0, pos])
        self.vel = np.concatenate([self.vel, d * self.start_speed])
        self.born = np.concatenate([self.born, np.full(count, float(t_birth))])

    def _fire_bursts(self, t0, t1):
        """Bursts whose cycle time lies in (t0, t1]."""
        for (bt, cnt) in self.bursts:
            if self.looping:
                k0 = math.floor((t0 - bt) / self.duration)
                k1 = math.floor((t1 - bt) / self.duration)
                for k in range(k0 + 1, k1 + 1):
                    if bt + k * self.duration >= 0.0:
                        self._emit(cnt, bt + k * self.duration)
            elif t0 < bt <= t1:
                self._emit(cnt, bt)

    # -- Unity-like surface -----------------------------------------------------------------------
    def Simulate(self, dt):
        """Advance by dt seconds (≙ one Update of the particle system)."""
        dt = float(dt)
        if dt <= 0.0:
            return
        t0, t1 = self.time, self.time + dt
        # age and retire first, then emit: a particle that dies in this step frees its slot for the step's emission
        alive = (t1 - self.born) < self.start_lifetime
        self.p0, self.vel, self.born = self.p0[alive], self.vel[alive], self.born[alive]
        emitting = self.looping or t0 < self.duration
        if emitting:
            self._fire_bursts(t0, t1)
            self._emit_debt += self.rate * dt
            n = int(math.floor(self._emit_debt))
            if n > 0:
                self._emit_debt -= n
                # births spread evenly over the step
                for i in range(n):
                    self._emit(1, t0 + (i + 1) * dt / n)
        self.time = t1

    @property
    def particleCount(self):
        return int(self.born.shape[0])

    def GetParticles(self):
        """≙ ParticleSystem.GetParticles(Particle[]) (VPR.cs:412-413): (n, 7) float32."""
        age = self.time - self.born
        out = np.empty((self.born.shape[0], 7), dtype=np.float32)
        out[:, 0:3] = self.p0 + self.vel * age[:, None]
        out[:, 3] = self.start_size
        out[:, 4] = np.mod(self.angular_velocity * age, 360.0)
        out[:, 5] = self.start_lifetime - age
        out[:, 6] = self.start_lifetime
        return out
