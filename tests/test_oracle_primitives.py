"""The oracle's own primitives and invariants (CPU)."""
import numpy as np
import pytest

from vpe_b200 import scenes
from oracle_lib import load_oracle, oracle_engine


def test_f16_to_f32_exhaustive():
    lib = load_oracle()
    bits = np.arange(65536, dtype=np.uint16)
    want = bits.view(np.float16).astype(np.float32)
    got = np.array([lib.vpe_ref_f16_to_f32(int(b)) for b in bits], dtype=np.float32)
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan)
    assert np.array_equal(got[~nan].view(np.uint32), want[~nan].view(np.uint32))


def test_f32_to_f16_round_to_nearest_even():
    lib = load_oracle()
    rng = np.random.default_rng(5)
    vals = np.concatenate([
        rng.uniform(-70000, 70000, 20000), rng.uniform(-1, 1, 20000), rng.uniform(-1e-4, 1e-4, 20000),
        rng.uniform(-1e-7, 1e-7, 5000), [0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e9, -1e9, 2.0 ** -24, 2.0 ** -25, 1.5 * 2.0 ** -25],
        # exact ties between neighbouring halves
        (np.arange(1024, 2048, dtype=np.float64) + 0.5) * 2.0 ** -10]).astype(np.float32)
    with np.errstate(over="ignore"):
        want = vals.astype(np.float16).view(np.uint16)
    got = np.array([lib.vpe_ref_f32_to_f16(float(v)) for v in vals], dtype=np.uint16)
    assert np.array_equal(got, want)


def test_empty_and_degenerate_inputs():
    sc = scenes.make_scene("cfg1", image=(32, 32))
    e = oracle_engine(sc)
    scenes.apply_scene(e, sc)
    e.fill(np.zeros((0, 7), dtype=np.float32), sc["emitter"])         # empty particle list
    img, smp = e.march(sc["camera"])
    assert e.stats()["numMetavoxelsCovered"] == 0 and not img.any() and not smp.any()
    assert (e.read_light_sheet() == 1.0).all()
    far = sc["particles"][:4].copy()
    far[:, 0:3] += 1000.0                                              # particles outside the grid bin nowhere
    e.fill(far, sc["emitter"])
    assert e.stats()["numParticlePairs"] == 0


def test_image_invariants_and_sample_counts():
    sc = scenes.make_scene("cfg1", image=(64, 64))
    e = oracle_engine(sc)
    scenes.apply_scene(e, sc)
    e.fill(sc["particles"], sc["emitter"])
    img, smp = e.march(sc["camera"])
    assert (img[..., 3] >= 0).all() and (img[..., 3] <= 1).all()       # coverage alpha
    assert (img[..., :3] >= 0).all() and (img[..., :3] <= 1.0).all()   # premultiplied colour <= 0.4 + 0.2*ao
    assert int(smp.sum()) == e.stats()["raySamples"]
    # a ray crosses at most sqrt(3) * G metavoxel widths, 64 steps each (March.shader:221-224)
    assert smp.max() <= int(np.sqrt(3) * 8 * 64) + 8
    pix = np.array([0, 63, 64 * 32 + 32, 64 * 64 - 1], dtype=np.int32)
    rgba, s = e.march_pixels(sc["camera"], pix)
    assert np.array_equal(rgba, img.reshape(-1, 4)[pix]) and np.array_equal(s, smp.reshape(-1)[pix])


def test_single_particle_centred_in_a_metavoxel():
    """Closed form: one sphere centred on metavoxel (4,4,4) of an axis-aligned 8^3 grid. Every voxel
    whose centre lies within the 0.7*nd core gets density = opacityFactor exactly; voxels outside the
    sphere get 0; the column through the centre attenuates light by prod 1/(1+density)."""
    sc = scenes.make_scene("cfg1", image=(16, 16))
    sc["light"] = {"position": (0.0, 0.0, -20.0), "rotation": (0.0, 0.0, 0.0, 1.0)}
    sc["displacementScale"] = 0.0                                       # nd == 1: a plain sphere
    e = oracle_engine(sc)
    scenes.apply_scene(e, sc)
    c = e.read_metavoxel_position(4, 4, 4)
    p = np.array([[c[0], c[1], c[2], 0.9, 0.0, 6.0, 6.0]], dtype=np.float32)   # radius 0.45 < half a metavoxel
    e.fill(p, sc["emitter"])
    assert e.stats()["numMetavoxelsCovered"] >= 1
    b = e.read_brick(4, 4, 4).view(np.float16).astype(np.float64)       # [k][y][x][rgba]
    N, sb = 8, 8.0 / 6.0
    k, y, x = np.meshgrid(np.arange(N), np.arange(N), np.arange(N), indexing="ij")
    # voxel centre in metavoxel units: x,y centred, z at the slice's near face
    pos = np.stack([(x + 0.5 - N / 2) / N, (y + 0.5 - N / 2) / N, (k - N / 2) / N], -1) * sb
    r2 = (pos ** 2).sum(-1) / 0.9 ** 2                                  # |q|^2 in particle space (size 0.9)
    d2 = 4 * r2
    inside = r2 <= 0.25
    t = np.clip((d2 - 1.0) / (0.7 - 1.0), 0, 1)
    want = np.where(inside, t * t * (3 - 2 * t) * 0.04, 0.0)
    far_from_edges = (np.abs(r2 - 0.25) > 1e-4)
    assert np.allclose(b[..., 3][far_from_edges], want[far_from_edges], rtol=2.0 ** -10, atol=1e-7)
    assert (b[..., 3][d2 <= 0.69] == np.float16(0.04)).all() and (b[..., 3][r2 > 0.2501] == 0).all()
    # ambient occlusion term: colour = 0.4*T + 0.2*ao with ao = nd = 1 inside, 0 outside
    outside_lit = (~inside) & (k == 0)
    assert np.allclose(b[..., 0][outside_lit], 0.4, atol=2e-4)
    # light sheet under the particle: product of the blend factors of the interior slices
    sheet = e.read_light_sheet()[4 * N:5 * N, 4 * N:5 * N]
    dens32 = b[..., 3]
    T = np.ones((N, N))
    prop = T.copy()
    for kk in range(N):
        if kk < N - 1:
            prop = T.copy()
        T = T / (1 + dens32[kk])
    assert np.allclose(sheet, prop, rtol=2e-3)
    assert sheet.min() < 0.9 and sheet.max() == 1.0
