"""What surrounds the path in the reference's frame, on the oracle (CPU):
light depth map rasterisation, march options (UNORM8 target, scene depth test, debug views) and the
composite over the scene. Closed forms and invariants; the GPU parity tests are in test_gpu_frame.py."""
import numpy as np
import pytest

from vpe_b200 import scenes
from oracle_lib import oracle_engine
import frame_scenes


def _engine(sc):
    e = oracle_engine(sc)
    scenes.apply_scene(e, sc)
    return e


def test_depth_map_of_a_plate_facing_the_light_is_its_linear_depth():
    sc = scenes.make_scene("cfg1")
    e = _engine(sc)
    right, up, fwd = frame_scenes.light_frame(sc)
    g = sc["grid"][0] * sc["mvScale"]
    centre = -1.5 * fwd   # 1.5 units before the grid centre, i.e. 198.5 from the light camera (VPR.cs:365)
    plate = frame_scenes.quad(centre, 0.25 * g * right, 0.25 * g * up)
    n = sc["grid"][0] * sc["numVoxels"]
    for tris in (plate, plate[:, ::-1, :]):
        e.render_light_depth_map(tris)
        d = e.read_light_depth_map()
        drawn = d < 1.0
        if drawn.any():
            break
    assert drawn.any(), "one of the two windings must be drawn (Cull Front)"
    e.render_light_depth_map(tris[:, ::-1, :])
    assert (e.read_light_depth_map() == 1.0).all(), "the other winding is culled"
    e.render_light_depth_map(tris)
    d = e.read_light_depth_map()
    # the plate spans the middle half of the map in both directions (orthographic extent = the grid, VPR.cs:340)
    ys, xs = np.nonzero(d < 1.0)
    assert xs.min() == n // 4 and xs.max() == 3 * n // 4 - 1 and ys.min() == n // 4 and ys.max() == 3 * n // 4 - 1
    expect = (198.5 - 0.3) / (1000.0 - 0.3)
    assert np.allclose(d[d < 1.0], expect, rtol=0, atol=2e-6)
    assert (d[d >= 1.0] == 1.0).all()


def test_nearest_back_face_wins_and_clipping():
    sc = scenes.make_scene("cfg1")
    e = _engine(sc)
    right, up, fwd = frame_scenes.light_frame(sc)
    g = sc["grid"][0] * sc["mvScale"]
    a = frame_scenes.quad(-1.0 * fwd, 0.3 * g * right, 0.3 * g * up)
    b = frame_scenes.quad(-3.0 * fwd, 0.1 * g * right, 0.1 * g * up)
    behind_camera = frame_scenes.quad(-250.0 * fwd, 0.3 * g * right, 0.3 * g * up)
    e.render_light_depth_map(a)
    flip = not (e.read_light_depth_map() < 1.0).any()
    tris = np.concatenate([a, b, behind_camera])
    if flip:
        tris = tris[:, ::-1, :]
    e.render_light_depth_map(tris)
    d = e.read_light_depth_map()
    n = d.shape[0]
    za, zb = (199.0 - 0.3) / 999.7, (197.0 - 0.3) / 999.7
    assert abs(d[n // 2, n // 2] - zb) < 2e-6           # ZTest Less: the nearer plate
    assert abs(d[n // 2 + n // 5, n // 2] - za) < 2e-6  # outside the small plate
    assert d.min() > 0.0                                 # the plate behind the light camera is clipped


def test_rendered_depth_map_shadows_the_volume():
    sc = scenes.make_scene("cfg1")
    lit = _engine(sc)
    lit.fill(sc["particles"], sc["emitter"])
    shadowed = _engine(sc)
    shadowed.render_light_depth_map(frame_scenes.occluders(sc))
    assert (shadowed.read_light_depth_map() < 1.0).mean() > 0.05
    shadowed.fill(sc["particles"], sc["emitter"])
    a, b = lit.read_light_sheet(), shadowed.read_light_sheet()
    assert (b >= a).all() and (b > a).any()      # Fill.shader:239-250: light stops propagating (and attenuating) in shadow
    img_lit, _ = lit.march(sc["camera"])
    img_sh, _ = shadowed.march(sc["camera"])
    assert img_sh[..., 0].sum() < img_lit[..., 0].sum()
    assert np.array_equal(img_sh[..., 3], img_lit[..., 3])   # coverage does not depend on light


def test_composite_over_the_scene():
    sc = scenes.make_scene("cfg1")
    e = _engine(sc)
    rng = np.random.default_rng(4)
    p = rng.uniform(0, 1, (50, 40, 4)).astype(np.float32)
    s = rng.uniform(0, 1, (50, 40, 4)).astype(np.float32)
    out = e.composite_scene(p, s)
    k = np.float32(1.0) - p[..., 3:4]
    assert np.array_equal(out[..., :3], p[..., :3] + s[..., :3] * k)   # Blend One OneMinusSrcAlpha
    assert np.array_equal(out[..., 3], p[..., 3] + s[..., 3])          # , One One
    out8 = e.composite_scene(p, s, target_format=1)
    assert np.array_equal(np.round(out8 * 255), out8 * 255) and out8.max() <= 1.0
    assert np.abs(out8 - np.clip(out, 0, 1)).max() <= 0.5 / 255 + 1e-6


def test_march_options_views_and_target():
    sc = scenes.make_scene("cfg1", image=(96, 96))
    e = _engine(sc)
    e.fill(sc["particles"], sc["emitter"])
    img, smp = e.march(sc["camera"])
    hit = smp > 0
    # UNORM8 target: every stored value is a multiple of 1/255, and close to the float image
    e.set_march_options(target_format=1)
    img8, smp8 = e.march(sc["camera"])
    assert np.array_equal(smp8, smp)
    assert np.array_equal(np.round(img8 * 255), img8 * 255)
    assert np.abs(img8 - img).max() < 0.1
    # blend-function view: yellow where only OVER fragments, cyan where only UNDER (all opaque: the last drawn wins)
    e.set_march_options(debug_mode=2)
    dbg, dsmp = e.march(sc["camera"])
    assert not dsmp.any()   # the shader returns before the sample loop (March.shader:170-181)
    cols = {tuple(c) for c in dbg.reshape(-1, 4).round(3).tolist()}
    assert cols <= {(0.0, 0.0, 0.0, 0.0), (0.5, 0.5, 0.0, 1.0), (0.0, 0.5, 0.5, 1.0), (0.5, 0.5, 0.5, 1.0)} or len(cols) < 12
    # draw-order view: opaque fragments of one channel
    e.set_march_options(debug_mode=1)
    order, _ = e.march(sc["camera"])
    lit = order[..., 3] > 0
    assert lit.any() and (order[lit][:, 3] == 1.0).all()
    assert ((order[lit][:, :3] > 0).sum(axis=1) <= 1).all()
    # sample-count view: blended bands with alpha 0.5 per fragment
    e.set_march_options(debug_mode=3)
    bands, bsmp = e.march(sc["camera"])
    assert np.array_equal(bsmp, smp)
    assert (bands[..., 2] == 0).all() and bands[hit][:, 3].min() >= 0.5
    e.set_march_options()
    again, _ = e.march(sc["camera"])
    assert np.array_equal(again, img)


def test_scene_depth_test_drops_fragments_behind_the_scene():
    sc = scenes.make_scene("cfg1", image=(96, 96))
    e = _engine(sc)
    e.fill(sc["particles"], sc["emitter"])
    img, smp = e.march(sc["camera"])
    far = np.full((96, 96), 3.0e38, dtype=np.float32)
    e.set_march_options(scene_depth=far)
    same, ssmp = e.march(sc["camera"])
    assert np.array_equal(same, img) and np.array_equal(ssmp, smp)
    e.set_march_options(scene_depth=np.zeros((96, 96), dtype=np.float32))
    none, nsmp = e.march(sc["camera"])
    assert not none.any() and not nsmp.any()
    depth = frame_scenes.scene_depth(sc)
    e.set_march_options(scene_depth=depth)
    part, psmp = e.march(sc["camera"])
    assert (psmp <= smp).all() and (psmp < smp).any()
    assert not psmp[:24, :24].any()                       # something right in front of the camera hides everything
    assert np.array_equal(psmp[40:, :40], smp[40:, :40])  # nothing in front of the scene there
    with pytest.raises(Exception):
        e.set_march_options(scene_depth=np.zeros((10, 10), dtype=np.float32))
        e.march(sc["camera"])
