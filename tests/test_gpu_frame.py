"""GPU parity for what surrounds the path in the reference's frame: the CUDA
engine against the oracle through the same C-ABI calls. Depth map, composite and debug views: bit-exact;
images: the RGBA tolerance of tests/parity.py."""
import numpy as np
import pytest

import vpe_b200
from vpe_b200 import scenes
from oracle_lib import oracle_engine
from parity import RTOL, rel_err
import frame_scenes

pytestmark = pytest.mark.gpu


def _pair(sc):
    gpu = vpe_b200.engine_for_scene(None, sc)
    ref = oracle_engine(sc)
    for e in (gpu, ref):
        scenes.apply_scene(e, sc)
    return gpu, ref


def test_light_depth_map_rasteriser_and_the_fill_that_reads_it():
    sc = scenes.make_scene("cfg1", image=(96, 96))
    gpu, ref = _pair(sc)
    rng = np.random.default_rng(8)
    g = sc["grid"][0] * sc["mvScale"]
    soup = rng.uniform(-0.6 * g, 0.6 * g, (300, 3, 3)).astype(np.float32)   # triangle soup: every winding, clipping, overlaps
    tris = np.concatenate([frame_scenes.occluders(sc), soup[:40] * np.float32(0.3) - np.float32(0.4 * g) * np.asarray(frame_scenes.light_frame(sc)[2], dtype=np.float32)])
    for e in (gpu, ref):
        e.render_light_depth_map(tris)
    dg, dr = gpu.read_light_depth_map(), ref.read_light_depth_map()
    assert (dr < 1.0).mean() > 0.05
    assert np.array_equal(dg, dr)
    for e in (gpu, ref):
        e.fill(sc["particles"], sc["emitter"])
    assert np.array_equal(gpu.read_light_sheet(), ref.read_light_sheet())
    for (x, y, z) in [(1, 2, 3), (3, 4, 5), (2, 2, 2), (5, 3, 6), (4, 4, 1)]:
        a, b = gpu.read_brick(x, y, z), ref.read_brick(x, y, z)
        assert (a is None) == (b is None)
        if a is not None:
            assert np.array_equal(a, b)
    img_g, smp_g = gpu.march(sc["camera"])
    img_r, smp_r = ref.march(sc["camera"])
    assert np.array_equal(smp_g, smp_r) and float(rel_err(img_g, img_r).max()) <= RTOL
    # an empty occluder list clears the map
    gpu.render_light_depth_map(np.zeros((0, 3, 3), dtype=np.float32))
    assert (gpu.read_light_depth_map() == 1.0).all()


def test_composite_over_the_scene_matches_the_oracle():
    sc = scenes.make_scene("cfg1")
    gpu, ref = _pair(sc)
    rng = np.random.default_rng(4)
    p = rng.uniform(0, 1.2, (270, 480, 4)).astype(np.float32)
    s = rng.uniform(0, 1, (270, 480, 4)).astype(np.float32)
    for fmt in (0, 1):
        assert np.array_equal(gpu.composite_scene(p, s, fmt), ref.composite_scene(p, s, fmt))


def test_march_options_match_the_oracle():
    sc = scenes.make_scene("cfg1", image=(128, 96))
    sc["camera"]["position"] = (0.3, 0.2, -3.0)
    gpu, ref = _pair(sc)
    for e in (gpu, ref):
        e.fill(sc["particles"], sc["emitter"])
    depth = frame_scenes.scene_depth(sc)
    base_g, _ = gpu.march(sc["camera"])
    for kw in (dict(debug_mode=1), dict(debug_mode=2), dict(debug_mode=3), dict(scene_depth=depth),
               dict(debug_mode=3, scene_depth=depth), dict(target_format=1), dict(target_format=1, scene_depth=depth)):
        for e in (gpu, ref):
            e.set_march_options(**kw)
        img_g, smp_g = gpu.march(sc["camera"])
        img_r, smp_r = ref.march(sc["camera"])
        assert np.array_equal(smp_g, smp_r), kw
        assert float(img_r[..., 3].max()) > 0.2, kw
        if kw.get("debug_mode") in (1, 2):
            assert np.array_equal(img_g, img_r), kw
        elif kw.get("target_format") == 1 or kw.get("debug_mode") == 3:
            # quantisation / band edges turn a last-bit difference into one step: allow a step on a few pixels
            diff = np.abs(img_g - img_r)
            assert diff.max() <= 1.0 / 255 + 1e-6 and (diff > 1e-6).mean() < 2e-3, kw
        else:
            assert float(rel_err(img_g, img_r).max()) <= RTOL, kw
    for e in (gpu, ref):
        e.set_march_options()
    again, _ = gpu.march(sc["camera"])
    assert np.array_equal(again, base_g)   # options off: the fast kernel again, same image


def test_march_options_are_refused_for_slab_partials():
    import torch
    from vpe_b200 import slabs
    sc = scenes.make_scene("cfg1", image=(64, 64))
    eng = slabs.CudaSlabEngine(sc, 0, 1, 0)
    eng.fill_prepare(sc["particles"], sc["emitter"])
    eng.fill_region(0, 8, 0, 8)
    eng.eng.set_march_options(target_format=1)
    with pytest.raises(vpe_b200.VpeError) as e:
        eng.march_partial(sc["camera"])
    assert e.value.code == vpe_b200._abi.VPE_E_UNSUPPORTED


def test_animated_emitter_frames_on_cuda_match_the_oracle():
    """SURVEY §8f row 3 on the GPU: the demo scene's cone emitter (vpe_b200.emitter, scene:2259-2900) drives the CUDA engine
    and the oracle through the host mirror at the reference's defaults (10^3 grid, mvScale 3, N = 32, fill every second
    frame, VPR.cs:186): same particle lists, bit-identical light sheet, images within the RGBA tolerance, frame by frame."""
    from vpe_b200.emitter import ConeEmitter
    from vpe_b200.renderer import VolumetricParticleRenderer
    from oracle_lib import load_oracle
    cam = {"position": (-10.0, 0.0, -20.0), "rotation": (0.0, 0.2164396, 0.0, 0.976296), "fovYDegrees": 60.0, "width": 160, "height": 120}
    gpu, ref = VolumetricParticleRenderer(), VolumetricParticleRenderer(load_oracle())
    for r in (gpu, ref):
        r.Start()
    em = ConeEmitter(seed=9)
    lit = 0
    for frame in range(4):
        p = em.GetParticles()
        a, b = gpu.OnPostRender(p, cam), ref.OnPostRender(p, cam)
        sg, sr = gpu.engine.stats(), ref.engine.stats()
        assert sg["numParticlePairs"] == sr["numParticlePairs"] and sg["numMetavoxelsCovered"] == sr["numMetavoxelsCovered"] > 0
        assert sg["raySamples"] == sr["raySamples"]
        assert np.array_equal(gpu.engine.read_light_sheet(), ref.engine.read_light_sheet())
        assert float(rel_err(a, b).max()) <= RTOL
        lit += int(b[..., 3].max() > 0.05)
        em.Simulate(0.4)
    assert lit == 4 and gpu.numParticlesEmitted == ref.numParticlesEmitted >= 50
