"""The host-side mirrors of the reference's VolumetricParticleRenderer (Python: vpe_b200.renderer,
C++: host/cpp) — frame structure (fill every updateInterval frames, march every frame), setters,
and that both drive the C-ABI to the same image.  CPU: the host logic runs over the oracle library
(the checker); GPU (-m gpu): the same over libvpe_cuda.so against the golden fixture."""
import os
import subprocess

import numpy as np
import pytest

import vpe_b200
from vpe_b200 import scenes
from vpe_b200.renderer import VolumetricParticleRenderer
from parity import RTOL, max_rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_renderer(lib, sc):
    r = VolumetricParticleRenderer(lib)
    r.numMetavoxelsX, r.numMetavoxelsY, r.numMetavoxelsZ = sc["grid"]
    r.mvScale = (sc["mvScale"],) * 3
    r.numVoxelsInMetavoxel = sc["numVoxels"]
    r.dirLight, r.particleSys, r.gridCenter = sc["light"], sc["emitter"], sc["gridCenter"]
    r.Start()
    return r


def check_frame_structure(lib):
    sc = scenes.make_scene("cfg1", image=(48, 48))
    r = make_renderer(lib, sc)
    assert r.updateInterval == 2 and not r.fadeOutParticles
    moved = sc["particles"].copy()
    moved[:, 0] += 0.25
    f0 = r.OnPostRender(sc["particles"], sc["camera"])       # frame 0: fill + march
    covered0 = r.numMetavoxelsCovered
    f1 = r.OnPostRender(moved, sc["camera"])                 # frame 1: march only (VPR.cs:186) -> same volume
    assert np.array_equal(f0, f1) and r.numMetavoxelsCovered == covered0
    f2 = r.OnPostRender(moved, sc["camera"])                 # frame 2: refilled with the moved particles
    assert not np.array_equal(f0, f2)
    r.SetUpdateInterval(1)
    r.SetParticleOpacityFactor(0.08)
    f3 = r.OnPostRender(moved, sc["camera"])
    assert f3[..., 3].sum() > f2[..., 3].sum()               # denser particles cover more
    r.SetRayMarchSteps(16)
    _, samples16 = r.RenderMetavoxels(sc["camera"], show_samples=True)
    r.SetRayMarchSteps(64)
    _, samples64 = r.RenderMetavoxels(sc["camera"], show_samples=True)
    assert 3.5 < samples64.sum() / samples16.sum() < 4.5
    # the rest of the frame (VPR.cs:184,204,210): occluders seen from the light, scene depth, blend onto the scene
    import frame_scenes
    h, w = sc["camera"]["height"], sc["camera"]["width"]
    scene_rt = np.zeros((h, w, 4), dtype=np.float32)
    scene_rt[..., 2] = 0.5
    full = r.OnPostRender(moved, sc["camera"], occluders=frame_scenes.occluders(sc), mainSceneRT=scene_rt,
                          sceneDepth=frame_scenes.scene_depth(sc))
    assert full.shape == (h, w, 4) and (full[:h // 4, :w // 4, 2] == 0.5).all()   # hidden behind the scene: untouched
    assert full[..., 0].sum() < r.engine.composite_scene(f3, scene_rt)[..., 0].sum()   # shadowed and partly hidden
    r.ShowRayMarchBlendFunc(True)
    view = r.RenderMetavoxels(sc["camera"])
    assert set(np.unique(view)) <= {0.0, 0.5, 1.0}
    r.ShowRayMarchBlendFunc(False)
    r.engine.render_light_depth_map(np.zeros((0, 3, 3), dtype=np.float32))
    r.SetGridScale(1.5)                                      # re-places the metavoxels (VPR.cs:1059-1063)
    assert np.allclose(r.engine.read_metavoxel_position(5, 4, 4) - r.engine.read_metavoxel_position(4, 4, 4), [1.5, 0, 0], atol=1e-5)
    return f0


def test_metavoxel_wire_grid_over_the_oracle():
    """DrawMetavoxelGrid (VPR.cs:962-1033): 12 edges per covered metavoxel, each of length mvScale, centred on mPos."""
    from oracle_lib import load_oracle
    sc = scenes.make_scene("cfg1", image=(32, 32))
    r = make_renderer(load_oracle(), sc)
    r.FillMetavoxels(sc["particles"])
    lines = r.DrawMetavoxelGrid()
    assert lines.shape == (r.numMetavoxelsCovered, 12, 2, 3) and r.numMetavoxelsCovered == 218
    length = np.linalg.norm(lines[:, :, 1] - lines[:, :, 0], axis=-1)
    assert np.allclose(length, sc["mvScale"], atol=1e-5)
    centre = lines.reshape(lines.shape[0], 24, 3).mean(axis=1)
    covered = [(x, y, z) for z in range(8) for y in range(8) for x in range(8) if r.engine.read_particle_list(x, y, z).shape[0]]
    want = np.array([r.engine.read_metavoxel_position(*c) for c in covered])
    assert np.allclose(centre, want, atol=1e-5)


def build_cli(tmp_path):
    exe = str(tmp_path / "vpe_cli")
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "host", "cpp", "vpe_cli.cpp"), "-ldl"], check=True)
    return exe


def run_cli(exe, lib_path, tmp_path, frames, occluders=None):
    sc = scenes.make_scene("cfg1", image=(48, 48))
    pfile, out = str(tmp_path / "particles.f32"), str(tmp_path / "out.rgba")
    sc["particles"].astype(np.float32).tofile(pfile)
    cube = os.path.join(scenes.ASSET_DIR, "displacement_r8.bin")
    extra = []
    if occluders is not None:
        tfile = str(tmp_path / "occluders.f32")
        np.ascontiguousarray(occluders, dtype=np.float32).tofile(tfile)
        extra = [tfile]
    res = subprocess.run([exe, lib_path, cube, pfile, "8", "8", "1.0", "48", "48", str(sc["camera"]["position"][2]), str(frames), out] + extra,
                         check=True, capture_output=True, text=True)
    summary = dict(kv.split("=") for kv in res.stdout.split())
    return np.fromfile(out, dtype=np.float32).reshape(48, 48, 4), summary


def test_python_mirror_frame_structure_over_the_oracle():
    from oracle_lib import load_oracle
    check_frame_structure(load_oracle())


def test_cpp_mirror_over_the_oracle(tmp_path):
    from oracle_lib import load_oracle, ORACLE_LIB
    img, summary = run_cli(build_cli(tmp_path), ORACLE_LIB, tmp_path, frames=5)
    assert summary["backend"] == "oracle" and summary["fills"] == "3"      # frames 0, 2, 4 (updateInterval 2)
    sc = scenes.make_scene("cfg1", image=(48, 48))
    r = make_renderer(load_oracle(), sc)
    want = r.OnPostRender(sc["particles"], sc["camera"])
    assert np.array_equal(img, want)
    assert int(summary["covered"]) == 218 and int(summary["pairs"]) == 457


def _whole_frame_through_python(lib, sc):
    import frame_scenes
    r = make_renderer(lib, sc)
    r.particlesRT_8bit = True
    scene_rt = np.zeros((48, 48, 4), dtype=np.float32)
    scene_rt[..., 2], scene_rt[..., 3] = 0.5, 1.0
    return r.OnPostRender(sc["particles"], sc["camera"], occluders=frame_scenes.occluders(sc), mainSceneRT=scene_rt)


def test_cpp_mirror_whole_frame_over_the_oracle(tmp_path):
    """The C++ mirror's RenderLightDepthMap / SetMarchOptions / CompositeParticles (VPR.cs:184,210,228) against the
    Python mirror, both over the oracle library."""
    import frame_scenes
    from oracle_lib import load_oracle, ORACLE_LIB
    sc = scenes.make_scene("cfg1", image=(48, 48))
    img, summary = run_cli(build_cli(tmp_path), ORACLE_LIB, tmp_path, frames=1, occluders=frame_scenes.occluders(sc))
    want = _whole_frame_through_python(load_oracle(), sc)
    assert np.array_equal(img, want)
    assert np.array_equal(np.round(img * 255), img * 255) and (img[..., 3] >= 1.0).all()   # 8-bit, opaque scene below
    assert img[..., 0].max() > 0.05 and abs(float(np.median(img[..., 2])) - 128.0 / 255) < 0.05


@pytest.mark.gpu
def test_cpp_mirror_whole_frame_on_cuda(tmp_path):
    import frame_scenes
    sc = scenes.make_scene("cfg1", image=(48, 48))
    img, summary = run_cli(build_cli(tmp_path), vpe_b200.CUDA_LIB_PATH, tmp_path, frames=1, occluders=frame_scenes.occluders(sc))
    assert summary["backend"] == "cuda"
    want = _whole_frame_through_python(None, sc)
    assert np.array_equal(img, want)
    from oracle_lib import load_oracle
    ref = _whole_frame_through_python(load_oracle(), sc)
    diff = np.abs(img - ref)
    assert diff.max() <= 1.0 / 255 + 1e-6 and (diff > 1e-6).mean() < 5e-3


@pytest.mark.gpu
def test_python_mirror_frame_structure_on_cuda():
    check_frame_structure(None)


@pytest.mark.gpu
def test_cpp_mirror_on_cuda(tmp_path):
    img, summary = run_cli(build_cli(tmp_path), vpe_b200.CUDA_LIB_PATH, tmp_path, frames=3)
    assert summary["backend"] == "cuda" and summary["fills"] == "2"
    assert int(summary["covered"]) == 218 and int(summary["pairs"]) == 457
    sc = scenes.make_scene("cfg1", image=(48, 48))
    r = make_renderer(None, sc)
    want = r.OnPostRender(sc["particles"], sc["camera"])
    assert max_rel_err(img, want) <= 1e-6
    from oracle_lib import load_oracle
    ref = make_renderer(load_oracle(), sc)
    assert max_rel_err(img, ref.OnPostRender(sc["particles"], sc["camera"])) <= RTOL
