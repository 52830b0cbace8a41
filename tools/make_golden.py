#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the CPU oracle (oracle/libvpe_ref.so).

The reference holds no golden vectors for this path: these are outputs of OUR
oracle, pinned so that (a) the oracle cannot drift silently (tests/test_golden.py, CPU) and (b) the
CUDA engine can be checked on the GPU box against committed numbers, including a full-size volume
hash for cfg2 that the scalar oracle would be too slow to recompute there on every run.

  python tools/make_golden.py            # rewrites tests/golden/*.npz
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from vpe_b200 import scenes  # noqa: E402
from oracle_lib import oracle_engine  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# name -> (scene name, image override, pixel subset size or None = full image, bin mode)
CASES = {
    "cfg1": ("cfg1", None, None, 0),
    "cfg1_exact_bins": ("cfg1", None, None, 1),
    "ref_defaults": ("ref-defaults", (160, 120), None, 0),
    "cfg2_subset": ("cfg2", None, 1500, 0),
}


def case_scene(name):
    scene_name, image, subset, bin_mode = CASES[name]
    sc = scenes.make_scene(scene_name)
    if image is not None:
        sc["camera"]["width"], sc["camera"]["height"] = image
    return sc, subset, bin_mode


def subset_pixels(sc, count):
    rng = np.random.default_rng(sc["seed"] + 77)
    w, h = sc["camera"]["width"], sc["camera"]["height"]
    return np.sort(rng.choice(w * h, size=count, replace=False)).astype(np.int32)


def volume_digest(engine):
    """sha256 over every covered brick (z-major, then y, x), each prefixed by its flat index; plus the
    particle lists as CSR."""
    gx, gy, gz = engine.grid
    h = hashlib.sha256()
    offsets, indices = [0], []
    for z in range(gz):
        for y in range(gy):
            for x in range(gx):
                lst = engine.read_particle_list(x, y, z)
                indices.append(lst)
                offsets.append(offsets[-1] + len(lst))
                if len(lst):
                    b = engine.read_brick(x, y, z)
                    h.update(np.int32((z * gy + y) * gx + x).tobytes())
                    h.update(np.ascontiguousarray(b).tobytes())
    return h.hexdigest(), np.asarray(offsets, dtype=np.int32), (np.concatenate(indices).astype(np.int32) if indices else np.zeros(0, np.int32))


def run_case(name, engine_factory):
    sc, subset, bin_mode = case_scene(name)
    e = engine_factory(sc, binMode=bin_mode)
    scenes.apply_scene(e, sc)
    e.fill(sc["particles"], sc["emitter"])
    digest, offsets, indices = volume_digest(e)
    out = {"volume_sha256": np.frombuffer(bytes.fromhex(digest), dtype=np.uint8), "list_offsets": offsets, "list_indices": indices,
           "sheet": e.read_light_sheet()}
    if subset is None:
        img, smp = e.march(sc["camera"])
    else:
        pix = subset_pixels(sc, subset)
        img, smp = e.march_pixels(sc["camera"], pix)
        out["pixels"] = pix
    st = e.stats()
    out.update(rgba=img, samples=smp, covered=np.int64(st["numMetavoxelsCovered"]), pairs=np.int64(st["numParticlePairs"]),
               ray_samples=np.int64(st["raySamples"]), z_boundary=np.int64(st["zBoundary"]))
    return out


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    for name in CASES:
        out = run_case(name, lambda sc, **kw: oracle_engine(sc, **kw))
        path = os.path.join(GOLDEN, name + ".npz")
        np.savez_compressed(path, **out)
        print("%-16s covered=%d pairs=%d ray_samples=%d -> %s (%d KB)" % (
            name, out["covered"], out["pairs"], out["ray_samples"], os.path.relpath(path, ROOT), os.path.getsize(path) // 1024))


if __name__ == "__main__":
    main()
