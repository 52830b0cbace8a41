"""Print where the CUDA march differs most from the oracle (run on the GPU box)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import vpe_b200
from vpe_b200 import scenes
from oracle_lib import oracle_engine

name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
sc = scenes.make_scene(name)
if len(sys.argv) > 2:
    sc["camera"]["width"], sc["camera"]["height"] = int(sys.argv[2]), int(sys.argv[3])
gpu = vpe_b200.engine_for_scene(None, sc)
ref = oracle_engine(sc)
for e in (gpu, ref):
    scenes.apply_scene(e, sc)
    e.fill(sc["particles"], sc["emitter"])
ig, sg = gpu.march(sc["camera"])
ir, sr = ref.march(sc["camera"])
print("samples equal:", np.array_equal(sg, sr))
ad = np.abs(ig - ir)
rel = ad / np.maximum(np.abs(ir), 1e-6)
print("max abs err", ad.max(), "max rel err", rel.max())
for thr in (1e-6, 1e-5, 1e-4, 1e-3, 1e-2):
    m = np.abs(ir) >= thr
    print("  |ref| >= %g: n=%d max rel %.3g" % (thr, m.sum(), rel[m].max() if m.any() else 0))
from parity import rel_err
print("parity metric (tests/parity.py):", rel_err(ig, ir).max())
idx = np.argsort(rel.ravel())[::-1][:4]
for i in idx:
    y, x, c = np.unravel_index(i, rel.shape)
    print("  pix (%d,%d) ch %d gpu %.9g ref %.9g rel %.3g samples %d" % (x, y, c, ig[y, x, c], ir[y, x, c], rel[y, x, c], sr[y, x]))
idx = np.argsort(ad.ravel())[::-1][:5]
for i in idx:
    y, x, c = np.unravel_index(i, rel.shape)
    print("  ABS pix (%d,%d) ch %d gpu %.9g ref %.9g abs %.3g" % (x, y, c, ig[y, x, c], ir[y, x, c], ad[y, x, c]))
