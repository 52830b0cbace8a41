"""Time the sweep kernel alone on one GPU: cfg3, whole grid as one slab, density pass then the sweep (16 B/voxel).
usage: sweep_kernel_bench.py [cfg] [tma|reg]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import vpe_b200
from vpe_b200 import scenes

sc = scenes.make_scene(sys.argv[1] if len(sys.argv) > 1 else "cfg3")
e = vpe_b200.engine_for_scene(None, sc)
mode = sys.argv[2] if len(sys.argv) > 2 else "tma"
e.set_debug_options(no_tma_sweep=(mode == "reg"))
scenes.apply_scene(e, sc)
gx, gy, gz = e.grid
ts, td = [], []
for it in range(5):
    e.fill_prepare(sc["particles"], sc["emitter"])
    e.fill_density()
    td.append(e.stats()["fillKernelMs"])
    e.fill_sweep_region(0, gx, 0, gy)
    st = e.stats()
    ts.append(st["fillKernelMs"])
vox = st["voxelsFilled"]
t = float(np.median(ts[1:]))
print(mode, "density %.3f ms   sweep %.3f ms  %.0f GB/s (16 B/voxel, %d voxels)" % (float(np.median(td[1:])), t, vox * 16 / t / 1e6, vox))
