"""Slab engine adapter over the CPU oracle (TEST INFRASTRUCTURE): lets the CPU tests run the
product's multi-GPU host logic (vpe_b200.slabs.SlabRenderer) with gloo and host tensors."""
import ctypes as C

import numpy as np
import torch

from vpe_b200 import slabs
from vpe_b200.engine import _camera
from oracle_lib import load_oracle, oracle_engine


class OracleSlabEngine:
    def __init__(self, scene, rank, world):
        from vpe_b200 import scenes
        self.lib = load_oracle()
        self.scene = scene
        self.slab = slabs.slab_range(scene["grid"][2], world, rank)
        self.eng = oracle_engine(scene, slab=self.slab)
        scenes.apply_scene(self.eng, scene)
        self.grid, self.N = self.eng.grid, self.eng.N
        gx, gy, _ = self.grid
        self._sheet = torch.ones((gy * self.N, gx * self.N), dtype=torch.float32)

    def new_tensor(self, shape):
        return torch.empty(tuple(shape), dtype=torch.float32)

    def fill_prepare(self, particles, emitter):
        self.eng.fill_prepare(particles, emitter)
        self._sheet.fill_(1.0)

    def fill_region(self, x0, x1, y0, y1):
        self.eng.fill_region(x0, x1, y0, y1)

    def fill_density(self):
        assert self.lib.vpe_fill_density(self.eng._ctx) == 0

    def fill_sweep_region(self, x0, x1, y0, y1):
        self.eng.fill_sweep_region(x0, x1, y0, y1)

    def sheet_tensor(self):
        return self._sheet

    def sheet_written(self, y0, y1):  # rows received from the previous slab -> oracle context
        full = self.eng.read_light_sheet()
        full[y0 * self.N:y1 * self.N] = self._sheet[y0 * self.N:y1 * self.N].numpy()
        assert self.lib.vpe_ref_write_light_sheet(self.eng._ctx, full.ctypes.data) == 0

    def sheet_read(self, y0, y1):     # oracle context -> rows to send to the next slab
        full = self.eng.read_light_sheet()
        self._sheet[y0 * self.N:y1 * self.N] = torch.from_numpy(full[y0 * self.N:y1 * self.N])

    def buffer(self, name, shape):
        return torch.zeros(tuple(shape), dtype=torch.float32)

    def copy_band_to_host(self, band, host_rows):
        host_rows[...] = band[:host_rows.shape[0]].numpy()

    def march_partial(self, camera, padded_rows=None):
        c = _camera(camera)
        rows = max(c.height, padded_rows or c.height)
        over = np.zeros((rows, c.width, 4), dtype=np.float32)
        under = np.zeros((rows, c.width, 4), dtype=np.float32)
        rc = self.lib.vpe_ref_march_partial(self.eng._ctx, C.byref(c), over.ctypes.data, under.ctypes.data, None)
        assert rc == 0
        return torch.from_numpy(over), torch.from_numpy(under)

    def last_ray_samples(self):
        return int(self.eng.stats()["raySamples"])

    def composite(self, parts, num_pixels):
        """Reference-order compositing of slab partials (≙ k_composite): OVER parts in ascending slab
        order with Blend One OneMinusSrcAlpha, then UNDER parts with Blend OneMinusDstAlpha One."""
        dst = np.zeros((num_pixels, 4), dtype=np.float32)
        one = np.float32(1.0)
        for s in range(len(parts) // 2):
            src = parts[2 * s].reshape(-1, 4).numpy()
            dst = src + dst * (one - src[:, 3:4])
        for s in range(len(parts) // 2):
            src = parts[2 * s + 1].reshape(-1, 4).numpy()
            dst = src * (one - dst[:, 3:4]) + dst
        return torch.from_numpy(dst.astype(np.float32))

    def all_reduce_sum(self, dist, value):
        t = torch.tensor([value], dtype=torch.int64)
        dist.all_reduce(t)
        return int(t.item())

    # -- load balancing (SlabRenderer.rebalance): the oracle's slab is fixed at create, so a moved slab is a new context;
    # "times" are deterministic stand-ins (particle pairs of the slab for the density pass, ray samples for the march)
    def set_slab(self, z0, z1):
        from vpe_b200 import scenes
        self.slab = (int(z0), int(z1))
        self.eng = oracle_engine(self.scene, slab=self.slab)
        scenes.apply_scene(self.eng, self.scene)

    def record_event(self):
        return None

    def elapsed_ms(self, a, b):
        return float(self.eng.stats()["numParticlePairs"]) * 1e-2

    def stats(self):
        st = self.eng.stats()
        st["marchKernelMs"] = float(st["raySamples"]) * 1e-5
        return st
