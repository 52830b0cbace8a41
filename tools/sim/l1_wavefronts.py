"""Offline model of the march's L1 gather cost (no GPU): wavefronts per warp-wide LDG.64 = the largest number of
distinct 8-byte words that fall into one of the 32 four-byte banks (measured rule, tools/microbench/l1_gather.cu),
for the cfg3 camera / light geometry, an 8x4 pixel warp tile and several brick layouts.
usage: python tools/sim/l1_wavefronts.py"""
import numpy as np

W, H, FOV, G, N, B = 1920, 1080, 60.0, 32, 32, 1
q = np.array([.185594, 0, 0, .982627]); q /= np.linalg.norm(q)
x, y, z, w = q
R = np.array([[1-2*(y*y+z*z), 2*(x*y-z*w), 2*(x*z+y*w)], [2*(x*y+z*w), 1-2*(x*x+z*z), 2*(y*z-x*w)], [2*(x*z-y*w), 2*(y*z+x*w), 1-2*(x*x+y*y)]])
cam = np.array([0, 0, -0.75 * G])
step = 1.73205 / 64.0  # metavoxel units per sample

def rays(px, py):
    dx = (2 * (px + .5) / W - 1) * W / H; dy = 2 * (py + .5) / H - 1; dz = 1 / np.tan(np.radians(FOV / 2))
    d = np.stack([dx, dy, np.full_like(dx, dz)], -1); d /= np.linalg.norm(d, axis=-1, keepdims=True)
    return d  # camera looks down +z in world (Unity), image y up or down does not matter here

def texels(px, py, t):
    """base texel + brick id of the sample at distance t (metavoxel units) along each ray"""
    d = rays(px, py)
    p = cam + d * t[..., None]
    ls = p @ R  # world -> light space (R^T p)
    gcoord = ls + G / 2  # metavoxel (i) spans [i-.5, i+.5]
    mv = np.floor(gcoord + .5).astype(int)
    loc = gcoord - mv  # [-.5,.5]
    f = (loc + .5) * (N - 2 * B) + B - .5
    base = np.floor(f).astype(int)
    inside = ((mv >= 0) & (mv < G)).all(-1)
    return mv, base, inside

def wavefronts(addr_bytes):
    """addr_bytes: (32,) int64 addresses of 8-byte loads, -1 = inactive lane"""
    a = np.unique(addr_bytes[addr_bytes >= 0])
    if len(a) == 0: return 0
    banks = np.zeros(32, int)
    for v in a:
        b = (v // 4) % 32
        banks[b] += 1; banks[(b + 1) % 32] += 1
    return banks.max()

def layout_addr(mv, base, RS, SP, swz=None):
    brick = (mv[..., 2] * G + mv[..., 1]) * G + mv[..., 0]
    SS = N * RS + SP
    x0, y0, z0 = base[..., 0], base[..., 1], base[..., 2]
    out = []
    for dy_ in (0, 1):
        for dx_ in (0, 1):
            xx = x0 + dx_
            if swz: xx = swz(xx, y0 + dy_, z0)
            out.append(((brick.astype(np.int64) * N * SS) + z0 * SS + (y0 + dy_) * RS + xx) * 8)
    return out

rng = np.random.default_rng(0)
layouts = {"row 32 (no pad)": (32, 0), "row 40 (r01)": (40, 0), "row 40 + slice pad 2": (40, 2), "row 40 + slice pad 4": (40, 4), "row 40 + slice pad 6": (40, 6),
           "row 36": (36, 0), "row 36 + slice pad 2": (36, 2), "row 34 + slice pad 4": (34, 4), "row 33": (33, 0), "row 33 + sp 3": (33, 3), "row 34": (34, 0), "row 34 + sp 6": (34, 6)}
for jitter in (0, 4, 12):
    res = {k: [] for k in layouts}
    for trial in range(3000):
        tx = rng.integers(0, W // 8) * 8; ty = rng.integers(0, H // 4) * 4
        px, py = np.meshgrid(np.arange(tx, tx + 8), np.arange(ty, ty + 4)); px = px.ravel().astype(float); py = py.ravel().astype(float)
        k0 = rng.uniform(8, 42) / step
        k = np.round(k0 + (rng.integers(-jitter, jitter + 1, 32) if jitter else 0))
        mv, base, inside = texels(px, py, k * step)
        if inside.sum() < 16: continue
        for name, (RS, SP) in layouts.items():
            tot = 0
            for a in layout_addr(mv, base, RS, SP):
                a = np.where(inside, a, -1)
                tot += wavefronts(a)
            res[name].append(tot / 4)
    print("k jitter +-%d samples between the lanes of a warp:" % jitter)
    for name in layouts:
        print("   %-24s %.2f wavefronts per LDG.64" % (name, np.mean(res[name])))
