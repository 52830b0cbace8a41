"""Offline geometry of a CTA-cooperative march (no GPU): what would have to be staged in shared memory if the work item
were (pixel tile, metavoxel) and the tile's rays sampled a brick sub-box fetched by TMA (an axis-aligned box of texels,
the only shape cp.async.bulk.tensor moves). cfg3 camera / light geometry as in l1_wavefronts.py.
For random (tile, metavoxel) pairs that the tile's rays cross: samples served, the axis-aligned bounding box of all
trilinear footprints, bytes staged per sample at 8 B/texel (z-paired grey texels) and 4 B/texel (un-paired r,density).
usage: python tools/sim/tile_boxes.py"""
import numpy as np

W, H, FOV, G, N, B = 1920, 1080, 60.0, 32, 32, 1
q = np.array([.185594, 0, 0, .982627]); q /= np.linalg.norm(q)
x, y, z, w = q
R = np.array([[1-2*(y*y+z*z), 2*(x*y-z*w), 2*(x*z+y*w)], [2*(x*y+z*w), 1-2*(x*x+z*z), 2*(y*z-x*w)], [2*(x*z-y*w), 2*(y*z+x*w), 1-2*(x*x+y*y)]])
cam = np.array([0, 0, -0.75 * G])
step = 1.73205 / 64.0  # metavoxel units per sample (March.shader:222-224 with 64 steps per metavoxel)


def ray_dirs(px, py):
    dx = (2 * (px + .5) / W - 1) * W / H; dy = 2 * (py + .5) / H - 1; dz = 1 / np.tan(np.radians(FOV / 2))
    d = np.stack([dx, dy, np.full_like(dx, dz)], -1)
    return d / np.linalg.norm(d, axis=-1, keepdims=True)


def tile_fragment(rng, tw, th):
    """One (tile, metavoxel) work item: returns (samples, box dims in texels) or None."""
    tx = rng.integers(0, W // tw) * tw; ty = rng.integers(0, H // th) * th
    px, py = np.meshgrid(np.arange(tx, tx + tw), np.arange(ty, ty + th))
    d = ray_dirs(px.ravel().astype(float), py.ravel().astype(float))           # (rays, 3)
    k = np.arange(int(8 / step), int(42 / step))                                # the depth range that holds the grid
    p = cam + d[:, None, :] * (k[None, :, None] * step)                         # (rays, steps, 3) world
    g = p @ R + G / 2                                                           # light space, metavoxel i spans [i-.5, i+.5]
    mv = np.floor(g + .5).astype(int)
    inside = ((mv >= 0) & (mv < G)).all(-1)
    if not inside.any():
        return None
    # pick the metavoxel under a random inside sample of the tile's centre ray
    c = d.shape[0] // 2
    ks = np.nonzero(inside[c])[0]
    if len(ks) == 0:
        return None
    target = mv[c, rng.choice(ks)]
    sel = inside & (mv == target).all(-1)
    if sel.sum() == 0:
        return None
    f = ((g - mv) + .5) * (N - 2 * B) + B - .5                                   # texel coordinate inside the brick
    base = np.floor(f).astype(int)[sel]
    lo, hi = base.min(0), base.max(0) + 1                                        # footprints reach base + 1
    return int(sel.sum()), tuple((hi - lo + 1).tolist())


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    print("tile      samples/item   box (x,y,z texels)      texels   KB @8B  KB @4B   staged B/sample @8B  @4B   (today: 32 B/sample from L1, 81 % hits)")
    for (tw, th) in [(8, 4), (16, 8), (32, 16), (64, 32)]:
        rows = [r for r in (tile_fragment(rng, tw, th) for _ in range(400)) if r]
        s = np.array([r[0] for r in rows], dtype=float); b = np.array([r[1] for r in rows], dtype=float)
        vol = b.prod(1)
        print("%3dx%-3d   %9.0f      %5.1f x %5.1f x %5.1f   %8.0f   %6.1f  %6.1f   %10.1f  %10.1f" % (
            tw, th, s.mean(), b[:, 0].mean(), b[:, 1].mean(), b[:, 2].mean(), vol.mean(), vol.mean() * 8 / 1024, vol.mean() * 4 / 1024,
            (vol * 8).sum() / s.sum(), (vol * 4).sum() / s.sum()))
