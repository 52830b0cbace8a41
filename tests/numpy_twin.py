"""An independent numpy (float64) restatement of the hot path, written from SURVEY.md Appendix A —
NOT from oracle/vpe_ref.cpp.  TEST INFRASTRUCTURE.

It exists to validate the C++ oracle by independent means (the reference ships no golden vectors,
SURVEY §4): a second implementation with different structure (vectorised over voxels, closed-form
positions instead of accumulated ones, double precision, numpy's own matrix inverse) must agree
with the oracle up to fp32/fp16 rounding and up to decisions that sit on a discontinuity.

Citations: VPR.cs = Assets/Main Scene/VolumetricParticleRenderer.cs, Fill.shader / March.shader =
Assets/Shaders/Metavoxel/{FillVolume,RayMarchVoxel}.shader, MathUtil.cs (all under /root/reference).
"""
import math

import numpy as np


# ---- UnityEngine math ----------------------------------------------------------------------------
def quat_matrix(q):
    x, y, z, w = [float(t) for t in q]
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def trs(t, q, s):
    m = np.eye(4)
    m[:3, :3] = quat_matrix(q) * float(s)
    m[:3, 3] = np.asarray(t, dtype=np.float64)
    return m


def angle_axis(deg, axis):
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    h = math.radians(float(deg)) / 2
    return (a[0] * math.sin(h), a[1] * math.sin(h), a[2] * math.sin(h), math.cos(h))


def xf(m, p):
    p = np.asarray(p, dtype=np.float64)
    return p @ m[:3, :3].T + m[:3, 3]


def round_half_even(v):
    return int(np.rint(v))  # Mathf.RoundToInt


class Twin:
    def __init__(self, sc, cubemap_r8):
        self.sc = sc
        self.G = np.array(sc["grid"], dtype=np.int64)
        self.N = int(sc["numVoxels"])
        self.b = int(sc["border"])
        self.s = float(np.float32(sc["mvScale"]))
        self.sb = self.s * self.N / (self.N - 2 * self.b)                    # VPR.cs:139
        self.lq = sc["light"]["rotation"]
        self.L2W = trs(sc["light"]["position"], self.lq, 1.0)
        self.W2L = np.linalg.inv(self.L2W)
        self.center = np.asarray(sc["gridCenter"], dtype=np.float64)
        self.f = quat_matrix(self.lq)[:, 2]                                  # dirLight.transform.forward
        self.cube = cubemap_r8.astype(np.float64) / 255.0
        self.E = cubemap_r8.shape[1]
        self.lists = None
        self.depth_map = None                                                # lightDepthMap (VPR.cs:274), None = cleared to 1

    # VPR.cs:320-367 InitCameraAtLight / UpdatePositionOfCameraAtLight: orthographic, at centre - forward * 200
    def light_camera_to_world(self):
        return trs(self.center - self.f * 200.0, self.lq, 1.0)

    # VPR.cs:184 + GenerateLightDepthMap.shader:6-31 (Cull Front, ZWrite On, ZTest Less, depth only)
    def rasterize_depth(self, tris, near=0.3, far=1000.0, edge_eps=1e-4):
        """Depth map [NY*N][NX*N] of the occluders seen from the light, and a mask of pixels whose centre lies within
        edge_eps (barycentric) of a drawn triangle's edge: there the fill rule decides, which this restatement does not
        model (the oracle uses D3D's top-left rule).  Barycentrics come from a 2x2 solve per triangle, not from edge
        functions; double precision."""
        W, H = int(self.G[0]) * self.N, int(self.G[1]) * self.N
        r, t = self.G[0] * self.s * 0.5, self.G[1] * self.s * 0.5                 # VPR.cs:340 orthographic extents
        w2lc = np.linalg.inv(self.light_camera_to_world())
        depth = np.ones((H, W))
        edge = np.zeros((H, W), dtype=bool)
        for tri in np.asarray(tris, dtype=np.float64):
            p = xf(w2lc, tri)
            sx = (p[:, 0] / r * 0.5 + 0.5) * W
            sy = (p[:, 1] / t * 0.5 + 0.5) * H
            sz = (p[:, 2] - near) / (far - near)
            m = np.array([[sx[1] - sx[0], sx[2] - sx[0]], [sy[1] - sy[0], sy[2] - sy[0]]])
            det = np.linalg.det(m)
            if not det > 0:                                                      # clockwise from the light = front face: culled
                continue
            x0, x1 = max(0, int(np.floor(sx.min() - 0.5))), min(W - 1, int(np.ceil(sx.max() - 0.5)))
            y0, y1 = max(0, int(np.floor(sy.min() - 0.5))), min(H - 1, int(np.ceil(sy.max() - 0.5)))
            if x1 < x0 or y1 < y0:
                continue
            gx, gy = np.meshgrid(np.arange(x0, x1 + 1) + 0.5, np.arange(y0, y1 + 1) + 0.5)
            bc = np.linalg.solve(m, np.stack([gx.ravel() - sx[0], gy.ravel() - sy[0]]))
            l1, l2 = bc[0].reshape(gx.shape), bc[1].reshape(gx.shape)
            l0 = 1 - l1 - l2
            lmin = np.minimum(l0, np.minimum(l1, l2))
            z = l0 * sz[0] + l1 * sz[1] + l2 * sz[2]
            inside = (lmin >= 0) & (z >= 0) & (z <= 1)
            sub = depth[y0:y1 + 1, x0:x1 + 1]
            sub[...] = np.where(inside & (z < sub), z, sub)
            edge[y0:y1 + 1, x0:x1 + 1] |= np.abs(lmin) < edge_eps
        return depth, edge

    # tex2D(_LightDepthMap, uv): bilinear, clamp (Fill.shader:216)
    def sample_depth(self, u, v):
        if self.depth_map is None:
            return np.ones_like(u)
        H, W = self.depth_map.shape
        fx, fy = u * W - 0.5, v * H - 0.5
        x0, y0 = np.floor(fx), np.floor(fy)
        wx, wy = fx - x0, fy - y0
        x0, y0 = x0.astype(np.int64), y0.astype(np.int64)
        cx, cy = (lambda a: np.clip(a, 0, W - 1)), (lambda a: np.clip(a, 0, H - 1))
        d = self.depth_map
        top = d[cy(y0), cx(x0)] * (1 - wx) + d[cy(y0), cx(x0 + 1)] * wx
        bot = d[cy(y0 + 1), cx(x0)] * (1 - wx) + d[cy(y0 + 1), cx(x0 + 1)] * wx
        return top * (1 - wy) + bot * wy

    # A.1  VPR.cs:370-394
    def mv_center(self, x, y, z):
        half = self.G // 2                                                   # integer halves
        off = np.array([(half[0] - x) * self.s, (half[1] - y) * self.s, (half[2] - z) * self.s])
        return xf(self.L2W, xf(self.W2L, self.center) - off)

    # A.2 / A.3  VPR.cs:397-457, 582-586; MathUtil.cs:11-25
    def bin_particles(self, particles, emitter, margin=None):
        """Returns {(x,y,z): [particle indices]} and, when `margin` is a list, appends the |r2| of every
        sphere/box decision so the caller can tell borderline pairs."""
        E2W = trs(emitter["position"], emitter["rotation"], 1.0)
        efwd = quat_matrix(emitter["rotation"])[:, 2]
        G = self.G
        lists = {}
        self.W2P, self.opacity, self.ws = [], [], []
        lsC = xf(self.W2L, self.center)
        for pi, p in enumerate(np.asarray(particles, dtype=np.float64)):
            ws = xf(E2W, p[0:3])
            size = p[3]
            self.ws.append(ws)
            self.W2P.append(np.linalg.inv(trs(ws, angle_axis(p[4], efwd), size)))
            self.opacity.append(p[5] / p[6])
            idx = (xf(self.W2L, ws) - lsC) / self.s + G * 0.5
            r = size / 2
            e = round_half_even(np.float32(r) / np.float32(self.s))
            lo = np.maximum(0.0, idx - e)
            hi = np.minimum(G - 1.0, idx + e)
            for z in range(int(lo[2]), int(hi[2]) + 1):
                for y in range(int(lo[1]), int(hi[1]) + 1):
                    for x in range(int(lo[0]), int(hi[0]) + 1):
                        M = trs(self.mv_center(x, y, z), self.lq, self.sb)
                        q = xf(np.linalg.inv(M), ws)
                        rr = r / self.sb
                        r2 = rr * rr
                        for a in range(3):
                            if q[a] < -0.5:
                                r2 -= (q[a] + 0.5) ** 2
                            elif q[a] > 0.5:
                                r2 -= (q[a] - 0.5) ** 2
                        if margin is not None:
                            margin.append(((x, y, z), pi, r2))
                        if r2 > 0:
                            lists.setdefault((x, y, z), []).append(pi)
        self.lists = lists
        return lists

    # texCUBE(...).x: D3D face selection, bilinear, clamp (Fill.shader:116)
    def sample_cube(self, d):
        d = np.asarray(d, dtype=np.float64)
        ax, ay, az = np.abs(d[..., 0]), np.abs(d[..., 1]), np.abs(d[..., 2])
        x, y, z = d[..., 0], d[..., 1], d[..., 2]
        isx = (ax >= ay) & (ax >= az)
        isy = ~isx & (ay >= az)
        isz = ~isx & ~isy
        face = np.where(isx, np.where(x >= 0, 0, 1), np.where(isy, np.where(y >= 0, 2, 3), np.where(z >= 0, 4, 5)))
        ma = np.where(isx, ax, np.where(isy, ay, az))
        sc = np.where(isx, np.where(x >= 0, -z, z), np.where(isy, x, np.where(z >= 0, x, -x)))
        tc = np.where(isx, -y, np.where(isy, np.where(y >= 0, z, -z), -y))
        ma = np.where(ma == 0, 1.0, ma)
        u = (sc / ma + 1) * 0.5
        v = (tc / ma + 1) * 0.5
        E = self.E
        fx, fy = u * E - 0.5, v * E - 0.5
        x0, y0 = np.floor(fx), np.floor(fy)
        wx, wy = fx - x0, fy - y0
        x0 = x0.astype(np.int64)
        y0 = y0.astype(np.int64)
        cl = lambda a: np.clip(a, 0, E - 1)
        t00 = self.cube[face, cl(y0), cl(x0)]
        t10 = self.cube[face, cl(y0), cl(x0 + 1)]
        t01 = self.cube[face, cl(y0 + 1), cl(x0)]
        t11 = self.cube[face, cl(y0 + 1), cl(x0 + 1)]
        top = t00 + wx * (t10 - t00)
        bot = t01 + wx * (t11 - t01)
        return top + wy * (bot - top)

    # A.4-A.6  Fill.shader:96-135, 152-269 for one metavoxel; sheet_in/out are (N,N) arrays [y][x]
    def fill_metavoxel(self, x, y, z, sheet_in):
        sc, N, b = self.sc, self.N, self.b
        plist = self.lists[(x, y, z)]
        N2W = trs(self.mv_center(x, y, z), self.lq, self.sb)
        u = (np.arange(N) + 0.5 - N / 2) / N
        k = np.arange(N)
        # voxel (k, yy, xx): x,y centred, z at the near face of the slice
        base = np.stack(np.meshgrid(u, u, indexing="xy"), axis=-1)            # [yy][xx] -> (u_x, u_y)
        p0 = np.concatenate([base, np.full((N, N, 1), (0 - N / 2) / N)], axis=-1)
        vox0 = xf(N2W, p0)                                                    # [yy][xx][3]
        vox = vox0[None] + (k[:, None, None, None] * (self.sb / N)) * self.f   # [k][yy][xx][3]
        density = np.zeros((N, N, N))
        ao = np.zeros((N, N, N))
        ds, of = float(np.float32(sc["displacementScale"])), float(np.float32(sc["opacityFactor"]))
        near_surface = np.zeros((N, N, N), dtype=bool)
        count = np.zeros((N, N, N), dtype=np.int64)                          # particles that contain the voxel
        for n_i, pi in enumerate(plist):
            q = xf(self.W2P[pi], vox)
            qq = (q * q).sum(-1)
            inside = qq <= 0.25
            near_surface |= np.abs(qq - 0.25) < 1e-5
            # texCUBE filters each face on its own (clamp): a direction whose two largest components tie sits on a face seam,
            # where fp32 and fp64 may pick different faces - a discontinuity like the sphere surface
            srt = np.sort(np.abs(q), axis=-1)
            near_surface |= inside & (srt[..., 1] > srt[..., 2] * (1 - 1e-4))
            count += inside
            raw = self.sample_cube(2 * q)
            nd = ds * raw + (1 - ds)
            d2 = 4 * qq
            t = np.clip((d2 - nd) / (0.7 * nd - nd), 0, 1)
            dens = t * t * (3 - 2 * t) * of
            if int(sc["fadeOutParticles"]) == 1:
                dens = dens * self.opacity[pi]
            density += np.where(inside, dens, 0.0)
            ao = np.where(inside, nd if n_i == 0 else np.maximum(ao, nd), ao)
        # occlusion, Fill.shader:211-221: the scene's depth seen from the light at this voxel column (cleared map = far plane)
        z0 = xf(np.linalg.inv(self.light_camera_to_world()), vox0)[..., 2]
        px = np.arange(N) + 0.5
        uu = np.broadcast_to((px[None, :] + x * N) / (self.G[0] * N), (N, N))
        vv = np.broadcast_to((px[:, None] + y * N) / (self.G[1] * N), (N, N))
        depth = self.sample_depth(uu, vv) * (1000.0 - 0.3) + 0.3
        fsi = (depth - z0) / (self.sb / N)
        shadow = np.trunc(fsi)
        self.last_shadow_margin = np.abs(fsi - np.rint(fsi))                  # columns whose shadow index sits on an integer
        self.last_shadow = shadow                                             # first slice in shadow, per voxel column
        T = np.ones((N, N)) if z == 0 else sheet_in.astype(np.float64).copy()
        prop = T.copy()
        amb = np.asarray(sc["ambient"], dtype=np.float32).astype(np.float64)
        out = np.zeros((N, N, N, 4))
        for kk in range(N):
            sh = kk >= shadow
            T = np.where(sh, 0.0, T)
            if kk < N - b:
                prop = np.where(sh, prop, T)
            out[kk, ..., 0:3] = 0.4 * T[..., None] + amb * ao[kk][..., None]
            out[kk, ..., 3] = density[kk]
            T = T * (1 / (1 + density[kk]))
        self.last_inside_count = count
        return out, prop, near_surface

    # A.7  VPR.cs:613-711: metavoxels in submission order with their blend mode
    def draw_order(self, cam_pos):
        G = self.G
        cam_pos = np.asarray(cam_pos, dtype=np.float64)
        keys = []
        for y in range(G[1]):
            for x in range(G[0]):
                d = self.mv_center(x, y, 0) - cam_pos
                keys.append((float(np.float32(d @ d)), y, x))
        asc = sorted(keys)
        ls_cam = xf(self.W2L, cam_pos)[2]
        ls0 = xf(self.W2L, self.mv_center(0, 0, 0))[2]
        zb = min(max(round_half_even((ls_cam - ls0) / self.s), -1), int(G[2]) - 1)
        order = []
        for z in range(0, zb + 1):
            order += [((x, y, z), True) for (_, y, x) in reversed(asc)]
        for z in range(zb + 1, int(G[2])):
            order += [((x, y, z), False) for (_, y, x) in asc]
        return [(mv, over) for (mv, over) in order if mv in self.lists], zb

    # A.8  March.shader:187-279,301 + ROP blend for one pixel; bricks: {(x,y,z): float array [k][y][x][4]}
    def march_pixel(self, cam, px, py, bricks, order):
        sc, N, b, s = self.sc, self.N, self.b, self.s
        W, H = cam["width"], cam["height"]
        camL2W = trs(cam["position"], cam["rotation"], 1.0)
        C2W = camL2W.copy()
        C2W[:3, 2] *= -1                                                     # OpenGL-style view space
        W2C = np.linalg.inv(C2W)
        d = np.array([(2 * (px + 0.5) / W - 1) * (W / H), 2 * (py + 0.5) / H - 1,
                      -1 / math.tan(math.radians(cam["fovYDegrees"]) / 2)])
        d /= np.linalg.norm(d)
        maxG = float(self.G.max())
        half = 1.73205 * 0.5 * maxG * s
        zmin = xf(W2C, self.center)[2] + half
        o_c = d * (zmin / d[2])
        step = (2 * half / s) / (maxG * sc["rayMarchSteps"])
        soft = int(sc["softDistance"])
        dst = np.zeros(4)
        total = 0
        for (mv, over) in order:
            C2M = np.linalg.inv(trs(self.mv_center(*mv), self.lq, s)) @ C2W
            o = xf(C2M, o_c)
            dd = d @ C2M[:3, :3].T
            dd /= np.linalg.norm(dd)
            with np.errstate(divide="ignore", invalid="ignore"):
                ta, tb = (-0.5 - o) / dd, (0.5 - o) / dd
            t1, t2 = np.minimum(ta, tb).max(), np.maximum(ta, tb).min()
            if t1 > t2:
                continue
            k_in, k_out = math.ceil(t1 / step), math.floor(t2 / step)
            k_cam = int(np.linalg.norm(xf(C2M, np.zeros(3)) - o) / step)
            k_in = max(k_in, k_cam)
            rgb = np.zeros(3)
            T = 1.0
            brick = bricks[mv]
            for k in range(k_out, k_in - 1, -1):
                pos = o + k * step * dd
                uvw = (pos + 0.5) * (1 - 2 * b / N) + b / N
                f = uvw * N - 0.5
                i0 = np.floor(f).astype(int)
                w = f - i0
                c = np.zeros(4)
                for dz in (0, 1):
                    for dy in (0, 1):
                        for dx in (0, 1):
                            wt = (w[0] if dx else 1 - w[0]) * (w[1] if dy else 1 - w[1]) * (w[2] if dz else 1 - w[2])
                            c += wt * brick[(i0[2] + dz) % N, (i0[1] + dy) % N, (i0[0] + dx) % N]
                dens = c[3]
                if k - k_cam < soft:
                    dens *= (k - k_cam) / soft
                bf = 1 / (1 + dens)
                rgb = c[:3] + bf * (rgb - c[:3])
                T *= bf
                total += 1
            src = np.array([rgb[0], rgb[1], rgb[2], 1 - T])
            if over:
                dst = src + (1 - src[3]) * dst
            else:
                dst = (1 - dst[3]) * src + dst
        return dst, total
