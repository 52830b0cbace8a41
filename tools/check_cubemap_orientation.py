#!/usr/bin/env python
"""Evidence for the orientation of the displacement cubemap (decision C-1).

The engine takes the faces and rows of Assets/Textures/DisplacementTexture.cubemap exactly as stored
(tools/decode_displacement_cubemap.py) and addresses them with the D3D11 rule: face order +X,-X,+Y,-Y,+Z,-Z,
row r <-> v = (r + 0.5) / E with v = 0 at the top of the face (+Y side for the four side faces), column c <-> u.
The reference cannot be run, so this cannot be checked against a rendered frame. What can be checked is how the stored
blob relates to the loose source images cm_c00..05.png (PNG row 0 = top of the picture): for every stored face this
script finds the source image and the dihedral transform (8 candidates) with the smallest mean absolute difference.

Result recorded in DESIGN.md: every stored face is ONE source image mirrored left-right, never flipped vertically,
never rotated (mean abs diff ~2/255 from re-encoding against 40-60/255 for every other candidate) - i.e. stored row 0
is the TOP row of the artist's picture, and the columns run right-to-left as seen in the picture, which is exactly the
layout Unity documents for cubemap faces (Cubemap.GetPixels: "laid out right to left, top to bottom") and the layout
D3D11 samples without any flip (a face seen from inside the cube is the mirror image of the face seen from outside).

Second check, seam continuity under the D3D addressing rule: the six source pictures are a SEAMLESS cube map only when
read as c00..c05 = +X,-X,+Y,-Y,+Z,-Z un-mirrored (3.4/255 across the cube edges, less than between neighbouring texels
inside a face). The stored asset assigns them to other slots (c02,c03,c04,c05,c00,c01) and is not seamless under ANY reading
of its rows and columns (44-47/255): whoever built the asset did not preserve the seams, so continuity cannot tell "as
stored" from "mirrored back" - both give a displacement field with the same texel statistics and seams along the cube
edges. Needs /root/reference and PIL for the first part; nothing at test / bench time depends on it."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/Assets/Textures"

TRANSFORMS = {
    "identity": lambda a: a, "mirror left-right": lambda a: a[:, ::-1], "flip top-bottom": lambda a: a[::-1],
    "rotate 180": lambda a: a[::-1, ::-1], "transpose": lambda a: a.T, "rotate 90": lambda a: a.T[:, ::-1],
    "rotate 270": lambda a: a.T[::-1], "anti-transpose": lambda a: a.T[::-1, ::-1]}


def analyse():
    from PIL import Image
    import vpe_b200
    from vpe_b200 import scenes
    stored = scenes.load_displacement_cubemap().astype(np.float64)
    src = [np.asarray(Image.open(os.path.join(REF, "cm_c%02d.png" % k)).convert("RGB"), dtype=np.float64)[..., 0] for k in range(6)]
    rows = []
    for f in range(6):
        cands = sorted((float(np.abs(fn(src[k]) - stored[f]).mean()), k, name) for k in range(6) for name, fn in TRANSFORMS.items())
        rows.append((f, cands[0], cands[1]))
    return rows


def d3d_lookup(cube, d):
    """Nearest texel of the D3D11 cube addressing rule (the rule of oracle sample_cube / k_fill_columns)."""
    E = cube.shape[1]
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    ax, ay, az = np.abs(x), np.abs(y), np.abs(z)
    isx = (ax >= ay) & (ax >= az)
    isy = ~isx & (ay >= az)
    face = np.where(isx, np.where(x >= 0, 0, 1), np.where(isy, np.where(y >= 0, 2, 3), np.where(z >= 0, 4, 5)))
    ma = np.where(isx, ax, np.where(isy, ay, az))
    sc = np.where(isx, np.where(x >= 0, -z, z), np.where(isy, x, np.where(z >= 0, x, -x)))
    tc = np.where(isx, -y, np.where(isy, np.where(y >= 0, z, -z), -y))
    c = np.clip(((sc / ma + 1) * 0.5 * E).astype(int), 0, E - 1)
    r = np.clip(((tc / ma + 1) * 0.5 * E).astype(int), 0, E - 1)
    return cube[face, r, c]


def seam_discontinuity(cube):
    """Mean |difference| between the texels on either side of each of the 12 cube edges (mean over edges, worst edge)."""
    E = cube.shape[1]
    eps, ts = 1.0 / E, (np.arange(E) + 0.5) / E * 2 - 1
    per_edge = []
    for a in range(3):
        for b in range(a + 1, 3):
            for sa in (1, -1):
                for sb in (1, -1):
                    d1, d2 = np.zeros((E, 3)), np.zeros((E, 3))
                    d1[:, a], d1[:, b], d1[:, 3 - a - b] = sa, sb * (1 - eps), ts
                    d2[:, a], d2[:, b], d2[:, 3 - a - b] = sa * (1 - eps), sb, ts
                    per_edge.append(float(np.abs(d3d_lookup(cube, d1) - d3d_lookup(cube, d2)).mean()))
    return float(np.mean(per_edge)), float(np.max(per_edge))


def seam_report():
    import vpe_b200
    from vpe_b200 import scenes
    cube = scenes.load_displacement_cubemap().astype(np.float64)
    inside = (np.abs(np.diff(cube, axis=1)).mean() + np.abs(np.diff(cube, axis=2)).mean()) / 2
    readings = {
        "as stored (the engine's reading: direct upload)": cube,
        "every face mirrored back (= the source pictures in the stored slots)": cube[:, :, ::-1],
        "rows flipped": cube[:, ::-1, :],
        "source pictures c00..c05 as +X,-X,+Y,-Y,+Z,-Z, un-mirrored": cube[[4, 5, 0, 1, 2, 3]][:, :, ::-1]}
    return float(inside), {k: seam_discontinuity(v) for k, v in readings.items()}


if __name__ == "__main__":
    names = ["+X", "-X", "+Y", "-Y", "+Z", "-Z"]
    for f, best, second in analyse():
        print("stored face %d (%s): cm_c%02d.png, %-17s mean |diff| %5.2f / 255   (next best: cm_c%02d %s %.1f)" % (
            f, names[f], best[1], best[2] + ",", best[0], second[1], second[2], second[0]))
    inside, seams = seam_report()
    print("neighbouring texels inside a face: mean |diff| %.2f / 255" % inside)
    for k, (m, worst) in seams.items():
        print("%-70s across the 12 cube edges: mean |diff| %5.2f, worst edge %5.2f" % (k, m, worst))
