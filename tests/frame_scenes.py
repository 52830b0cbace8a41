"""Shared inputs of the tests for what surrounds the path in the reference's frame: occluder
meshes for the light depth map, a scene depth buffer, a scene colour buffer."""
import numpy as np

from vpe_b200 import scenes


def quad(center, half_u, half_v, clockwise_from_light=False):
    """Two triangles of a parallelogram centre +- half_u +- half_v (world space). The winding is
    counter-clockwise for a viewer on the side the cross product u x v points AWAY from ... the tests
    simply render both windings and check that exactly one of them is drawn."""
    c, u, v = [np.asarray(a, dtype=np.float64) for a in (center, half_u, half_v)]
    p00, p10, p11, p01 = c - u - v, c + u - v, c + u + v, c - u + v
    tris = np.array([[p00, p10, p11], [p00, p11, p01]])
    if clockwise_from_light:
        tris = tris[:, ::-1, :]
    return tris.astype(np.float32)


def light_frame(sc):
    """World-space right / up / forward unit vectors of the directional light."""
    q = sc["light"]["rotation"]
    r = scenes.quat_rotate(q, np.eye(3))
    return r[0], r[1], r[2]


def occluders(sc):
    """A tilted plate in front of the left third of the grid and a small plate deeper in, both given
    with the winding the light camera draws (Cull Front keeps the faces that look away from the light)."""
    right, up, fwd = light_frame(sc)
    g = sc["grid"][0] * sc["mvScale"]
    c1 = -0.25 * g * right - 0.7 * g * fwd
    plate1 = quad(c1, 0.2 * g * right + 0.05 * g * fwd, 0.45 * g * up)
    c2 = 0.2 * g * right + 0.1 * g * up - 0.1 * g * fwd
    plate2 = quad(c2, 0.12 * g * right, 0.1 * g * up + 0.03 * g * fwd)
    return np.concatenate([plate1, plate2])


def scene_depth(sc, near_fraction=0.5):
    """An opaque wall across the right part of the image at the grid centre's depth, nothing elsewhere."""
    cam = sc["camera"]
    h, w = cam["height"], cam["width"]
    d = np.full((h, w), 3.0e38, dtype=np.float32)
    dist = float(np.linalg.norm(np.asarray(cam["position"], dtype=np.float64)))
    d[:, int(w * near_fraction):] = dist
    d[: h // 4, : w // 4] = 0.05  # something right in front of the camera
    return d
