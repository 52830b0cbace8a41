#!/usr/bin/env python
"""Decode the reference's displacement cubemap into the R8 fixture the engine takes as input.

Source : /root/reference/Assets/Textures/DisplacementTexture.cubemap (Unity YAML; lines 10-27:
         128x128, 6 images, TextureFormat 5 = ARGB32, no mips, bilinear, clamp).
Rule   : `_typelessdata` hex -> bytes -> [6][128][128][4] in byte order A,R,G,B; keep R, the only
         channel FillVolume.shader:116 reads (`texCUBE(...).x`). Faces stay in stored (Unity) order
         +X,-X,+Y,-Y,+Z,-Z, rows in stored order.
Output : volumetric-particles-for-unity_b200/assets/displacement_r8.bin (6*128*128 bytes).
The reference tree is only needed to regenerate the fixture; nothing reads it at test/bench time.
"""
import hashlib
import os
import re
import sys

import numpy as np

SRC = "/root/reference/Assets/Textures/DisplacementTexture.cubemap"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                   "volumetric-particles-for-unity_b200", "assets", "displacement_r8.bin")
EXPECTED_SHA256 = "b064481200bbb407c089caf8238cf0d8e0fe30c31a66f310c98c696c422929a3"


def main():
    text = open(SRC).read()
    blob = bytes.fromhex(re.search(r"_typelessdata: ([0-9a-f]+)", text).group(1))
    argb = np.frombuffer(blob, dtype=np.uint8).reshape(6, 128, 128, 4)
    assert (argb[..., 0] == 255).all(), "byte 0 must be alpha = 255 (ARGB32)"
    r8 = np.ascontiguousarray(argb[..., 1])
    digest = hashlib.sha256(r8.tobytes()).hexdigest()
    assert digest == EXPECTED_SHA256, digest
    r8.tofile(DST)
    print("wrote", os.path.normpath(DST), r8.shape, "sha256", digest)


if __name__ == "__main__":
    sys.exit(main())
