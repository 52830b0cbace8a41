"""Evidence behind decision C-1 of DESIGN.md §2 (orientation of the displacement cubemap): the stored faces are the loose
source pictures mirrored left-right - never flipped vertically, never rotated - and the stored arrangement is not a
seamless cube map under any reading of rows / columns, while the source pictures in their own order are."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import check_cubemap_orientation as cco


def test_seams_do_not_discriminate_between_the_readings():
    inside, seams = cco.seam_report()
    stored = seams["as stored (the engine's reading: direct upload)"][0]
    mirrored = seams["every face mirrored back (= the source pictures in the stored slots)"][0]
    source_order = seams["source pictures c00..c05 as +X,-X,+Y,-Y,+Z,-Z, un-mirrored"]
    assert source_order[0] < inside and source_order[1] < 1.5 * inside     # the artist's set is a seamless cube map
    assert stored > 5 * inside and mirrored > 5 * inside                   # the asset is not, whichever way it is read
    assert abs(stored - mirrored) < 0.1 * stored


@pytest.mark.skipif(not os.path.isdir(cco.REF), reason="needs the reference's source pictures (not on the GPU box)")
def test_stored_faces_are_the_source_pictures_mirrored_left_right():
    pytest.importorskip("PIL")
    rows = cco.analyse()
    assert [best[1] for (_, best, _) in rows] == [2, 3, 4, 5, 0, 1]
    for (_, best, second) in rows:
        assert best[2] == "mirror left-right" and best[0] < 3.0 and second[0] > 10 * best[0]
