"""Golden fixtures (tests/golden/*.npz, written by tools/make_golden.py from the CPU oracle).

CPU: the oracle still reproduces them bit for bit (it cannot drift silently).
GPU (-m gpu): the CUDA engine against the committed numbers — discrete results, light sheet and
the sha256 of the whole fp16 volume are exact; RGBA within the parity metric.  For cfg2 this is a
full-size (3 526 bricks, 115 M voxels) bit-exact volume check that needs no oracle run on the box."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import vpe_b200
import make_golden
from oracle_lib import oracle_engine
from parity import RTOL, max_rel_err

CASES = sorted(make_golden.CASES)


def load(name):
    return np.load(os.path.join(make_golden.GOLDEN, name + ".npz"))


def check_discrete(out, gold):
    for k in ("covered", "pairs", "ray_samples", "z_boundary"):
        assert int(out[k]) == int(gold[k]), k
    assert np.array_equal(out["list_offsets"], gold["list_offsets"])
    assert np.array_equal(out["list_indices"], gold["list_indices"])
    assert np.array_equal(out["samples"], gold["samples"])
    if "pixels" in gold.files:
        assert np.array_equal(out["pixels"], gold["pixels"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(name):
    gold = load(name)
    out = make_golden.run_case(name, lambda sc, **kw: oracle_engine(sc, **kw))
    check_discrete(out, gold)
    assert bytes(out["volume_sha256"]) == bytes(gold["volume_sha256"])
    assert np.array_equal(out["sheet"], gold["sheet"])
    assert np.array_equal(out["rgba"], gold["rgba"])


def test_golden_figures_match_the_survey_model():
    """SURVEY.md §8d lists pair / covered-metavoxel counts predicted by an independent throwaway numpy
    model of VPR.cs:415-457 written during the survey; the oracle lands on the same integers."""
    g1, g1e, g2 = load("cfg1"), load("cfg1_exact_bins"), load("cfg2_subset")
    assert (int(g1["pairs"]), int(g1["covered"])) == (457, 218)
    assert int(g1e["pairs"]) == 685
    assert (int(g2["pairs"]), int(g2["covered"])) == (13000, 3526)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_golden(name):
    gold = load(name)
    out = make_golden.run_case(name, lambda sc, **kw: vpe_b200.engine_for_scene(None, sc, **kw))
    check_discrete(out, gold)
    assert bytes(out["volume_sha256"]) == bytes(gold["volume_sha256"]), "fp16 volume differs from the oracle's"
    assert np.array_equal(out["sheet"], gold["sheet"])
    assert max_rel_err(out["rgba"], gold["rgba"]) <= RTOL
