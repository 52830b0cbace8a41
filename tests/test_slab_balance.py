"""Slab load balancing (slabs.balance_slabs): contiguous light-axis slabs that minimise the modelled
frame time of the multi-GPU pipeline (density pass -> chained light sweep -> march)."""
import itertools

import numpy as np

from vpe_b200 import slabs


def frame_time(cuts, dc, sc, mc, prepare=0.1, lag=0.07, head=1.0):
    """The model of balance_slabs, evaluated for one partition (cuts = [(z0, z1), ...])."""
    sweep_prev, worst = -1e30, 0.0
    for i, (a, b) in enumerate(cuts):
        d = prepare + float(np.sum(dc[a:b]))
        se = max(d + float(np.sum(sc[a:b])) * (head if i == 0 else 1.0), sweep_prev + lag)
        worst = max(worst, se + float(np.sum(mc[a:b])))
        sweep_prev = se
    return worst


def brute_force(dc, sc, mc, world, head=1.0):
    nz = len(dc)
    best = None
    for inner in itertools.combinations(range(1, nz), world - 1):
        edges = (0,) + inner + (nz,)
        cuts = [(edges[i], edges[i + 1]) for i in range(world)]
        t = frame_time(cuts, dc, sc, mc, head=head)
        if best is None or t < best[0] - 1e-12:
            best = (t, cuts)
    return best


def test_partition_is_contiguous_and_complete():
    rng = np.random.default_rng(5)
    dc, sc, mc = rng.uniform(0.1, 1, 32), np.full(32, 0.05), rng.uniform(0.1, 1, 32)
    cuts, t = slabs.balance_slabs(dc, sc, mc, 8)
    assert cuts[0][0] == 0 and cuts[-1][1] == 32 and len(cuts) == 8
    assert all(a < b for a, b in cuts) and all(cuts[i][1] == cuts[i + 1][0] for i in range(7))
    assert abs(t - frame_time(cuts, dc, sc, mc)) < 1e-9


def test_matches_exhaustive_search_on_small_grids():
    rng = np.random.default_rng(11)
    for trial in range(6):
        nz, world = 10, int(rng.integers(2, 5))
        dc, sc, mc = rng.uniform(0, 1, nz), rng.uniform(0, 0.2, nz), rng.uniform(0, 1, nz)
        cuts, t = slabs.balance_slabs(dc, sc, mc, world)
        t_best, _ = brute_force(dc, sc, mc, world)
        assert abs(t - t_best) < 1e-9, (trial, cuts)


def test_uniform_costs_without_a_chain_give_equal_slabs():
    cuts, _ = slabs.balance_slabs(np.ones(32), np.zeros(32), np.ones(32), 8, prepare=0.0, lag=0.0)
    assert cuts == [(4 * i, 4 * i + 4) for i in range(8)]


def test_march_heavy_slices_near_the_camera_get_thinner_slabs():
    """The trace of cfg3 at 8 GPUs: march cost falls with z (perspective), density is flat. Equal slabs
    leave the first ranks with the longest frames; the balanced partition is better under the model
    and gives the march-heavy end thinner slabs."""
    nz, world = 32, 8
    dc = np.full(nz, 1.35 / 4)
    mc = np.linspace(0.42, 0.12, nz)
    sc = np.full(nz, 0.12)
    equal = [slabs.slab_range(nz, world, r) for r in range(world)]
    cuts, t = slabs.balance_slabs(dc, sc, mc, world)
    assert t < frame_time(equal, dc, sc, mc) - 0.05
    assert (cuts[0][1] - cuts[0][0]) <= (cuts[-1][1] - cuts[-1][0])


def test_one_rank_and_as_many_ranks_as_slices():
    assert slabs.balance_slabs(np.ones(5), np.ones(5), np.ones(5), 1)[0] == [(0, 5)]
    assert slabs.balance_slabs(np.ones(5), np.ones(5), np.ones(5), 5)[0] == [(i, i + 1) for i in range(5)]


def test_head_rank_with_the_fused_kernel_takes_more_slices():
    """The rank nearest the light runs the fused fill kernel (its sweep costs a fraction of a separate sweep kernel):
    the model scales slab 0's sweep cost; exact against exhaustive search, and slab 0 never gets thinner."""
    rng = np.random.default_rng(3)
    for trial in range(4):
        nz, world = 10, int(rng.integers(2, 4))
        dc, sc, mc = rng.uniform(0.2, 1, nz), rng.uniform(0.2, 0.5, nz), rng.uniform(0, 0.5, nz)
        cuts, t = slabs.balance_slabs(dc, sc, mc, world, head_sweep_factor=0.35)
        t_best, _ = brute_force(dc, sc, mc, world, head=0.35)
        assert abs(t - t_best) < 1e-9
        assert t <= slabs.balance_slabs(dc, sc, mc, world)[1] + 1e-12
