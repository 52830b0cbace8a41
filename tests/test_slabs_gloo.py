"""Multi-GPU host logic on CPU: two gloo ranks run the product's SlabRenderer over the oracle.

Checks that the slab decomposition reproduces the single-context result: the volume and the light
sheet bit for bit (the sheet hand-off is exact), the image to rounding (ordered compositing of
slab partials is the reference's blend sequence re-associated)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vpe_b200 import scenes, slabs
from parity import max_rel_err


def test_partition_helpers():
    for nz, world in [(8, 1), (8, 2), (8, 3), (32, 8), (10, 4), (5, 5)]:
        ranges = [slabs.slab_range(nz, world, r) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == nz
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        assert all(b > a for a, b in ranges)
    with pytest.raises(ValueError):
        slabs.slab_range(4, 8, 0)
    assert slabs.row_bands(8, 3) == [(0, 2), (2, 5), (5, 8)]
    assert slabs.row_bands(4, 16) == [(0, 1), (1, 2), (2, 3), (3, 4)]
    assert slabs.row_bands(7, 1) == [(0, 7)]
    bands = [slabs.image_band(1080, 8, r) for r in range(8)]
    assert bands[0] == (0, 135) and bands[-1] == (945, 1080)
    bands = [slabs.image_band(10, 4, r) for r in range(4)]
    assert bands == [(0, 3), (3, 6), (6, 9), (9, 10)]


def _scene():
    sc = scenes.make_scene("cfg1", image=(48, 40))
    # camera inside the grid, looking sideways: both blend phases and every slab contribute
    sc["camera"]["position"] = (0.3, 0.2, 0.1)
    sc["camera"]["rotation"] = (0.0, 0.6427876, 0.0, 0.7660444)
    return sc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rebalance_worker(rank, world, port, out_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle_slab import OracleSlabEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = _scene()
        eng = OracleSlabEngine(sc, rank, world)
        r = slabs.SlabRenderer(eng, dist, fill_bands=2)
        r.profile = True
        r.fill(sc["particles"], sc["emitter"])
        r.march(sc["camera"], gather=False)
        new_slabs = r.rebalance()
        r.profile = False
        assert tuple(eng.slab) == tuple(new_slabs[rank]) == (r.z0, r.z1)
        r.fill(sc["particles"], sc["emitter"])
        img, total = r.march(sc["camera"])
        np.savez(os.path.join(out_dir, "rb%d.npz" % rank), slabs=np.asarray(new_slabs), total=np.int64(total),
                 img=img.numpy() if img is not None else np.zeros(0, np.float32), sheet=eng.eng.read_light_sheet())
    finally:
        dist.destroy_process_group()


def test_rebalanced_slabs_render_the_same_image(tmp_path):
    """SlabRenderer.rebalance over gloo: every rank derives the same contiguous partition from the gathered costs,
    moves its slab, and the frame rendered with the new slabs is the single-context frame."""
    from oracle_lib import oracle_engine
    world = 3
    sc = _scene()
    ref = oracle_engine(sc)
    scenes.apply_scene(ref, sc)
    ref.fill(sc["particles"], sc["emitter"])
    img_ref, smp_ref = ref.march(sc["camera"])
    mp.spawn(_rebalance_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(os.path.join(str(tmp_path), "rb%d.npz" % r)) for r in range(world)]
    parts = [tuple(map(tuple, o["slabs"])) for o in outs]
    assert parts[0] == parts[1] == parts[2]
    cuts = parts[0]
    assert cuts[0][0] == 0 and cuts[-1][1] == ref.grid[2] and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
    assert all(b > a for a, b in cuts)
    assert all(int(o["total"]) == int(smp_ref.sum()) for o in outs)
    assert np.array_equal(outs[-1]["sheet"], ref.read_light_sheet())
    assert max_rel_err(outs[0]["img"], img_ref) <= 1e-5


def _worker(rank, world, port, out_dir, bands):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle_slab import OracleSlabEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = _scene()
        eng = OracleSlabEngine(sc, rank, world)
        r = slabs.SlabRenderer(eng, dist, fill_bands=bands)
        r.fill(sc["particles"], sc["emitter"])
        img, total = r.march(sc["camera"])
        # the same frame again, every rank copying its band into one shared host image instead of gathering on rank 0
        host = slabs.SharedHostImage(dist, sc["camera"]["height"], sc["camera"]["width"])
        r.march(sc["camera"], count_samples=False, host_image=host)
        dist.barrier()
        if rank == 0:
            assert np.array_equal(host.array, img.numpy()), "shared host image differs from the gathered image"
        dist.barrier()
        host.close()
        z0, z1 = eng.slab
        bricks = {}
        for z in range(z0, z1):
            for y in range(eng.grid[1]):
                for x in range(eng.grid[0]):
                    b = eng.eng.read_brick(x, y, z)
                    if b is not None:
                        bricks["%d_%d_%d" % (x, y, z)] = b
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), sheet=eng.eng.read_light_sheet(), total=np.int64(total),
                 img=img.numpy() if img is not None else np.zeros(0, np.float32), **bricks)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,bands", [(2, 4), (3, 8)])
def test_slab_renderer_matches_single_context(tmp_path, world, bands):
    from oracle_lib import oracle_engine
    sc = _scene()
    ref = oracle_engine(sc)
    scenes.apply_scene(ref, sc)
    ref.fill(sc["particles"], sc["emitter"])
    img_ref, smp_ref = ref.march(sc["camera"])
    assert 0 <= ref.stats()["zBoundary"] < 7

    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), bands), nprocs=world, join=True)
    outs = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    # volume: every covered metavoxel is held by exactly one rank, bit-exact
    n_cov = 0
    gx, gy, gz = ref.grid
    for z in range(gz):
        owner = [r for r in range(world) if slabs.slab_range(gz, world, r)[0] <= z < slabs.slab_range(gz, world, r)[1]]
        assert len(owner) == 1
        for y in range(gy):
            for x in range(gx):
                b = ref.read_brick(x, y, z)
                key = "%d_%d_%d" % (x, y, z)
                assert (b is None) == (key not in outs[owner[0]].files)
                if b is not None:
                    n_cov += 1
                    assert np.array_equal(b, outs[owner[0]][key])
    assert n_cov == ref.stats()["numMetavoxelsCovered"]
    # the last slab ends with the single-context sheet
    assert np.array_equal(outs[-1]["sheet"], ref.read_light_sheet())
    # image on rank 0, total samples everywhere
    assert all(int(o["total"]) == int(smp_ref.sum()) for o in outs)
    img = outs[0]["img"]
    assert img.shape == img_ref.shape
    assert max_rel_err(img, img_ref) <= 1e-5
    assert float(img_ref[..., 3].max()) > 0.3
