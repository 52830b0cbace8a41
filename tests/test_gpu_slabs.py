"""GPU: the light-axis slab entry points of the C-ABI (vpe_fill_prepare / vpe_fill_region /
vpe_march_partial_device / vpe_composite_device) against the single-context result.

test_two_slab_contexts_on_one_gpu needs one GPU (two contexts share it, the sheet rows are copied
device to device exactly as NCCL would move them); test_nccl_two_ranks needs two."""
import os
import socket

import numpy as np
import pytest

import vpe_b200
from vpe_b200 import scenes, slabs
from oracle_lib import oracle_engine
from parity import RTOL, max_rel_err

pytestmark = pytest.mark.gpu


def _scene():
    sc = scenes.make_scene("cfg1", image=(96, 80))
    sc["camera"]["position"] = (0.3, 0.2, 0.1)
    sc["camera"]["rotation"] = (0.0, 0.6427876, 0.0, 0.7660444)
    return sc


def test_two_slab_contexts_on_one_gpu():
    import torch
    sc = _scene()
    cam = sc["camera"]
    h, w = cam["height"], cam["width"]
    one = vpe_b200.engine_for_scene(None, sc)
    scenes.apply_scene(one, sc)
    one.fill(sc["particles"], sc["emitter"])
    img_one, smp_one = one.march(cam)

    ranks = [slabs.CudaSlabEngine(sc, r, 2, 0) for r in range(2)]
    bands = slabs.row_bands(ranks[0].grid[1], 4)
    n = ranks[0].N
    gx = ranks[0].grid[0]
    for e in ranks:
        e.fill_prepare(sc["particles"], sc["emitter"])
        e.fill_density()
    for (y0, y1) in bands:  # the pipeline of SlabRenderer.fill, with a device copy in place of NCCL
        ranks[0].fill_sweep_region(0, gx, y0, y1)
        ranks[1].sheet_tensor()[y0 * n:y1 * n].copy_(ranks[0].sheet_tensor()[y0 * n:y1 * n])
        ranks[1].fill_sweep_region(0, gx, y0, y1)
    torch.cuda.synchronize()
    # volume: bit-exact with the single context, each brick on its owner
    cov = 0
    for z in range(ranks[0].grid[2]):
        owner = ranks[0] if z < ranks[0].slab[1] else ranks[1]
        for y in range(ranks[0].grid[1]):
            for x in range(gx):
                a, b = one.read_brick(x, y, z), owner.eng.read_brick(x, y, z)
                assert (a is None) == (b is None)
                if a is not None:
                    cov += 1
                    assert np.array_equal(a, b)
    assert cov == one.stats()["numMetavoxelsCovered"]
    assert np.array_equal(ranks[1].eng.read_light_sheet(), one.read_light_sheet())
    # march: slab partials composited in slab order == single-context image (to rounding)
    parts, total = [], 0
    for e in ranks:
        over, under = e.march_partial(cam)
        parts += [over, under]
        total += e.last_ray_samples()
    out = ranks[0].composite([p.clone() for p in parts], h * w).reshape(h, w, 4).cpu().numpy()
    assert total == int(smp_one.sum())
    assert max_rel_err(out, img_one) <= 1e-5
    ref = oracle_engine(sc)
    scenes.apply_scene(ref, sc)
    ref.fill(sc["particles"], sc["emitter"])
    img_ref, _ = ref.march(cam)
    assert max_rel_err(out, img_ref) <= RTOL


def test_linked_sweep_two_contexts_on_one_gpu():
    """vpe_fill_sweep_linked: the sweep kernel itself hands the sheet to the next slab (peer stores + a flag
    per block of voxel columns). Two contexts of one process share the GPU here; the upstream kernel is
    launched first. Three fills in a row exercise the epoch and acknowledge flags. Result: bit-identical
    to the single-context fill."""
    import torch
    sc = _scene()
    one = vpe_b200.engine_for_scene(None, sc)
    scenes.apply_scene(one, sc)
    one.fill(sc["particles"], sc["emitter"])
    ranks = [slabs.CudaSlabEngine(sc, r, 2, 0) for r in range(2)]
    ptrs = [e.eng.sheet_link_create()[1] for e in ranks]
    ranks[0].eng.sheet_link_connect(None, ptrs[1])
    ranks[1].eng.sheet_link_connect(ptrs[0], None)
    gx, gy, gz = ranks[0].grid
    for it in range(3):
        for e in ranks:
            e.fill_prepare(sc["particles"], sc["emitter"])
            e.fill_density()
        torch.cuda.synchronize()
        for e in ranks:  # upstream first: on one GPU the downstream kernel must not occupy the SMs alone
            e.fill_sweep_linked()
            if it == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        assert ranks[0].eng.sheet_link_timeouts() == 0 and ranks[1].eng.sheet_link_timeouts() == 0
        assert np.array_equal(ranks[1].eng.read_light_sheet(), one.read_light_sheet())

    def check_volume():
        cov = 0
        for z in range(gz):
            owner = ranks[0] if z < ranks[0].slab[1] else ranks[1]
            for y in range(gy):
                for x in range(gx):
                    a, b = one.read_brick(x, y, z), owner.eng.read_brick(x, y, z)
                    assert (a is None) == (b is None)
                    if a is not None:
                        cov += 1
                        assert np.array_equal(a, b)
        assert cov == one.stats()["numMetavoxelsCovered"]
    check_volume()
    # load balancing moves the slab boundary between fills (SlabRenderer.rebalance): same volume again
    ranks[0].set_slab(0, 3)
    ranks[1].set_slab(3, gz)
    with pytest.raises(vpe_b200.VpeError):
        ranks[1].march_partial(sc["camera"])  # the old slab's volume is gone: fill first
    for e in ranks:
        e.fill_prepare(sc["particles"], sc["emitter"])
        e.fill_density()
    for e in ranks:
        e.fill_sweep_linked()
    torch.cuda.synchronize()
    assert np.array_equal(ranks[1].eng.read_light_sheet(), one.read_light_sheet())
    check_volume()


@pytest.mark.parametrize("n_vox", [32, 12])
def test_head_rank_runs_the_fused_kernel_and_feeds_the_link(n_vox):
    """vpe_fill_linked: the rank nearest the light has nothing upstream, so it does not split its fill - the fused
    kernel (k_fill_columns<false, ., HEAD>) stores its exit sheet into the next rank's inbox and raises the flags the
    linked sweep would; the other ranks run density + linked sweep. Three contexts of one process, three frames
    (epochs, acknowledge flags): volume, final sheet and the image are those of the single-context fill; the switch
    noHeadFused gives the same bits through the split path."""
    import torch
    sc = _scene()
    sc["numVoxels"] = n_vox
    cam = sc["camera"]
    h, w = cam["height"], cam["width"]
    one = vpe_b200.engine_for_scene(None, sc)
    scenes.apply_scene(one, sc)
    one.fill(sc["particles"], sc["emitter"])
    img_one, smp_one = one.march(cam)
    ranks = [slabs.CudaSlabEngine(sc, r, 3, 0) for r in range(3)]
    ptrs = [e.eng.sheet_link_create()[1] for e in ranks]
    for r, e in enumerate(ranks):
        e.eng.sheet_link_connect(ptrs[r - 1] if r > 0 else None, ptrs[r + 1] if r < 2 else None)
    gx, gy, gz = ranks[0].grid
    for it in range(4):
        if it == 3:
            ranks[0].eng.set_debug_options(no_head_fused=True)
        for e in ranks:
            e.fill_prepare(sc["particles"], sc["emitter"])
        for e in ranks:  # upstream first (one GPU)
            e.fill_linked()
        torch.cuda.synchronize()
        st = ranks[0].eng.stats()
        assert st["fillLaunches"] >= 1
        assert all(e.eng.sheet_link_timeouts() == 0 for e in ranks), ranks[0].eng.lib.vpe_last_error(ranks[0].eng._ctx)
        assert np.array_equal(ranks[2].eng.read_light_sheet(), one.read_light_sheet())
        cov = 0
        for z in range(gz):
            owner = [e for e in ranks if e.slab[0] <= z < e.slab[1]][0]
            for y in range(gy):
                for x in range(gx):
                    a, b = one.read_brick(x, y, z), owner.eng.read_brick(x, y, z)
                    assert (a is None) == (b is None)
                    if a is not None:
                        cov += 1
                        assert np.array_equal(a, b)
        assert cov == one.stats()["numMetavoxelsCovered"]
        parts, total = [], 0
        for e in ranks:  # the head's empty-space bitmap comes from the fused kernel
            over, under = e.march_partial(cam)
            parts += [over, under]
            total += e.last_ray_samples()
        out = ranks[0].composite([p.clone() for p in parts], h * w).reshape(h, w, 4).cpu().numpy()
        assert total == int(smp_one.sum())
        assert max_rel_err(out, img_one) <= 1e-5


@pytest.mark.parametrize("n_vox,grid,ranks_z", [(64, (2, 2, 4), [(0, 1), (1, 4)]), (12, (3, 3, 5), [(0, 2), (2, 3), (3, 5)]),
                                               (32, (2, 3, 3), [(0, 1), (1, 2), (2, 3)])])
def test_linked_sweep_other_brick_sizes_and_three_slabs(n_vox, grid, ranks_z):
    """The linked sweep with 64^3 bricks (16 blocks of voxel columns per metavoxel column), with a brick edge that is
    not a multiple of the 8x4 warp tile (idle lanes must still take part in the block's hand-shake) and through a chain
    of three contexts: bricks and sheet bit-identical to the single-context fill."""
    import torch
    sc = scenes.make_scene("cfg1", image=(48, 48))
    sc["grid"], sc["numVoxels"] = tuple(grid), n_vox
    rng = np.random.default_rng(77)
    half = (np.asarray(grid, dtype=np.float64) / 2 - 0.6) * sc["mvScale"]
    p = scenes.make_particles_uniform(rng, 14, 8, sc["mvScale"])
    p[:, 0:3] = scenes.quat_rotate(scenes.LIGHT_ROTATION, rng.uniform(-half, half, size=(14, 3))).astype(np.float32)
    sc["particles"] = p
    one = vpe_b200.engine_for_scene(None, sc)
    scenes.apply_scene(one, sc)
    one.fill(sc["particles"], sc["emitter"])
    world = len(ranks_z)
    ranks = [slabs.CudaSlabEngine(sc, r, world, 0) for r in range(world)]
    for e, (z0, z1) in zip(ranks, ranks_z):
        e.set_slab(z0, z1)
    ptrs = [e.eng.sheet_link_create()[1] for e in ranks]
    for r, e in enumerate(ranks):
        e.eng.sheet_link_connect(ptrs[r - 1] if r > 0 else None, ptrs[r + 1] if r < world - 1 else None)
    for it in range(2):
        for e in ranks:
            e.fill_prepare(sc["particles"], sc["emitter"])
            e.fill_density()
        for e in ranks:
            e.fill_sweep_linked()
        torch.cuda.synchronize()
        assert all(e.eng.sheet_link_timeouts() == 0 for e in ranks)
    assert np.array_equal(ranks[-1].eng.read_light_sheet(), one.read_light_sheet())
    cov = 0
    for z in range(grid[2]):
        owner = next(e for e, (z0, z1) in zip(ranks, ranks_z) if z0 <= z < z1)
        for y in range(grid[1]):
            for x in range(grid[0]):
                a, b = one.read_brick(x, y, z), owner.eng.read_brick(x, y, z)
                assert (a is None) == (b is None)
                if a is not None:
                    cov += 1
                    assert np.array_equal(a, b)
    assert cov == one.stats()["numMetavoxelsCovered"] > 0


def test_image_link_two_contexts_on_one_gpu():
    """vpe_march_linked / vpe_composite_linked: the march kernel stores the partial images into the compositing
    context's receive buffer and raises a flag; compositing from there equals compositing the partial images
    (bit for bit: same kernel arithmetic), two frames in a row (both buffer parities)."""
    import torch
    sc = _scene()
    cam = sc["camera"]
    h, w = cam["height"], cam["width"]
    per = -(-h // 2)
    ranks = [slabs.CudaSlabEngine(sc, r, 2, 0) for r in range(2)]
    for e in ranks:
        e.fill_prepare(sc["particles"], sc["emitter"])
    # fill through the two-phase path with a device copy of the sheet (as in test_two_slab_contexts_on_one_gpu)
    for e in ranks:
        e.fill_density()
    ranks[0].fill_sweep_region(0, ranks[0].grid[0], 0, ranks[0].grid[1])
    ranks[1].sheet_tensor().copy_(ranks[0].sheet_tensor())
    ranks[1].fill_sweep_region(0, ranks[1].grid[0], 0, ranks[1].grid[1])
    parts = []
    for e in ranks:
        over, under = e.march_partial(cam, per * 2)
        parts += [over.clone(), under.clone()]
    want = ranks[0].composite(parts, per * 2 * w).reshape(per * 2, w, 4).cpu().numpy()
    ptrs = [e.eng.image_link_create(2, r, w, h)[1] for r, e in enumerate(ranks)]
    for r, e in enumerate(ranks):
        e.eng.image_link_connect([None if q == r else ptrs[q] for q in range(2)])
    for frame in range(3):
        for e in ranks:
            e.march_linked(cam)
        bands = [e.composite_linked(per, w).cpu().numpy() for e in ranks]
        torch.cuda.synchronize()
        assert all(e.eng.image_link_timeouts() == 0 for e in ranks)
        got = np.concatenate(bands, axis=0)
        assert np.array_equal(got, want), frame
    assert float(want[..., 3].max()) > 0.3


def test_linked_sweep_reports_a_missing_peer():
    """A downstream rank whose upstream never arrives gives up after the spin limit and says so; it does
    not hang the GPU."""
    import torch
    sc = _scene()
    ranks = [slabs.CudaSlabEngine(sc, r, 2, 0) for r in range(2)]
    ranks[1].eng.set_debug_options(link_spin_ms=20)
    ptrs = [e.eng.sheet_link_create()[1] for e in ranks]
    ranks[1].eng.sheet_link_connect(ptrs[0], None)
    ranks[1].fill_prepare(sc["particles"], sc["emitter"])
    ranks[1].fill_density()
    ranks[1].fill_sweep_linked()
    torch.cuda.synchronize()
    assert ranks[1].eng.sheet_link_timeouts() > 0


def _nccl_worker(rank, world, port, out_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        sc = _scene()
        eng = slabs.CudaSlabEngine(sc, rank, world, rank)
        r = slabs.SlabRenderer(eng, dist, fill_bands=4)   # NCCL send/recv band pipeline
        r.fill(sc["particles"], sc["emitter"])
        img, total = r.march(sc["camera"])
        torch.cuda.synchronize()
        sheet_nccl = eng.eng.read_light_sheet()
        # the same over the sheet link (peer memory, CUDA IPC between the two processes), twice
        eng2 = slabs.CudaSlabEngine(sc, rank, world, rank)
        r2 = slabs.SlabRenderer(eng2, dist)
        assert r2.linked and r2._image_link(sc["camera"]["width"], sc["camera"]["height"])
        for _ in range(2):
            r2.fill(sc["particles"], sc["emitter"])
        img2, total2 = r2.march(sc["camera"])
        torch.cuda.synchronize()
        assert eng2.eng.sheet_link_timeouts() == 0 and eng2.eng.image_link_timeouts() == 0
        assert np.array_equal(eng2.eng.read_light_sheet(), sheet_nccl)
        if rank == 0:
            assert total2 == total and np.array_equal(img2.cpu().numpy(), img.cpu().numpy())
            np.savez(os.path.join(out_dir, "nccl.npz"), img=img.cpu().numpy(), total=np.int64(total))
    finally:
        dist.destroy_process_group()


def test_nccl_two_ranks(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    out = np.load(os.path.join(str(tmp_path), "nccl.npz"))
    sc = _scene()
    ref = oracle_engine(sc)
    scenes.apply_scene(ref, sc)
    ref.fill(sc["particles"], sc["emitter"])
    img_ref, smp_ref = ref.march(sc["camera"])
    assert int(out["total"]) == int(smp_ref.sum())
    assert max_rel_err(out["img"], img_ref) <= RTOL


def test_overlapped_sweep_many_blocks_back_to_back_frames():
    """The persistent sweep kernel (k_sweep_overlapped) runs concurrently with the density pass of the same fill. With more
    blocks of voxel columns than SMs every sweep CTA takes several blocks; several frames are enqueued without a host
    synchronisation in between, as bench.py does. Two contexts share the GPU; result: the single-context volume, bit for bit."""
    import torch
    sc = scenes.make_scene("cfg1", image=(64, 48))
    sc["grid"] = (10, 9, 4)
    sc["numVoxels"] = 32
    rng = np.random.default_rng(5)
    n = 160
    p = np.zeros((n, 7), dtype=np.float32)
    p[:, 0:3] = rng.uniform(-3.5, 3.5, size=(n, 3)) * np.array([1.0, 1.0, 0.4])
    p[:, 3] = rng.uniform(1.2, 2.4, size=n)
    p[:, 4] = rng.uniform(0, 360, size=n)
    p[:, 5] = rng.uniform(0.1, 6.0, size=n)
    p[:, 6] = 6.0
    sc["particles"] = p
    one = vpe_b200.engine_for_scene(None, sc)
    scenes.apply_scene(one, sc)
    one.fill(sc["particles"], sc["emitter"])
    assert one.stats()["numMetavoxelsCovered"] > 100
    ranks = [slabs.CudaSlabEngine(sc, r, 2, 0) for r in range(2)]
    for e in ranks:
        e.eng.set_debug_options(link_spin_ms=300, sweep_overlap=True)
    ptrs = [e.eng.sheet_link_create()[1] for e in ranks]
    ranks[0].eng.sheet_link_connect(None, ptrs[1])
    ranks[1].eng.sheet_link_connect(ptrs[0], None)
    for frame in range(4):
        for e in ranks:
            e.fill_prepare(sc["particles"], sc["emitter"])
            e.fill_density()
            e.fill_sweep_linked()
    torch.cuda.synchronize()
    for e in ranks:
        assert e.eng.sheet_link_timeouts() == 0, e.eng.lib.vpe_last_error(e.eng._ctx)
    assert np.array_equal(ranks[1].eng.read_light_sheet(), one.read_light_sheet())
    gx, gy, gz = ranks[0].grid
    for z in range(gz):
        owner = ranks[0] if z < ranks[0].slab[1] else ranks[1]
        for y in range(gy):
            for x in range(gx):
                a, b = one.read_brick(x, y, z), owner.eng.read_brick(x, y, z)
                assert (a is None) == (b is None)
                if a is not None:
                    assert np.array_equal(a, b)


@pytest.mark.parametrize("n_vox,tma", [(32, True), (64, True), (32, False)])
def test_density_plus_sweep_equals_the_fused_fill(n_vox, tma):
    """vpe_fill_density + vpe_fill_sweep_region == vpe_fill, bit for bit, for the brick sizes the TMA sweep pipeline
    (k_sweep_tma: cp.async.bulk.tensor loads/stores of [8 slices][rows][N] boxes, mbarrier ring) handles, and for the
    register-staged kernel it replaces (VpeDebugOptions.noTmaSweep)."""
    sc = scenes.make_scene("cfg1", image=(32, 32))
    sc["grid"] = (3, 2, 3)
    sc["numVoxels"] = n_vox
    a = vpe_b200.engine_for_scene(None, sc)
    b = vpe_b200.engine_for_scene(None, sc)
    if not tma:
        b.set_debug_options(no_tma_sweep=True)
    for e in (a, b):
        scenes.apply_scene(e, sc)
    p = sc["particles"].copy()
    p[:, 0:3] *= 0.35
    a.fill(p, sc["emitter"])
    for frame in range(2):
        b.fill_prepare(p, sc["emitter"])
        b.fill_density()
        b.fill_sweep_region(0, 3, 0, 2)
    assert np.array_equal(a.read_light_sheet(), b.read_light_sheet())
    cov = 0
    for z in range(3):
        for y in range(2):
            for x in range(3):
                ba, bb = a.read_brick(x, y, z), b.read_brick(x, y, z)
                assert (ba is None) == (bb is None)
                if ba is not None:
                    cov += 1
                    assert np.array_equal(ba, bb), (x, y, z)
    assert cov >= 4
