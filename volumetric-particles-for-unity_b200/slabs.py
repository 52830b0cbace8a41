"""Multi-GPU host logic: light-axis slabs for the fill, slab-local march + ordered compositing.

One process per GPU (torch.distributed; NCCL over NVLink on the GPU box, gloo in the CPU tests).
Rank r owns the metavoxel slices z in [r*NZ/R, (r+1)*NZ/R) — slice 0 is nearest the light
(VPR.cs:388-390, 505) — and keeps their bricks in its own HBM; bricks never cross the link.

Fill (≙ FillMetavoxels, VPR.cs:495-520).  The only dependency between slabs is the light sheet
(lightPropogationUAV, VPR.cs:266; Fill.shader:224,250): slab r+1 needs the sheet as slab r left it.
So the fill is split there.  Phase 1 (vpe_fill_density): every rank runs the particle loop of its
whole slab — density and ambient-occlusion term of every voxel, the dominant cost, no light involved
— at once.  Phase 2 (vpe_fill_sweep_region): the light sweep, an HBM-bound pass.  The metavoxel
columns are cut into bands of rows; for every band, in order, a rank receives the band's sheet rows
from rank r-1, sweeps the band through its own slices and sends the rows on to rank r+1.  Bands
pipeline through the ranks: after R-1 bands every GPU is busy.  Per band the message is
(rows*N) x (NX*N) fp32 — the exit-plane light sheet of north_star.

March (≙ RenderMetavoxels, VPR.cs:637-713).  The reference composites metavoxels slice-major:
slices 0..zB far-to-near with OVER, then zB+1..NZ-1 near-to-far with UNDER (VPR.cs:652-711).  A
slab's slices are therefore one contiguous run of each phase, and because both blend equations are
associative, a rank can march its own bricks into two premultiplied partial images (OVER part,
UNDER part; vpe_march_partial_device) and the partials can be composited afterwards in slab order
(vpe_composite_device).  The image is cut into R bands of rows; one all-to-all moves every rank's
partial rows to the band's owner, which composites them; rank 0 gathers the bands.

The engine adapter is what touches memory: `CudaSlabEngine` (below) wraps the CUDA library and
device tensors; the CPU tests drive the very same `SlabRenderer` with an adapter over the oracle.
"""
import numpy as np


# ------------------------------------------------------------------------------------------------
# pure partitioning helpers (unit-tested on CPU)
# ------------------------------------------------------------------------------------------------
def slab_range(num_slices, world, rank):
    """Slices [z0, z1) owned by `rank`: contiguous, ascending with rank, sizes differ by at most 1."""
    if not (0 <= rank < world) or world > num_slices:
        raise ValueError("need 0 <= rank < world <= numMetavoxelsZ (got rank %d, world %d, NZ %d)" % (rank, world, num_slices))
    return rank * num_slices // world, (rank + 1) * num_slices // world


def row_bands(num_rows, bands):
    """Cut [0, num_rows) into at most `bands` contiguous non-empty bands."""
    bands = max(1, min(int(bands), num_rows))
    return [(b * num_rows // bands, (b + 1) * num_rows // bands) for b in range(bands)]


def image_band(height, world, rank):
    """Image rows [r0, r1) composited by `rank`; every rank gets ceil(H/R) rows, the last may be short."""
    per = -(-height // world)
    return min(rank * per, height), min((rank + 1) * per, height)


# Cost of the sweep inside the fused fill kernel relative to the sweep as a kernel of its own (cfg3, one B200: fused 9.24 ms,
# density pass 8.21 ms, sweep alone 3.09 ms): what the rank at the head of the chain pays for its sweep (profiles/r02_scaling.md).
HEAD_FUSED_SWEEP_FACTOR = 0.35


def balance_slabs(density_cost, sweep_cost, march_cost, world, prepare=0.1, lag=0.07, head_sweep_factor=1.0):
    """Contiguous light-axis slabs [z0, z1) per rank that minimise the modelled frame time.

    Per-slice costs (ms): density_cost = the particle loop, sweep_cost = the light sweep, march_cost = the
    ray march. Model of one frame on rank r owning slices [a, b) (what tools/slab_trace.py shows):
        density ends   d_r = prepare + sum(density_cost[a:b])
        sweep ends     s_r = max(d_r + sum(sweep_cost[a:b]), s_{r-1} + lag)     (the sheet chain)
        march ends     e_r = s_r + sum(march_cost[a:b])
    and the frame ends at max_r e_r (the compositing exchange waits for every rank). head_sweep_factor scales the sweep
    cost of rank 0, which may run the fused kernel (HEAD_FUSED_SWEEP_FACTOR) instead of density + sweep. The sweep of a rank
    cannot end before that of the rank nearer the light, so ranks late in the chain should carry less
    march work; exact search by dynamic programming over (rank, first slice) with Pareto-pruned states."""
    nz = len(density_cost)
    if not (1 <= world <= nz):
        raise ValueError("need 1 <= world <= number of slices")
    cd = np.concatenate([[0.0], np.cumsum(np.asarray(density_cost, dtype=np.float64))])
    cs = np.concatenate([[0.0], np.cumsum(np.asarray(sweep_cost, dtype=np.float64))])
    cm = np.concatenate([[0.0], np.cumsum(np.asarray(march_cost, dtype=np.float64))])
    # states[start] = list of (sweep_end_prev, worst_end_so_far, cuts) not dominated by another
    states = {0: [(-1e30, 0.0, ())]}
    for r in range(world):
        nxt = {}
        remaining = world - r - 1
        for a, lst in states.items():
            hi = nz - remaining
            ends = [nz] if remaining == 0 else range(a + 1, hi + 1)
            for b in ends:
                d = prepare + cd[b] - cd[a]
                sw, ma = cs[b] - cs[a], cm[b] - cm[a]
                if r == 0:
                    sw *= head_sweep_factor
                for (sp, worst, cuts) in lst:
                    se = max(d + sw, sp + lag)
                    cand = (se, max(worst, se + ma), cuts + (b,))
                    cur = nxt.setdefault(b, [])
                    if any(o[0] <= cand[0] and o[1] <= cand[1] for o in cur):
                        continue
                    cur[:] = [o for o in cur if not (cand[0] <= o[0] and cand[1] <= o[1])]
                    cur.append(cand)
        states = nxt
    best = min(states[nz], key=lambda t: t[1])
    cuts = (0,) + best[2]
    return [(cuts[i], cuts[i + 1]) for i in range(world)], best[1]


def default_fill_bands(grid, n_voxels, world, min_ctas_per_launch=512):
    """Bands of metavoxel rows for the fill pipeline: as many as possible (the pipeline bubble is
    (R-1)/(T+R-1)) while one band's launch still has >= min_ctas_per_launch CTAs (k_fill_columns uses
    one CTA per 8 warp tiles of 8x4 voxel columns per metavoxel column)."""
    if world <= 1:
        return 1
    gx, gy, _ = grid
    ctas_per_row = gx * -(-(-(-n_voxels // 8) * -(-n_voxels // 4)) // 8)
    rows_per_band = max(1, -(-min_ctas_per_launch // ctas_per_row))
    return max(1, gy // rows_per_band)


class SlabRenderer:
    """fill() + march() of one rank.  `engine` is a slab engine adapter, `dist` is torch.distributed
    (already initialised) or None for a single process."""

    def __init__(self, engine, dist=None, fill_bands=None, image_link=True, head_fused=True):
        self.e = engine
        self.dist = dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        gx, gy, gz = engine.grid
        self.z0, self.z1 = slab_range(gz, self.world, self.rank)
        if (self.z0, self.z1) != tuple(engine.slab):
            raise ValueError("engine owns slab %s, rank %d of %d must own %s" % (engine.slab, self.rank, self.world, (self.z0, self.z1)))
        self.bands = row_bands(gy, fill_bands if fill_bands is not None else default_fill_bands(engine.grid, engine.N, self.world))
        # sheet link: engines that can hand the sheet over through peer memory (CUDA, one node) do so;
        # fill_bands given explicitly keeps the NCCL send/recv band pipeline; image_link=False keeps the
        # NCCL all-to-all of the partial images
        self._image_links = {}
        self._use_image_link = bool(image_link)
        self.profile = False          # record device times of the density pass and the march kernel (rebalance)
        self._times = None
        self.linked = False
        if self.world > 1 and fill_bands is None and hasattr(engine, "link_neighbours"):
            self.linked = bool(engine.link_neighbours(dist, self.rank, self.world))
        # the head of the chain does not split its fill (engines that can: vpe_fill_linked)
        self.head_fused = bool(self.linked and self.rank == 0 and head_fused and hasattr(engine, "fill_linked"))

    # -- fill -------------------------------------------------------------------------------------
    def fill(self, particles, emitter):
        e, d = self.e, self.dist
        gx, gy, gz = e.grid
        n = e.N
        e.fill_prepare(particles, emitter)          # bins into this slab only, clears the sheet to 1
        if self.world == 1:
            e.fill_region(0, gx, 0, gy)             # fused: nothing to wait for
            return
        t0 = e.record_event() if self.profile else None
        if self.head_fused:
            # nothing upstream of the slab nearest the light: the fused kernel runs and feeds the sheet link itself
            e.fill_linked()
            if self.profile:
                self._times = [t0, e.record_event()]
            return
        e.fill_density()                            # phase 1: the particle loop of the whole slab, no dependency
        t1 = e.record_event() if self.profile else None
        if self.profile:
            self._times = [t0, t1]
        if self.linked:
            # phase 2 in ONE kernel per rank: the sweep hands its exit values to the next rank's inbox over
            # NVLink peer memory and raises a flag per block of voxel columns (k_sweep_columns<., true>)
            e.fill_sweep_linked()
            return
        sheet = e.sheet_tensor()                    # (NY*N, NX*N) fp32 view of the context's sheet
        for (y0, y1) in self.bands:                 # phase 2: the light sweep, pipelined through the ranks
            rows = sheet[y0 * n:y1 * n]
            if self.rank > 0:
                d.recv(rows, src=self.rank - 1)     # the sheet as the previous slab left it
                e.sheet_written(y0, y1)
            e.fill_sweep_region(0, gx, y0, y1)
            if self.rank < self.world - 1:
                e.sheet_read(y0, y1)
                d.send(rows, dst=self.rank + 1)

    def _image_link(self, w, h):
        """Set up (once per image size) the peer-memory exchange of the partial images; False = use NCCL."""
        if not self.linked or not self._use_image_link or not hasattr(self.e, "image_link_neighbours"):
            return False
        key = (w, h)
        if key not in self._image_links:
            self._image_links[key] = bool(self.e.image_link_neighbours(self.dist, self.rank, self.world, w, h))
        return self._image_links[key]

    # -- load balance -----------------------------------------------------------------------------
    def set_slab(self, z0, z1):
        self.e.set_slab(z0, z1)
        self.z0, self.z1 = z0, z1

    def rebalance(self):
        """Move the slab boundaries so that the modelled frame time (balance_slabs) is minimal, from the
        device times of the last profiled frame (self.profile = True during fill() and march()). A rank's density
        and march-kernel time is spread over its slices in proportion to the work the engine counted per slice
        ((particle, metavoxel) pairs of the fill, ray samples of the march; evenly where the engine has no such
        profile), which gives a per-slice cost profile of the whole grid; all ranks compute the same partition
        from the same gathered numbers. The volume must be filled again afterwards. Returns the new list of slabs."""
        e, d = self.e, self.dist
        if self.world == 1 or self._times is None:
            return [(self.z0, self.z1)]
        density_ms = e.elapsed_ms(self._times[0], self._times[1])
        march_ms = float(e.stats()["marchKernelMs"])
        nz = e.grid[2]
        wd, wm = np.ones(nz), np.ones(nz)
        prof = e.slice_profile() if hasattr(e, "slice_profile") else None
        if prof is not None:
            pairs, covered, samples = prof
            # a covered metavoxel costs its voxels' loop overhead even with few particles: ~2 pairs' worth (measured shape)
            wd = pairs.astype(np.float64) + 2.0 * covered.astype(np.float64) + 1e-9
            wm = samples.astype(np.float64) + 1e-9
        # the sweep moves 16 bytes per voxel of a covered metavoxel; ~0.6 of the HBM rate measured alone
        gx, gy, _ = e.grid
        sweep = np.full(nz, gx * gy * float(e.N) ** 3 * 16.0 / 4.0e12 * 1e3)
        if self.head_fused:   # rank 0 timed the fused kernel: take the sweep's share out, the model adds it back
            density_ms = max(0.1 * density_ms, density_ms - HEAD_FUSED_SWEEP_FACTOR * float(sweep[self.z0:self.z1].sum()))
        mine = (self.z0, self.z1, density_ms, march_ms, wd[self.z0:self.z1].tolist(), wm[self.z0:self.z1].tolist(), self.head_fused)
        rows = [None] * self.world
        d.all_gather_object(rows, mine)
        fused0 = bool(rows[0][6])
        rows = [r[:6] for r in rows]
        dc, mc = np.zeros(nz), np.zeros(nz)
        for (a, b, dm, mm, w1, w2) in rows:
            w1, w2 = np.asarray(w1), np.asarray(w2)
            dc[a:b] = dm * w1 / w1.sum()
            mc[a:b] = mm * w2 / w2.sum()
        slabs_, _ = balance_slabs(dc, sweep, mc, self.world, head_sweep_factor=HEAD_FUSED_SWEEP_FACTOR if fused0 else 1.0)
        self.set_slab(*slabs_[self.rank])
        self._times = None
        return slabs_

    def check_links(self):
        """Raises when a bounded wait of the sheet link or the image link gave up (a peer that never signalled): the
        frame just rendered is then not the reference's image.  Synchronises this rank's stream."""
        if self.world > 1 and hasattr(self.e, "link_timeouts"):
            n = self.e.link_timeouts()
            if n:
                raise RuntimeError("rank %d: %d sheet/image link waits gave up; the frame is not valid" % (self.rank, n))

    # -- march ------------------------------------------------------------------------------------
    def march(self, camera, gather=True, count_samples=True, host_image=None):
        """Returns (rgba, total_ray_samples): rgba is the full H x W x 4 image on rank 0 when
        `gather` (None elsewhere), else this rank's band. count_samples=False skips the read-back
        and all-reduce of the sample counter (a host synchronisation) and returns None for it.
        host_image (a SharedHostImage): every rank copies its band into the shared host image instead of
        gathering on rank 0's device; the copy is asynchronous, the caller synchronises and barriers."""
        e, d = self.e, self.dist
        h, w = int(camera["height"]), int(camera["width"])
        per = -(-h // self.world)                    # image rows per owner; the image is padded to R * per rows
        if self.world > 1 and self._image_link(w, h):
            # the march kernel stores every pixel's two partials straight into the receive buffer of the rank that
            # composites its row (peer memory) and raises a flag; the composite kernel waits for all ranks' flags
            e.march_linked(camera)
            samples = e.last_ray_samples() if count_samples else None
            band = e.composite_linked(per, w)
        else:
            over, under = e.march_partial(camera, per * self.world)   # 2 x (R*per, W, 4), premultiplied; rows >= H stay 0
            samples = e.last_ray_samples() if count_samples else None
            if self.world == 1:
                out = e.composite([over, under], h * w).reshape(h, w, 4)
                return out, samples
            # rows [q*per, (q+1)*per) of both partials go to rank q: the buffers are already laid out by owner
            recv_over = e.buffer("recv_over", (self.world, per, w, 4))    # [slab][row][col][rgba]
            recv_under = e.buffer("recv_under", (self.world, per, w, 4))
            d.all_to_all_single(recv_over.view(-1), over.view(-1))
            d.all_to_all_single(recv_under.view(-1), under.view(-1))
            parts = []
            for s in range(self.world):                  # ascending slab order = ascending z
                parts += [recv_over[s], recv_under[s]]
            band = e.composite(parts, per * w).reshape(per, w, 4)
        total = e.all_reduce_sum(d, samples) if count_samples else None
        if count_samples:
            self.check_links()                       # the host has synchronised for the counter anyway
        if host_image is not None:
            e.copy_band_to_host(band, host_image.band())
            return None, total
        if not gather:
            r0, r1 = image_band(h, self.world, self.rank)
            return band[:r1 - r0], total
        full = e.buffer("full", (self.world, per, w, 4)) if self.rank == 0 else None
        d.gather(band, gather_list=list(full.unbind(0)) if self.rank == 0 else None, dst=0)
        if self.rank != 0:
            return None, total
        return full.reshape(self.world * per, w, 4)[:h], total


class SharedHostImage:
    """The frame's H x W x 4 float image in POSIX shared memory, mapped by every rank of the node (and page-locked
    for CUDA when `pin` is given): each rank copies the band of rows it composited straight home over its own PCIe
    link, rank 0 reads the whole image after a barrier. Replaces gather-to-rank-0 followed by one big device-to-host
    copy: R copies of 1/R of the image run in parallel."""

    def __init__(self, dist, height, width, pin=None):
        from multiprocessing import shared_memory
        self.dist = dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.h, self.w = int(height), int(width)
        nbytes = self.h * self.w * 16
        name = [None]
        if self.rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
            name[0] = self.shm.name
        if dist is not None and self.world > 1:
            dist.broadcast_object_list(name, src=0)
        if self.rank != 0:
            self.shm = shared_memory.SharedMemory(name=name[0])
        self.array = np.ndarray((self.h, self.w, 4), dtype=np.float32, buffer=self.shm.buf)
        self.r0, self.r1 = image_band(self.h, self.world, self.rank)
        self._pin = pin
        self._registered = False
        if pin is not None:                       # pin(address, nbytes) -> True when the pages are locked for the GPU
            self._registered = bool(pin(self.array.ctypes.data, nbytes))

    def band(self):
        """This rank's rows of the shared image (a view)."""
        return self.array[self.r0:self.r1]

    def close(self, unpin=None):
        if self._registered and unpin is not None:
            unpin(self.array.ctypes.data)
        self.array = None
        try:
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# CUDA adapter (the product path)
# ------------------------------------------------------------------------------------------------
class _DevPtr:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class CudaSlabEngine:
    """Slab engine over libvpe_cuda.so with torch CUDA tensors as buffers (torch is plumbing only:
    device memory, the current stream, and NCCL through torch.distributed)."""

    def __init__(self, scene, rank, world, device):
        import torch
        from . import engine as _engine
        from . import scenes
        self.torch = torch
        self.device = torch.device("cuda", device)
        gz = scene["grid"][2]
        self.slab = slab_range(gz, world, rank)
        self.eng = _engine.engine_for_scene(None, scene, device=device, slab=self.slab)
        self.eng.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        scenes.apply_scene(self.eng, scene)
        self.grid, self.N = self.eng.grid, self.eng.N
        self._sheet = None
        self._particles = None
        self._buffers = {}

    def new_tensor(self, shape):
        return self.torch.empty(tuple(shape), dtype=self.torch.float32, device=self.device)

    def fill_prepare(self, particles, emitter):
        t = self.torch
        if not isinstance(particles, t.Tensor):
            particles = t.from_numpy(np.ascontiguousarray(particles, dtype=np.float32))
        self._particles = particles.to(self.device, non_blocking=True).contiguous()
        self.eng.fill_prepare_device(self._particles.data_ptr(), self._particles.shape[0], emitter)

    def fill_region(self, x0, x1, y0, y1):
        self.eng.fill_region(x0, x1, y0, y1)

    def fill_density(self):
        self.eng.fill_density()

    def fill_sweep_region(self, x0, x1, y0, y1):
        self.eng.fill_sweep_region(x0, x1, y0, y1)

    def link_neighbours(self, dist, rank, world):
        """Exchange the CUDA IPC handles of the sheet link buffers and map the neighbours' buffers.
        Returns False (the caller falls back to NCCL send/recv) when the ranks are not on one node."""
        import socket
        handle, _ = self.eng.sheet_link_create()
        mine = (socket.gethostname(), handle)
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        if len({h for h, _ in everyone}) != 1:
            return False
        up = everyone[rank - 1][1] if rank > 0 else None
        down = everyone[rank + 1][1] if rank < world - 1 else None
        ok = True
        try:
            self.eng.sheet_link_connect(up, down)   # cudaIpcOpenMemHandle + peer access
        except Exception:
            ok = False
        # the link is used only if every rank could map its neighbours (all ranks must take the same path)
        results = [None] * world
        dist.all_gather_object(results, ok)
        if not all(results):
            try:
                self.eng.sheet_link_connect(None, None)
            except Exception:
                pass
            return False
        return True

    def fill_sweep_linked(self):
        self.eng.fill_sweep_linked()

    def fill_linked(self):
        self.eng.fill_linked()

    def image_link_neighbours(self, dist, rank, world, w, h):
        """Exchange the IPC handles of the image receive buffers and map every rank's buffer (same node only)."""
        handle, _ = self.eng.image_link_create(world, rank, w, h)
        everyone = [None] * world
        dist.all_gather_object(everyone, handle)
        ok = True
        try:
            self.eng.image_link_connect([None if q == rank else everyone[q] for q in range(world)])
        except Exception:
            ok = False
        results = [None] * world
        dist.all_gather_object(results, ok)
        return all(results)

    def march_linked(self, camera):
        self.eng.march_linked(camera)

    def composite_linked(self, per, w):
        out = self.buffer("band", (per, w, 4))
        self.eng.composite_linked(out.data_ptr())
        return out

    def set_slab(self, z0, z1):
        self.eng.set_config(slabZBegin=int(z0), slabZEnd=int(z1))
        self.slab = (int(z0), int(z1))

    def record_event(self):
        ev = self.torch.cuda.Event(enable_timing=True)
        ev.record(self.torch.cuda.current_stream(self.device))
        return ev

    def elapsed_ms(self, a, b):
        b.synchronize()
        return float(a.elapsed_time(b))

    def sheet_tensor(self):
        if self._sheet is None:
            gx, gy, _ = self.grid
            ptr = self.eng.light_sheet_device_ptr()
            self._sheet = self.torch.as_tensor(_DevPtr(ptr, (gy * self.N, gx * self.N)), device=self.device)
        return self._sheet

    def sheet_written(self, y0, y1):  # device memory is shared with the context: nothing to copy
        pass

    def sheet_read(self, y0, y1):
        pass

    def buffer(self, name, shape):
        """A cached, zero-initialised device tensor (reused across frames)."""
        key = (name, tuple(shape))
        if key not in self._buffers:
            self._buffers[key] = self.torch.zeros(tuple(shape), dtype=self.torch.float32, device=self.device)
        return self._buffers[key]

    def march_partial(self, camera, padded_rows=None):
        h, w = int(camera["height"]), int(camera["width"])
        rows = max(h, padded_rows or h)
        over, under = self.buffer("over", (rows, w, 4)), self.buffer("under", (rows, w, 4))
        self.eng.march_partial_device(camera, over.data_ptr(), under.data_ptr())
        return over, under

    def last_ray_samples(self):
        return int(self.eng.stats()["raySamples"])

    def composite(self, parts, num_pixels):
        out = self.buffer("composite", (num_pixels, 4))
        self._keep = [p.contiguous() for p in parts]
        self.eng.composite_device([p.data_ptr() for p in self._keep], num_pixels, out.data_ptr())
        return out

    def all_reduce_sum(self, dist, value):
        t = self.torch.tensor([value], dtype=self.torch.int64, device=self.device)
        dist.all_reduce(t)
        return int(t.item())

    def stats(self):
        return self.eng.stats()

    def profile_slices(self, on):
        """Keep per-slice work figures of the fills and marches that follow (SlabRenderer.rebalance reads them).
        self.debug holds the other VpeDebugOptions of this engine (set_debug_options replaces all of them)."""
        self.eng.set_debug_options(profile_slices=bool(on), **getattr(self, "debug", {}))
        self._profiling = bool(on)

    def slice_profile(self):
        if not getattr(self, "_profiling", False):
            return None
        return self.eng.read_slice_profile()

    def link_timeouts(self):
        """Waits of the sheet link and the image link that gave up (a peer that never arrived): must be 0."""
        return int(self.eng.sheet_link_timeouts()) + int(self.eng.image_link_timeouts())

    def copy_band_to_host(self, band, host_rows):
        """Asynchronous device-to-host copy of this rank's composited rows into (pinned, shared) host memory."""
        t = self.torch
        dst = t.from_numpy(host_rows)
        dst.copy_(band[:host_rows.shape[0]], non_blocking=True)

    @staticmethod
    def pin_host(address, nbytes):
        import torch
        return int(torch.cuda.cudart().cudaHostRegister(address, nbytes, 0)) == 0

    @staticmethod
    def unpin_host(address):
        import torch
        torch.cuda.cudart().cudaHostUnregister(address)
