"""Per-stage timeline of one multi-GPU frame (run under torchrun on the GPU box): device events and
host clocks around every stage of SlabRenderer.fill / .march, printed per rank."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import vpe_b200
from vpe_b200 import scenes, slabs

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
sc = scenes.make_scene(sys.argv[1] if len(sys.argv) > 1 else "cfg3")
mode = sys.argv[2] if len(sys.argv) > 2 else "nooverlap"      # overlap | nooverlap | regsweep | split (= nooverlap, head rank splits its fill too)
eng = slabs.CudaSlabEngine(sc, rank, world, lr)
eng.debug = dict(sweep_overlap=(mode == "overlap"), no_tma_sweep=(mode == "regsweep"))
r = slabs.SlabRenderer(eng, dist, head_fused=(mode in ("nooverlap", "regsweep")))
cam = sc["camera"]
parts = torch.from_numpy(sc["particles"]).cuda()
eng.profile_slices(True)
for _ in range(3):   # the bench's warm-up: balance the slabs
    r.profile = True
    for _ in range(2):
        r.fill(parts, sc["emitter"]); r.march(cam, gather=False, count_samples=False)
    slab_list = r.rebalance()
    r.profile = False
eng.profile_slices(False)
for _ in range(3):
    r.fill(parts, sc["emitter"]); r.march(cam, gather=False, count_samples=False)
e, d = r.e, dist
h, w = int(cam["height"]), int(cam["width"])
per = -(-h // world)
for rep in range(2):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ev = []
    def mark(tag):
        x = torch.cuda.Event(enable_timing=True); x.record(); ev.append((tag, x, time.perf_counter()))
    h0 = time.perf_counter()
    mark("start")
    e.fill_prepare(parts, sc["emitter"]); mark("prepare")
    if r.head_fused:
        mark("density"); e.fill_linked(); mark("sweep")      # head rank: ONE fused kernel
    else:
        e.fill_density(); mark("density")
        e.fill_sweep_linked(); mark("sweep")
    e.march_linked(cam); mark("march")
    e.composite_linked(per, w); mark("composite")
    torch.cuda.synchronize()
    h1 = time.perf_counter()
    for rr in range(world):
        dist.barrier()
        if rr == rank and rep == 1:
            st = e.stats()
            print("%s rank %d slab %s host %.2f ms | marchK %.2f | " % (mode, rank, e.slab, (h1 - h0) * 1e3, st["marchKernelMs"]) +
                  " ".join("%s:%.2f" % (t, ev[0][1].elapsed_time(x)) for (t, x, ht) in ev[1:]), flush=True)
assert e.link_timeouts() == 0
dist.destroy_process_group()
