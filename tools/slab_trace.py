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
eng = slabs.CudaSlabEngine(sc, rank, world, lr)
r = slabs.SlabRenderer(eng, dist)
cam = sc["camera"]
parts = torch.from_numpy(sc["particles"]).cuda()
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
    e.fill_density(); mark("density")
    e.fill_sweep_linked(); mark("sweep")
    over, under = e.march_partial(cam, per * world); mark("march_k")
    ro = e.buffer("recv_over", (world, per, w, 4)); ru = e.buffer("recv_under", (world, per, w, 4))
    d.all_to_all_single(ro.view(-1), over.view(-1)); mark("a2a_over")
    d.all_to_all_single(ru.view(-1), under.view(-1)); mark("a2a_under")
    ps = []
    for s in range(world):
        ps += [ro[s], ru[s]]
    e.composite(ps, per * w); mark("composite")
    torch.cuda.synchronize()
    h1 = time.perf_counter()
    for rr in range(world):
        dist.barrier()
        if rr == rank and rep == 1:
            st = e.stats()
            print("rank %d host %.2f ms | fillK %.2f marchK %.2f | " % (rank, (h1 - h0) * 1e3, st["fillKernelMs"], st["marchKernelMs"]) +
                  " ".join("%s:%.2f(h%.2f)" % (t, ev[0][1].elapsed_time(x), (ht - h0) * 1e3) for (t, x, ht) in ev[1:]), flush=True)
dist.destroy_process_group()
