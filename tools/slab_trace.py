"""Per-band timeline of the multi-GPU fill pipeline (run under torchrun on the GPU box)."""
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
bands = int(sys.argv[1]) if len(sys.argv) > 1 else None
sc = scenes.make_scene(sys.argv[2] if len(sys.argv) > 2 else "cfg3")
eng = slabs.CudaSlabEngine(sc, rank, world, lr)
r = slabs.SlabRenderer(eng, dist, fill_bands=bands)
parts = torch.from_numpy(sc["particles"]).cuda()
for _ in range(3):
    r.fill(parts, sc["emitter"])
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
# instrumented copy of SlabRenderer.fill
e, d = r.e, dist
gx, gy, gz = e.grid; n = e.N
ev = []
def mark(tag):
    x = torch.cuda.Event(enable_timing=True); x.record(); ev.append((tag, x, time.perf_counter()))
h0 = time.perf_counter()
mark("start")
e.fill_prepare(parts, sc["emitter"])
mark("prepared")
sheet = e.sheet_tensor()
for bi, (y0, y1) in enumerate(r.bands):
    rows = sheet[y0 * n:y1 * n]
    if rank > 0:
        d.recv(rows, src=rank - 1)
        mark("recv%d" % bi)
    e.fill_region(0, gx, y0, y1)
    mark("fill%d" % bi)
    if rank < world - 1:
        d.send(rows, dst=rank + 1)
        mark("send%d" % bi)
torch.cuda.synchronize()
h1 = time.perf_counter()
for rr in range(world):
    dist.barrier()
    if rr == rank:
        print("rank %d host total %.2f ms" % (rank, (h1 - h0) * 1e3))
        print("  " + " ".join("%s:%.2f(h%.2f)" % (t, ev[0][1].elapsed_time(x), (ht - h0) * 1e3) for (t, x, ht) in ev[1:]), flush=True)
dist.destroy_process_group()
