#!/usr/bin/env python
"""bench.py — Fill Volume + Ray March throughput on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # CUDA engine (the product)
  python bench.py --impl reference --steps K --warmup W    # CPU oracle on the host cores

A step is one frame of the hot path: vpe_fill (bin + fill every covered metavoxel) followed by
vpe_march (full image), on synthetic displaced-sphere particles.  At N=1 the
workload is BASELINE.json's headline configuration "32^3 grid x 32^3 voxels, 8k particles, 1080p"
(cfg3).  Reported, per BASELINE.json's metric "ray-samples/sec (march) + voxels/sec (fill)":
  value        march ray-samples/s, device-timed, inputs resident in HBM
  fill.value   fill voxels/s, device-timed
  e2e          the same two rates through the C-ABI with HOST buffers (pinned): H2D of the particle
               array and D2H of the float4 image inside the timed region
  roofline     march kernel: compulsory bytes (8 B x distinct texels touched + 16 B x pixels) / time
  cpu_baseline the CPU oracle timed on this box's host cores on a bounded sample of the workload
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "ray-samples/sec (march) + voxels/sec (fill), 32^3x32^3 @1080p, 1/2/4/8 GPU"


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def recorded_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
    (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.out, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.close()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (a C++ restatement; the reference itself is HLSL + C#/Unity and cannot run here)
# ------------------------------------------------------------------------------------------------
def cpu_sample(cfg_name, steps, warmup, tile=128):
    """Time the oracle on a bounded sample of the workload: a tile x tile pixel block at the image
    centre, and the metavoxel columns those rays enter (all z)."""
    import ctypes as C
    from vpe_b200 import scenes
    from oracle_lib import load_oracle, oracle_engine
    lib = load_oracle()
    sc = scenes.make_scene(cfg_name)
    eng = oracle_engine(sc)
    scenes.apply_scene(eng, sc)
    cam = sc["camera"]
    W, H = cam["width"], cam["height"]
    ys, xs = np.mgrid[H // 2 - tile // 2:H // 2 + tile // 2, W // 2 - tile // 2:W // 2 + tile // 2]
    pix = (ys * W + xs).astype(np.int32).ravel()
    from vpe_b200.engine import _camera
    ccam = _camera(cam)
    gx, gy, gz = eng.grid
    t0 = time.perf_counter()
    eng.fill_prepare(sc["particles"], sc["emitter"])
    bin_s = time.perf_counter() - t0
    touched = np.zeros(gx * gy * gz, dtype=np.uint8)
    rc = lib.vpe_ref_touched_metavoxels(eng._ctx, C.byref(ccam), pix.ctypes.data, len(pix), touched.ctypes.data)
    assert rc == 0
    t3 = touched.reshape(gz, gy, gx)
    cols = t3.any(axis=0)
    yy, xx = np.nonzero(cols)
    x0, x1, y0, y1 = int(xx.min()), int(xx.max()) + 1, int(yy.min()), int(yy.max()) + 1
    covered_in_region = 0
    for z in range(gz):
        for y in range(y0, y1):
            for x in range(x0, x1):
                covered_in_region += 1 if len(eng.read_particle_list(x, y, z)) else 0
    vox = covered_in_region * eng.N ** 3
    total_vox = eng.stats()["voxelsFilled"]
    fill_s, march_s, samples = [], [], 0
    for it in range(warmup + steps):
        eng.fill_prepare(sc["particles"], sc["emitter"])
        a = time.perf_counter()
        eng.fill_region(x0, x1, y0, y1)
        b = time.perf_counter()
        _, smp = eng.march_pixels(cam, pix)
        c = time.perf_counter()
        if it >= warmup:
            fill_s.append(b - a)
            march_s.append(c - b)
            samples = int(smp.sum())
    fill_t = float(np.mean(fill_s)) + bin_s * (vox / max(total_vox, 1))  # binning amortised per voxel
    march_t = float(np.mean(march_s))
    return {
        "march_samples_per_s": samples / march_t, "fill_voxels_per_s": vox / fill_t,
        "cores": int(lib.vpe_ref_num_threads()), "fill_ms": fill_t * 1e3, "march_ms": march_t * 1e3,
        "sample": "%s: march = %dx%d centre pixel tile (%d rays, %d ray-samples); fill = the %dx%d metavoxel columns those "
                  "rays enter, all %d slices (%d covered metavoxels, %d voxels); binning of all particles amortised per voxel"
                  % (cfg_name, tile, tile, len(pix), samples, x1 - x0, y1 - y0, gz, covered_in_region, vox),
    }


def workload_name(cfg_name, sc):
    return "%s: %d^3 grid x %d^3 voxels, %d particles, %dx%d, %d steps/metavoxel" % (
        cfg_name, sc["grid"][0], sc["numVoxels"], sc["particles"].shape[0], sc["camera"]["width"], sc["camera"]["height"],
        sc["rayMarchSteps"])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg_name = args.config or "cfg3"
    r = cpu_sample(cfg_name, max(1, args.steps), max(0, args.warmup))
    from vpe_b200 import scenes
    line = {
        "impl": "reference", "metric": METRIC, "value": r["march_samples_per_s"], "unit": "ray-samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["fill_ms"] + r["march_ms"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg_name, scenes.make_scene(cfg_name)),
                   "sample": "each step is a bounded sample of the workload, see cpu_baseline.sample",
                   "note": "CPU oracle (C++ restatement of the reference's shader + dispatch math, OpenMP over all host cores); "
                           "the reference itself is HLSL + C#/Unity and cannot run headless"},
        "fill": {"value": r["fill_voxels_per_s"], "unit": "voxels/s", "ms": r["fill_ms"]},
        "march": {"value": r["march_samples_per_s"], "unit": "ray-samples/s", "ms": r["march_ms"]},
        "cpu_baseline": {"value": r["march_samples_per_s"], "unit": "ray-samples/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"], "fill_voxels_per_s": r["fill_voxels_per_s"]},
        "e2e": {"value": r["march_samples_per_s"], "unit": "ray-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "fill_value": r["fill_voxels_per_s"], "fill_unit": "voxels/s"},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import vpe_b200
    from vpe_b200 import scenes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from vpe_b200 import slabs
        return slabs.bench_multi_gpu(args, METRIC, measured_peak_hbm, ClockSampler)

    cfg_name = args.config or "cfg3"
    sc = scenes.make_scene(cfg_name)
    eng = vpe_b200.engine_for_scene(None, sc, device=local_rank, earlyOut=args.early_out)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    scenes.apply_scene(eng, sc)
    if args.march_kernel or args.no_skip:
        eng.set_debug_options(march_kernel=args.march_kernel, no_skip=args.no_skip)
    cam = sc["camera"]
    W, H = cam["width"], cam["height"]
    n = sc["particles"].shape[0]
    parts_host = torch.from_numpy(sc["particles"]).pin_memory()
    parts_dev = parts_host.cuda()
    rgba_dev = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    rgba_host = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
    rgba_host_np = rgba_host.numpy()

    def step():
        eng.fill_device(parts_dev.data_ptr(), n, sc["emitter"])
        eng.march_device(cam, rgba_dev.data_ptr())

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    K = max(1, args.steps)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    kern_fill, kern_march = [], []
    clocks = ClockSampler(local_rank)
    clocks.start()
    torch.cuda.synchronize()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(K):
        ev[i][0].record()
        eng.fill_device(parts_dev.data_ptr(), n, sc["emitter"])
        ev[i][1].record()
        eng.march_device(cam, rgba_dev.data_ptr())
        ev[i][2].record()
        st = eng.stats()  # (syncs) kernel-only device times of this step
        kern_fill.append(st["fillKernelMs"])
        kern_march.append(st["marchKernelMs"])
    t_end.record()
    torch.cuda.synchronize()
    clk = clocks.stop()
    total_ms = t_start.elapsed_time(t_end)
    fill_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    march_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    st = eng.stats()
    samples, voxels = st["raySamples"], st["voxelsFilled"]
    launches = (st["fillLaunches"] + st["marchLaunches"]) * K

    # end to end through the host-buffer C-ABI (pinned host memory): H2D particles, D2H image
    e2e_fill, e2e_march = [], []
    parts_host_np = parts_host.numpy()
    for i in range(2 + K):
        a = time.perf_counter()
        eng.fill(parts_host_np, sc["emitter"])
        b = time.perf_counter()
        eng.march(cam, want_samples=False, out=rgba_host_np)
        c = time.perf_counter()
        if i >= 2:
            e2e_fill.append(b - a)
            e2e_march.append(c - b)
    e2e_fill_s, e2e_march_s = float(np.mean(e2e_fill)), float(np.mean(e2e_march))

    # roofline of the dominant kernel (march): compulsory read set / kernel time
    peak, peak_src = measured_peak_hbm()
    uniq = eng.march_footprint(cam) if args.early_out == 0.0 else None
    mk = float(np.mean(kern_march))
    fk = float(np.mean(kern_fill))
    N = eng.N
    fill_bytes = voxels * (8.0 + 8.0 / N)
    roof = None
    if uniq is not None:
        march_bytes = 8.0 * uniq + 16.0 * W * H
        ach = march_bytes / (mk * 1e-3) / 1e9
        roof = {"kernel": "k_march", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": recorded_traffic("k_march"), "peak_source": peak_src, "algorithmic_bytes": march_bytes,
                "kernel_ms": mk, "bytes_per_ray_sample": march_bytes / max(samples, 1),
                "distinct_texels": uniq, "note": "compulsory read set = 8 B x distinct texels in the union of all samples' "
                "trilinear footprints + 16 B x pixels (SURVEY 8d); the kernel is bound by L1 wavefronts / issue, not HBM (DESIGN.md 5.3); "
                "traffic = dram bytes of the committed ncu capture (below the compulsory set: zero-density cells are skipped)"}
    fach = fill_bytes / (fk * 1e-3) / 1e9
    roof_fill = {"kernel": "k_fill_columns", "bound": "hbm", "achieved": fach, "peak": peak, "unit": "GB/s", "frac": fach / peak,
                 "traffic": recorded_traffic("k_fill_columns"), "algorithmic_bytes": fill_bytes, "kernel_ms": fk,
                 "bytes_per_voxel": 8.0 + 8.0 / N}

    cpu = None
    if not args.no_cpu_baseline:
        r = cpu_sample(cfg_name, 1, 0)
        cpu = {"value": r["march_samples_per_s"], "unit": "ray-samples/s", "cores": r["cores"], "kind": "port",
               "sample": r["sample"], "fill_voxels_per_s": r["fill_voxels_per_s"]}

    line = {
        "metric": METRIC, "value": samples / (march_ms * 1e-3), "unit": "ray-samples/s", "n_gpus": 1, "steps": K,
        "warmup": max(3, args.warmup), "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg_name, sc),
            "cache": "inputs larger than L2 (brick pool %.2f GB vs 126 MB L2); no flush between iterations" % (st["brickPoolBytes"] / 1e9),
            "early_out_transmittance": args.early_out, "covered_metavoxels": st["numMetavoxelsCovered"],
            "particle_metavoxel_pairs": st["numParticlePairs"]},
        "fill": {"value": voxels / (fill_ms * 1e-3), "unit": "voxels/s", "ms": fill_ms, "kernel_ms": fk, "voxels": voxels},
        "march": {"value": samples / (march_ms * 1e-3), "unit": "ray-samples/s", "ms": march_ms, "kernel_ms": mk, "ray_samples": samples},
        "e2e": {"value": samples / e2e_march_s, "unit": "ray-samples/s", "h2d_bytes_per_step": int(n * 28),
                "d2h_bytes_per_step": int(W * H * 16), "march_ms": e2e_march_s * 1e3,
                "fill_value": voxels / e2e_fill_s, "fill_unit": "voxels/s", "fill_ms": e2e_fill_s * 1e3},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roof, "roofline_fill": roof_fill,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", default=None, help="cfg1..cfg5 (default: cfg3 at N=1)")
    ap.add_argument("--early-out", type=float, default=0.0, help="marchEarlyOutTransmittance (0 = exact reference semantics)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--march-kernel", type=int, default=0, help="VpeDebugOptions.marchKernel (experiments): 1 = general kernel, 2 = round 1's per-fragment loop")
    ap.add_argument("--no-skip", action="store_true", help="experiments: sample every step (ignore the empty-space bitmap)")
    ap.add_argument("--no-rebalance", action="store_true", help="N>1: keep equal slabs (default: balance the slab boundaries during warm-up)")
    ap.add_argument("--fill-bands", type=int, default=0, help="N>1: row bands of the fill pipeline (default: one per metavoxel row)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
