#!/usr/bin/env python
"""bench.py — Fill Volume + Ray March throughput on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # CUDA engine (the product)
  python bench.py --impl reference --steps K --warmup W    # CPU oracle on the host cores

A step is one frame of the hot path: vpe_fill (bin + fill every covered metavoxel) followed by
vpe_march (full image), on synthetic displaced-sphere particles.  The workload is
BASELINE.json's headline configuration "32^3 grid x 32^3 voxels, 8k particles, 1080p" (cfg3) at every
N (strong scaling: N > 1 cuts the grid into light-axis slabs).  Reported, per BASELINE.json's metric
"ray-samples/sec (march) + voxels/sec (fill)":
  value        march ray-samples/s, device-timed, inputs resident in HBM
  fill.value   fill voxels/s, device-timed
  e2e          the WHOLE FRAME (fill + march) through the host-buffer API: pinned host particles in (H2D) and the
               float4 image out in host memory (D2H) inside the timed region; value = ray-samples / frame time
  roofline     march kernel: compulsory bytes (8 B x distinct texels touched + 16 B x pixels) / kernel time
  cpu_baseline the CPU oracle timed on this box's host cores on a bounded sample of the workload
  parity       the image of the timed configuration against the CPU oracle on random pixels (N > 1: also
               against the single-GPU image, and the link time-out counters, which must be 0)
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
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=self.out, stderr=subprocess.DEVNULL)
            time.sleep(0.25)  # nvidia-smi needs a moment before its first sample
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
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


def workload_name(cfg_name, sc):
    return "%s: %d^3 grid x %d^3 voxels, %d particles, %dx%d, %d steps/metavoxel" % (
        cfg_name, sc["grid"][0], sc["numVoxels"], sc["particles"].shape[0], sc["camera"]["width"], sc["camera"]["height"],
        sc["rayMarchSteps"])


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (a C++ restatement; the reference itself is HLSL + C#/Unity and cannot run here).
# The only places of this file that touch oracle/: cpu_sample (the reported CPU baseline / the reference
# arm) and oracle_pixels (the checker of the image the timed configuration produced).
# ------------------------------------------------------------------------------------------------
def host_cores():
    """Cores this process may use: the affinity mask, capped by a cgroup CPU quota if there is one."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period))))
    except Exception:
        pass
    return n


def oracle_threads(lib):
    """torch.distributed.run exports OMP_NUM_THREADS=1 and libgomp may already be initialised: the oracle's
    thread count is set explicitly to the number of host cores."""
    want = host_cores()
    got = int(lib.vpe_ref_set_num_threads(want))
    assert got == want and (got > 1 or want == 1), "the CPU baseline must use all %d host cores, got %d" % (want, got)
    return got


def cpu_sample(cfg_name, steps, warmup, tile=128):
    """Time the oracle on a bounded sample of the workload: a tile x tile pixel block at the image
    centre, and the metavoxel columns those rays enter (all z)."""
    import ctypes as C
    from vpe_b200 import scenes
    from oracle_lib import load_oracle, oracle_engine
    lib = load_oracle()
    cores = oracle_threads(lib)
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
        "cores": cores, "fill_ms": fill_t * 1e3, "march_ms": march_t * 1e3, "samples": samples, "voxels": vox,
        "sample": "%s: march = %dx%d centre pixel tile (%d rays, %d ray-samples); fill = the %dx%d metavoxel columns those "
                  "rays enter, all %d slices (%d covered metavoxels, %d voxels); binning of all particles amortised per voxel; "
                  "OpenMP over %d threads (set explicitly)"
                  % (cfg_name, tile, tile, len(pix), samples, x1 - x0, y1 - y0, gz, covered_in_region, vox, cores),
    }


def cpu_baseline_entry(r):
    # whole-frame rate of the sample: its ray-samples over its fill + march time (same definition as e2e.value)
    frame = r["samples"] / ((r["fill_ms"] + r["march_ms"]) * 1e-3)
    return {"value": r["march_samples_per_s"], "unit": "ray-samples/s", "cores": r["cores"], "kind": "port",
            "sample": r["sample"], "fill_voxels_per_s": r["fill_voxels_per_s"], "frame_ray_samples_per_s": frame}


def oracle_pixels(cfg_name, image, n_pixels=10000, seed=7):
    """Checker: the CPU oracle's RGBA on `n_pixels` random pixels of the frame (full CPU fill, OpenMP) against
    `image` (H, W, 4). Returns the parity figures, or a reason why the configuration is too large."""
    from vpe_b200 import scenes
    from oracle_lib import load_oracle, oracle_engine
    from parity import RTOL, rel_err
    sc = scenes.make_scene(cfg_name)
    voxels = sc["grid"][0] * sc["grid"][1] * sc["grid"][2] * sc["numVoxels"] ** 3
    if voxels > 1.2e9:
        return {"checked": False, "why": "a full CPU fill of %.1e voxels does not fit the bench's time budget" % voxels}
    lib = load_oracle()
    oracle_threads(lib)
    t0 = time.perf_counter()
    ref = oracle_engine(sc)
    scenes.apply_scene(ref, sc)
    ref.fill(sc["particles"], sc["emitter"])
    cam = sc["camera"]
    W, H = cam["width"], cam["height"]
    pix = np.sort(np.random.default_rng(seed).choice(W * H, size=min(n_pixels, W * H), replace=False)).astype(np.int32)
    want, _ = ref.march_pixels(cam, pix)
    got = np.asarray(image, dtype=np.float32).reshape(-1, 4)[pix]
    err = rel_err(got, want)
    strict = np.abs(got - want) / np.maximum(np.abs(want), 1e-30)
    return {"checked": True, "pixels": int(len(pix)), "max_rel_err": float(err.max()), "tolerance": RTOL,
            "metric": "|gpu - oracle| / max(|oracle|, 1e-2) over RGBA (tests/parity.py)",
            "strict_relative_outlier_frac": float((strict > RTOL).mean()),
            "oracle": "CPU port, parity unpinned by the reference (DESIGN.md 2)", "seconds": time.perf_counter() - t0}


def cpu_legs(cfg_name, image, parity_pixels, timeout_s=420):
    """The CPU legs of the CUDA arm - the reported CPU baseline and the oracle check of the benchmarked image - in CHILD
    processes with a clean environment (no launcher-imposed OMP_NUM_THREADS, no rank variables) and a time limit, so
    that neither a launcher's environment nor a slow host can distort or hang the GPU benchmark."""
    env = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "RANK", "LOCAL_RANK", "WORLD_SIZE", "GROUP_RANK",
                                                            "ROLE_RANK", "LOCAL_WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "CUDA_VISIBLE_DEVICES")}
    env["CUDA_VISIBLE_DEVICES"] = ""
    me = os.path.abspath(__file__)
    cpu, parity = None, None
    try:
        out = subprocess.run([sys.executable, me, "--impl", "reference", "--config", cfg_name, "--steps", "1", "--warmup", "0"],
                             env=env, capture_output=True, text=True, timeout=timeout_s)
        cpu = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
    except Exception as ex:
        cpu = {"value": None, "unit": "ray-samples/s", "cores": host_cores(), "kind": "port", "sample": "failed: %r" % (ex,)}
    fd, path = tempfile.mkstemp(suffix=".npy")
    os.close(fd)
    try:
        np.save(path, np.asarray(image, dtype=np.float32))
        out = subprocess.run([sys.executable, me, "--check-image", path, "--config", cfg_name, "--parity-pixels", str(parity_pixels)],
                             env=env, capture_output=True, text=True, timeout=timeout_s)
        parity = json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as ex:
        parity = {"checked": False, "why": "failed: %r" % (ex,)}
    finally:
        os.unlink(path)
    if parity.get("checked"):
        assert parity["max_rel_err"] <= parity["tolerance"], "the benchmarked frame differs from the oracle: %r" % parity
    return cpu, parity


def run_check_image(args):
    print(json.dumps(oracle_pixels(args.config or "cfg3", np.load(args.check_image), args.parity_pixels)), flush=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg_name = args.config or "cfg3"
    r = cpu_sample(cfg_name, max(1, args.steps), max(0, args.warmup))
    from vpe_b200 import scenes
    frame = r["samples"] / ((r["fill_ms"] + r["march_ms"]) * 1e-3)
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
        "cpu_baseline": cpu_baseline_entry(r),
        "e2e": {"value": frame, "unit": "ray-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "definition": "ray-samples of the sample / (its fill + march time): the whole frame, as in the CUDA arm's e2e",
                "march_value": r["march_samples_per_s"], "fill_value": r["fill_voxels_per_s"], "fill_unit": "voxels/s"},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# CUDA arm, one GPU
# ------------------------------------------------------------------------------------------------
def general_path_legs(sc, cam, parts_dev, n, device, steps=3):
    """The paths the headline does not take: a coloured ambient (half4 texels, four-channel
    filter) and the reference's real target (ARGB32, quantised per blend) with the scene depth test - the general kernel."""
    import torch
    import vpe_b200
    from vpe_b200 import scenes
    out = {}
    W, H = cam["width"], cam["height"]
    rgba = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")

    def timed(eng):
        fk, mk = [], []
        for i in range(2 + steps):
            eng.fill_device(parts_dev.data_ptr(), n, sc["emitter"])
            eng.march_device(cam, rgba.data_ptr())
            st = eng.stats()
            if i >= 2:
                fk.append(st["fillKernelMs"])
                mk.append(st["marchKernelMs"])
        return float(np.mean(fk)), float(np.mean(mk)), st

    sc2 = dict(sc)
    sc2["ambient"] = (0.3, 0.2, 0.1)
    eng = vpe_b200.engine_for_scene(None, sc2, device=device)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    scenes.apply_scene(eng, sc2)
    f, m, st = timed(eng)
    out["coloured_ambient"] = {"fill_kernel_ms": f, "march_kernel_ms": m, "ray_samples": st["raySamples"],
                               "note": "ambient (0.3,0.2,0.1): half4 (r,g,b,density) texels, 8 loads per sample"}
    eng.close()
    eng = vpe_b200.engine_for_scene(None, sc, device=device)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    scenes.apply_scene(eng, sc)
    depth = np.full((H, W), 0.9 * sc["grid"][0] * sc["mvScale"], dtype=np.float32)  # an opaque plane behind the grid centre
    eng.set_march_options(target_format=1, scene_depth=depth)
    f, m, st = timed(eng)
    out["argb32_target_scene_depth"] = {"fill_kernel_ms": f, "march_kernel_ms": m, "ray_samples": st["raySamples"],
                                        "note": "targetFormat=1 (UNORM8 after every metavoxel's blend, VPR.cs:228) + ZTest Less "
                                                "against a scene depth plane at 0.9 G: the general kernel (shader's unfused sequence)"}
    eng.close()
    return out


def run_cuda(args):
    import torch
    import vpe_b200
    from vpe_b200 import scenes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        return run_cuda_slabs(args)

    cfg_name = args.config or "cfg3"
    sc = scenes.make_scene(cfg_name)
    eng = vpe_b200.engine_for_scene(None, sc, device=local_rank, earlyOut=args.early_out)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    scenes.apply_scene(eng, sc)
    experiment = bool(args.march_kernel or args.no_skip or args.tile_log2w is not None)
    if experiment:
        eng.set_debug_options(march_kernel=args.march_kernel, no_skip=args.no_skip, march_tile_log2w=args.tile_log2w)
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
    total_ms = t_start.elapsed_time(t_end)
    fill_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    march_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    st = eng.stats()
    samples, voxels = st["raySamples"], st["voxelsFilled"]
    launches = (st["fillLaunches"] + st["marchLaunches"]) * K

    # end to end through the host-buffer C-ABI (pinned host memory): H2D particles, D2H image; the whole frame
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
    clk = clocks.stop()
    e2e_fill_s, e2e_march_s = float(np.mean(e2e_fill)), float(np.mean(e2e_march))
    e2e_frame_s = e2e_fill_s + e2e_march_s
    image = rgba_host_np.copy()

    # roofline of the dominant kernel (march): compulsory read set / kernel time
    peak, peak_src = measured_peak_hbm()
    uniq, skipped_frac = None, None
    if args.early_out == 0.0:
        uniq = eng.march_footprint(cam)
        sf = eng.stats()
        skipped_frac = sf["raySamplesSkipped"] / max(sf["raySamples"], 1)
    mk = float(np.mean(kern_march))
    fk = float(np.mean(kern_fill))
    N = eng.N
    fill_bytes = voxels * (8.0 + 8.0 / N)
    roof = None
    if uniq is not None:
        march_bytes = 8.0 * uniq + 16.0 * W * H
        ach = march_bytes / (mk * 1e-3) / 1e9
        roof = {"kernel": "k_march_flat" if args.march_kernel == 0 else "k_march", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": recorded_traffic("k_march"), "peak_source": peak_src, "algorithmic_bytes": march_bytes,
                "kernel_ms": mk, "bytes_per_ray_sample": march_bytes / max(samples, 1), "distinct_texels": uniq,
                "note": "compulsory read set = 8 B x distinct texels in the union of all samples' trilinear footprints + 16 B x pixels "
                "(SURVEY 8d); the kernel is bound by instruction issue, load latency and L1 wavefronts, not HBM (DESIGN.md 5.3); traffic = "
                "dram bytes of the committed ncu capture (below the compulsory set: samples in empty space are not fetched)"}
    fach = fill_bytes / (fk * 1e-3) / 1e9
    roof_fill = {"kernel": "k_fill_columns", "bound": "hbm", "achieved": fach, "peak": peak, "unit": "GB/s", "frac": fach / peak,
                 "traffic": recorded_traffic("k_fill_columns"), "algorithmic_bytes": fill_bytes, "kernel_ms": fk,
                 "bytes_per_voxel": 8.0 + 8.0 / N}

    legs = None
    if not args.no_general_paths and not experiment:
        eng.close()
        legs = general_path_legs(sc, cam, parts_dev, n, local_rank)
    cpu, parity = None, None
    if not args.no_cpu_baseline:
        cpu, parity = cpu_legs(cfg_name, image, args.parity_pixels)

    line = {
        "metric": METRIC, "value": samples / (march_ms * 1e-3), "unit": "ray-samples/s", "n_gpus": 1, "steps": K,
        "warmup": max(3, args.warmup), "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg_name, sc),
            "cache": "inputs larger than L2 (brick pool %.2f GB vs 126 MB L2); no flush between iterations" % (st["brickPoolBytes"] / 1e9),
            "early_out_transmittance": args.early_out, "covered_metavoxels": st["numMetavoxelsCovered"],
            "particle_metavoxel_pairs": st["numParticlePairs"]},
        "fill": {"value": voxels / (fill_ms * 1e-3), "unit": "voxels/s", "ms": fill_ms, "kernel_ms": fk, "voxels": voxels},
        "march": {"value": samples / (march_ms * 1e-3), "unit": "ray-samples/s", "ms": march_ms, "kernel_ms": mk, "ray_samples": samples,
                  "skipped_sample_frac": skipped_frac,
                  "skipped_note": "ray-samples count every iteration of March.shader:254-279 (the metric's unit); this fraction of them lies "
                                  "in empty space (all 8 texels of the footprint have density 0: the blend is the identity) and is not fetched"},
        "e2e": {"value": samples / e2e_frame_s, "unit": "ray-samples/s", "h2d_bytes_per_step": int(n * 28),
                "d2h_bytes_per_step": int(W * H * 16), "frame_ms": e2e_frame_s * 1e3,
                "definition": "whole frame through the host-buffer C-ABI: vpe_fill (pinned particles in) + vpe_march (float4 image out to "
                              "pinned host memory); value = ray-samples / (fill + march wall time)",
                "march_ms": e2e_march_s * 1e3, "march_value": samples / e2e_march_s,
                "fill_value": voxels / e2e_fill_s, "fill_unit": "voxels/s", "fill_ms": e2e_fill_s * 1e3},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roof, "roofline_fill": roof_fill,
        "general_paths": legs,
        "cpu_baseline": cpu,
        "parity": parity,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# CUDA arm, N > 1: strong scaling of the same workload over light-axis slabs, one process per GPU
# ------------------------------------------------------------------------------------------------
def run_cuda_slabs(args):
    import torch
    import torch.distributed as dist
    import vpe_b200
    from vpe_b200 import scenes
    from vpe_b200.slabs import CudaSlabEngine, SharedHostImage, SlabRenderer, slab_range

    rank, world = dist.get_rank(), dist.get_world_size()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    cfg_name = args.config or "cfg3"
    sc = scenes.make_scene(cfg_name)
    torch.cuda.set_device(dev)
    eng = CudaSlabEngine(sc, rank, world, local_rank)
    eng.debug = dict(sweep_overlap=bool(args.sweep_overlap))
    eng.profile_slices(False)   # pushes eng.debug to the library
    r = SlabRenderer(eng, dist, fill_bands=args.fill_bands if getattr(args, "fill_bands", 0) else None,
                     head_fused=not args.no_head_fused and not args.sweep_overlap)
    cam = sc["camera"]
    W, H = cam["width"], cam["height"]
    n = sc["particles"].shape[0]
    parts_host = torch.from_numpy(sc["particles"]).pin_memory()
    parts_dev = parts_host.to(dev)

    def sync():
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        return float(t.item())

    warm = max(3, args.warmup)
    slab_list = [slab_range(eng.grid[2], world, q) for q in range(world)]
    if not getattr(args, "no_rebalance", False):
        # untimed: measure a frame, move the slab boundaries (SlabRenderer.rebalance), twice
        eng.profile_slices(True)
        for _ in range(3):
            r.profile = True
            for _ in range(2):
                r.fill(parts_dev, sc["emitter"])
                r.march(cam, gather=False, count_samples=False)
            slab_list = r.rebalance()
            r.profile = False
        eng.profile_slices(False)
    for _ in range(warm):
        r.fill(parts_dev, sc["emitter"])
        r.march(cam, gather=False)
    K = max(1, args.steps)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    clocks = ClockSampler(local_rank)
    clocks.start()
    sync()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(K):
        ev[i][0].record()
        r.fill(parts_dev, sc["emitter"])
        ev[i][1].record()
        r.march(cam, gather=False, count_samples=False)
        ev[i][2].record()
    t1.record()
    sync()
    total_ms = max_over_ranks(t0.elapsed_time(t1))
    fill_ms = max_over_ranks(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    march_ms = max_over_ranks(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    _, total_samples = r.march(cam, gather=False)        # untimed: the ray-sample count of the frame
    st = eng.stats()
    mk = max_over_ranks(st["marchKernelMs"])
    voxels = sum_over_ranks(st["voxelsFilled"])
    covered = sum_over_ranks(st["numMetavoxelsCovered"])
    pairs = sum_over_ranks(st["numParticlePairs"])
    pool = sum_over_ranks(st["brickPoolBytes"])
    launches = sum_over_ranks((st["fillLaunches"] + st["marchLaunches"] + 1) * K)

    # end to end with HOST buffers: pinned particles in on every rank, every rank copies the band it composited into one
    # shared pinned host image over its own PCIe link; the frame ends when rank 0 can read the whole image
    host = SharedHostImage(dist, H, W, pin=CudaSlabEngine.pin_host)
    e2e = []
    for i in range(2 + K):
        sync()
        a = time.perf_counter()
        r.fill(parts_host, sc["emitter"])
        r.march(cam, count_samples=False, host_image=host)
        sync()
        if i >= 2:
            e2e.append(time.perf_counter() - a)
    clk = clocks.stop()
    e2e_s = max_over_ranks(float(np.mean(e2e)))
    image = host.array.copy() if rank == 0 else None

    # correctness of what was just timed: no link wait may have given up, and the image must be the reference's
    timeouts = int(sum_over_ranks(eng.link_timeouts()))
    assert timeouts == 0, "a sheet/image link wait gave up (%d): the frame is not valid" % timeouts
    peak, peak_src = measured_peak_hbm()
    uniq = sum_over_ranks(eng.eng.march_footprint(cam))
    sf = eng.stats()
    skipped = sum_over_ranks(sf["raySamplesSkipped"])
    march_bytes = 8.0 * uniq + 16.0 * W * H * world
    ach = march_bytes / (mk * 1e-3) / 1e9
    host_pinned = bool(host._registered)
    # rank 0 runs the single-GPU comparison and the CPU legs; the others wait on the rendezvous store (a blocking socket
    # wait: an NCCL barrier would keep a GPU and a host core spinning next to the CPU baseline)
    store = dist.distributed_c10d._get_default_store()
    if rank != 0:
        store.wait(["vpe_bench_rank0_done"])
        host.close(CudaSlabEngine.unpin_host)
        return
    N = eng.N
    # the same frame on one GPU (this rank's): the slab composite must agree with it to rounding
    single = None
    try:
        one = vpe_b200.engine_for_scene(None, sc, device=local_rank)
        one.set_stream(torch.cuda.current_stream().cuda_stream)
        scenes.apply_scene(one, sc)
        one.fill(sc["particles"], sc["emitter"])
        img1, _ = one.march(cam, want_samples=False)
        from parity import max_rel_err
        single = {"max_rel_err": max_rel_err(image, img1), "single_gpu_ray_samples": one.stats()["raySamples"]}
        assert single["single_gpu_ray_samples"] == total_samples, "the slabs' ray-sample counts do not add up to the single-GPU count"
        assert single["max_rel_err"] <= 1e-4, "the slab composite differs from the single-GPU image: %r" % single
        one.close()
    except vpe_b200.VpeError as ex:   # e.g. the whole volume does not fit one GPU next to this rank's slab
        single = {"skipped": str(ex)}
    cpu, parity = None, None
    try:
        if not args.no_cpu_baseline:
            cpu, parity = cpu_legs(cfg_name, image, args.parity_pixels)
    finally:
        store.set("vpe_bench_rank0_done", "1")
    host.close(CudaSlabEngine.unpin_host)
    fill_bytes = voxels * (8.0 + 8.0 / N)
    line = {
        "metric": METRIC, "value": total_samples / (march_ms * 1e-3), "unit": "ray-samples/s", "n_gpus": world, "steps": K,
        "warmup": warm, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg_name, sc),
            "parallelism": "light-axis slabs x%d (fill: %s; march: slab-local, %s, ordered compositing by screen band)" % (
                world, ("persistent TMA sweep kernel concurrent with the density pass, sheet handed to the next rank over NVLink peer memory"
                        if args.sweep_overlap else "TMA sweep kernel after the density pass, sheet handed to the next rank over NVLink peer memory"
                        + ("; the rank nearest the light runs the fused fill kernel and feeds the link itself" if r.head_fused else ""))
                if r.linked else "sheet rows over NCCL send/recv in %d bands" % len(r.bands),
                "non-zero partials stored into the compositing rank's memory by the march kernel (peer memory + flags)"
                if r._image_links.get((W, H)) else "NCCL all-to-all of the partial images"),
            "slabs": [list(x) for x in slab_list],
            "cache": "inputs larger than L2 (brick pools %.2f GB in total); no flush between iterations" % (pool / 1e9),
            "covered_metavoxels": int(covered), "particle_metavoxel_pairs": int(pairs)},
        "fill": {"value": voxels / (fill_ms * 1e-3), "unit": "voxels/s", "ms": fill_ms, "voxels": int(voxels)},
        "march": {"value": total_samples / (march_ms * 1e-3), "unit": "ray-samples/s", "ms": march_ms, "kernel_ms": mk,
                  "ray_samples": int(total_samples), "skipped_sample_frac": skipped / max(total_samples, 1)},
        "e2e": {"value": total_samples / e2e_s, "unit": "ray-samples/s", "h2d_bytes_per_step": int(n * 28) * world,
                "d2h_bytes_per_step": int(W * H * 16), "frame_ms": e2e_s * 1e3, "host_image_pinned": host_pinned,
                "definition": "whole frame through SlabRenderer: pinned host particles in on every rank (H2D), fill + march, every rank "
                              "copies its composited band into one shared pinned host image (D2H); value = ray-samples / frame wall time"},
        "gpu_launches": int(launches), "clocks": clk,
        "roofline": {"kernel": "k_march_flat", "bound": "hbm", "achieved": ach, "peak": peak * world, "unit": "GB/s",
                     "frac": ach / (peak * world), "traffic": recorded_traffic("k_march"), "peak_source": peak_src + " x n_gpus",
                     "algorithmic_bytes": march_bytes, "kernel_ms": mk, "distinct_texels": int(uniq),
                     "note": "all ranks' compulsory bytes / the slowest rank's kernel time; traffic = the single-GPU ncu capture"},
        "roofline_fill": {"kernel": "k_fill_columns<density> + k_sweep_tma", "bound": "hbm", "achieved": fill_bytes / (fill_ms * 1e-3) / 1e9,
                          "peak": peak * world, "unit": "GB/s", "frac": fill_bytes / (fill_ms * 1e-3) / 1e9 / (peak * world),
                          "traffic": recorded_traffic("k_fill_columns"), "algorithmic_bytes": fill_bytes, "ms": fill_ms,
                          "bytes_per_voxel": 8.0 + 8.0 / N, "traffic_sweep": recorded_traffic("k_sweep_tma"),
                          "note": "whole fill step (bin + density + sweep), slowest rank; algorithmic bytes are the fused single-GPU "
                                  "figure (8 B/voxel written once); the split really moves 8 (density, written) + 16 (sweep, in place) "
                                  "B/voxel; traffic = the single-GPU ncu capture of the fused kernel, traffic_sweep = of k_sweep_tma over the whole grid"},
        "cpu_baseline": cpu,
        "parity": {"link_timeouts": timeouts, "vs_single_gpu": single, "vs_oracle": parity},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", default=None, help="cfg1..cfg5 (default: cfg3)")
    ap.add_argument("--early-out", type=float, default=0.0, help="marchEarlyOutTransmittance (0 = exact reference semantics)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU legs (baseline sample and the oracle parity check)")
    ap.add_argument("--no-general-paths", action="store_true", help="skip the two extra timed legs of the non-headline kernels (N=1)")
    ap.add_argument("--parity-pixels", type=int, default=10000)
    ap.add_argument("--check-image", default=None, help="internal: check an H x W x 4 .npy image of --config against the CPU oracle, print JSON")
    ap.add_argument("--march-kernel", type=int, default=0, help="VpeDebugOptions.marchKernel (experiments): 1 = general kernel, 2 = round 1's per-fragment loop")
    ap.add_argument("--tile-log2w", type=int, default=None, help="experiments: warp pixel tile width = 2^n (default 3: 8x4)")
    ap.add_argument("--no-skip", action="store_true", help="experiments: sample every step (ignore the empty-space bitmap)")
    ap.add_argument("--no-rebalance", action="store_true", help="N>1: keep equal slabs (default: balance the slab boundaries during warm-up)")
    ap.add_argument("--sweep-overlap", action="store_true", help="N>1 experiments: linked sweep concurrently with the density pass instead of after it")
    ap.add_argument("--no-head-fused", action="store_true", help="N>1 experiments: the rank nearest the light splits its fill (density + linked sweep) like the others")
    ap.add_argument("--fill-bands", type=int, default=0, help="N>1: row bands of the NCCL fill pipeline (default: the peer-memory sheet link)")
    args = ap.parse_args()
    if args.check_image:
        run_check_image(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
