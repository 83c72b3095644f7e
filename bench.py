#!/usr/bin/env python
"""bench.py — BASELINE.json metric on its headline configuration.

Workload (configs[1]): vszip.BoxBlur(hradius=13, hpasses=5, vradius=13, vpasses=5) on 1920x1080 YUV420P16
uniform-noise frames.  One "step" = one pass of that filter over a batch of FRAMES_PER_STEP frames that
are already resident in HBM (value) or that live in pinned host memory and go through the
getFrame-style C-ABI entry point with both PCIe copies inside the timed region (e2e).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA, through the C ABI)
  python bench.py --impl reference [...]                       the reference's CPU algorithm (oracle port)
  python bench.py --extras                                     also time the other BASELINE configs (stderr + profiles/)

Multi-GPU (torchrun, one rank per GPU): frames are independent, every rank processes its own frames
(frame n -> GPU n mod k), no collective on the data path; scaling is weak.  The timed region is
bracketed by barrier + synchronize, timed with CUDA events on the launching stream, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

W, H, FMT = 1920, 1080, "YUV420P16"
ARGS = dict(hradius=13, hpasses=5, vradius=13, vpasses=5)
FRAMES_PER_STEP = 1024
FRAME_BYTES = (W * H + 2 * (W // 2) * (H // 2)) * 2          # 6,220,800 B read per frame
ALGO_BYTES = 2 * FRAME_BYTES                                  # read once + written once
METRIC = "fps @1080p YUV420P16, vszip.BoxBlur(hradius=13,hpasses=5,vradius=13,vpasses=5), device-resident"
WORKLOAD = "configs[1]: BoxBlur 13/5/13/5 on 1920x1080 YUV420P16 uniform-noise frames"


def frames_of_rank(rank, world, per_rank):
    """Frame numbers rank `rank` of `world` processes: frame n runs on GPU n mod k."""
    return [rank + world * i for i in range(per_rank)]


def max_over_ranks(x, world, device=None):
    """Slowest rank's value (the job is as slow as its slowest GPU); works with nccl (cuda) and gloo (cpu)."""
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_arm(steps, warmup, frames_per_step=None):
    """The reference's algorithm on the host cores: one frame per thread, like VapourSynth's fmParallel
    workers (src/vapoursynth/boxblur.zig:211).  The Zig plugin cannot be built in this image (no zig, no
    VapourSynth), so this is the C++ restatement that reproduces the reference's goldens (kind "port")."""
    import numpy as np

    import oracle
    oracle.lib()
    cores = os.cpu_count() or 1
    n = frames_per_step or cores
    rng = np.random.default_rng(1234)
    frames = [[rng.integers(0, 65536, size=s, dtype=np.uint32).astype(np.uint16) for s in ((H, W), (H // 2, W // 2), (H // 2, W // 2))]
              for _ in range(min(n, 2 * cores))]

    def one(i):
        for p in frames[i % len(frames)]:
            oracle.boxblur_plane(p, ARGS["hradius"], ARGS["hpasses"], ARGS["vradius"], ARGS["vpasses"])

    with ThreadPoolExecutor(cores) as ex:
        for _ in range(warmup):
            list(ex.map(one, range(n)))
        t0 = time.perf_counter()
        for _ in range(steps):
            list(ex.map(one, range(n)))
        dt = time.perf_counter() - t0
    fps = n * steps / dt
    return fps, dt / steps * 1e3, cores, n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), max(0, min(args.warmup, 1))
    fps, ms, cores, n = cpu_arm(steps, warmup)
    sample = f"{n} frames per step ({cores} host threads, one frame per thread), {steps} timed steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": n, "note": "CPU restatement of vszip 19.0.0 (oracle/), not the Zig binary"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except FileNotFoundError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.split(",") for r in Path(self.f.name).read_text().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            # under load = samples in the upper half of what was seen (idle samples before/after the loop drop out)
            hi = sorted(sm)[len(sm) // 2:]
            out.update(sm_mhz=hi[len(hi) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import vapoursynth_zip_b200 as vz

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    vz.core.init([local])
    dev = torch.device("cuda", local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_ranks(x):
        return max_over_ranks(x, world, dev)

    n = FRAMES_PER_STEP
    src = vz.DeviceClip(FMT, W, H, n)
    dst = vz.DeviceClip(FMT, W, H, n)
    mine = frames_of_rank(rank, world, n)                    # frame numbers n with n mod world == rank
    src.fill_noise(seed=1234, first_frame_no=mine[0], frame_no_stride=world)
    flt = vz.BoxBlurFilter(src.info(), **ARGS)
    # launch on a torch-owned stream so torch's CUDA events bracket exactly the kernels (a NULL stream would
    # select the library's own stream, which torch events on the default stream do not see)
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step():
        flt.run_device(src, dst, 0, n, stream)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        barrier()
        return max_ranks(ms) / steps

    # ---- headline: device-resident
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = vz.core.kernel_launches
    ms_step = timed(step, args.steps, args.warmup)
    launches = (vz.core.kernel_launches - l0) * args.steps // (args.steps + args.warmup)
    clocks = sampler.stop() if sampler else None
    fps = world * n / (ms_step * 1e-3)

    # ---- per-kernel durations for the roofline (H passes only / V passes only, same batch)
    fh = vz.BoxBlurFilter(src.info(), hradius=13, hpasses=5, vradius=0, vpasses=0)
    fv = vz.BoxBlurFilter(src.info(), hradius=0, hpasses=0, vradius=13, vpasses=5)
    ms_h = timed(lambda: fh.run_device(src, dst, 0, n, stream), max(3, args.steps // 2), 2)
    ms_v = timed(lambda: fv.run_device(src, dst, 0, n, stream), max(3, args.steps // 2), 2)
    peak, peak_src = peaks()
    dom_name, dom_ms = ("blur_h_kernel<u16,P=5>", ms_h) if ms_h >= ms_v else ("blur_v_kernel<u16,P=5>", ms_v)
    achieved = ALGO_BYTES * n / (dom_ms * 1e-3) / 1e9
    traffic = None  # DRAM bytes per launch of that kernel, from the committed ncu --set full capture
    tfiles = sorted((ROOT / "profiles").glob("traffic_r*.json"))
    if tfiles:
        per_frame = json.loads(tfiles[-1].read_text())["dram_bytes_per_frame"].get(dom_name.split("<")[0])
        traffic = per_frame * n if per_frame else None
    path_gbs = ALGO_BYTES * n / (ms_step * 1e-3) / 1e9

    # ---- end to end: pinned host frames -> vszip_boxblur_get_frame (H2D + kernels + D2H per frame), E2E_IN_FLIGHT requests in flight
    # (8 saturate PCIe in both directions; more only add contention on the copy engines)
    ne = 64
    host_in = [torch.empty(FRAME_BYTES, dtype=torch.uint8).pin_memory() for _ in range(ne)]
    host_out = [torch.empty(FRAME_BYTES, dtype=torch.uint8).pin_memory() for _ in range(ne)]
    shapes = [(H, W), (H // 2, W // 2), (H // 2, W // 2)]

    def planes_of(t):
        a = t.numpy().view(np.uint16)
        out, off = [], 0
        for (h, w) in shapes:
            out.append(a[off:off + h * w].reshape(h, w)); off += h * w
        return out

    rng = np.random.default_rng(rank)
    for t in host_in:
        t.numpy()[:] = rng.integers(0, 256, size=FRAME_BYTES, dtype=np.uint8)
    clip = vz.core.clip_from_frames(FMT, [planes_of(t) for t in host_in])
    lib = vz.load_library()
    import ctypes as C
    frames_in = [vz._cframe(planes_of(t)) for t in host_in]
    frames_out = [vz._cframe(planes_of(t)) for t in host_out]

    def one(i):
        rc = lib.vszip_boxblur_get_frame(flt.handle, mine[i], C.byref(frames_in[i]), C.byref(frames_out[i]))
        if rc:
            raise RuntimeError(vz._last_error())

    E2E_IN_FLIGHT = 8
    pool = ThreadPoolExecutor(E2E_IN_FLIGHT)

    def e2e_step():
        list(pool.map(one, range(ne)))

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = max_ranks(time.perf_counter() - t0)
    barrier()
    e2e_fps = world * ne * e2e_steps / e2e_s
    pool.shutdown()

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload on the host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cfps, cms, cores, cn = cpu_arm(2, 1)
        cpu = {"value": cfps, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"{cn} frames per step, 2 timed steps, one frame per host thread ({cores} threads); C++ restatement of vszip 19.0.0"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": n, "parallelism": f"frame-parallel x{world} (frame n -> GPU n mod k), no collective",
                       "l2": f"each step reads {n * FRAME_BYTES / 1e6:.0f} MB and writes {n * FRAME_BYTES / 1e6:.0f} MB per GPU, far above the 126 MB L2 (no flush needed)"},
            "clocks": clocks,
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": ne * FRAME_BYTES, "d2h_bytes_per_step": ne * FRAME_BYTES,
                    "frames_per_step_per_gpu": ne, "in_flight": E2E_IN_FLIGHT, "api": "vszip_boxblur_get_frame on pinned host frames"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "algorithmic_bytes": ALGO_BYTES * n, "peak_source": peak_src,
                         "note": "algorithmic bytes = 12,441,600 B per frame (read once + write once) x frames per launch / launch duration"},
            "path": {"h_kernel_ms": ms_h, "v_kernel_ms": ms_v, "step_ms": ms_step, "algorithmic_GBps": path_gbs, "frac_of_peak": path_gbs / peak},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner there) must not add to it:
    # point fd 1 at stderr for the whole run and hand the real stdout to the two print(json.dumps(..)) calls only.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        real_stdout.flush()


if __name__ == "__main__":
    main()
