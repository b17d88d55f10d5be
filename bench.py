#!/usr/bin/env python
"""bench.py -- headline benchmark of the path-tracing hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): C1 of BASELINE.md -- Cornell Box of
examples/cornell_box/cornell_box_shortest.py, 1024 x 1024, 64 spp, max 8 bounces.  One "step" =
one pass of the hot path over that batch: refresh() + pathtrace(spp) (+ the NCCL tile reduce
when N > 1).  Metric: Msamples/s = pixels x spp / seconds.

N > 1 (launched by torchrun, one process per GPU): the image is sharded by 4-column bands
(rank = (i / 4) mod N) and the spp is scaled by N, so every GPU traces the same number of
samples as at N = 1 (weak scaling); the only collective is the NCCL sum of per-tile sample
sums at the end of the step (tonemap time).  torch.distributed is used for the barrier, the
max-over-ranks reduction of the timings and the NCCL-id broadcast only.

Legs
  value         device-resident: scene, camera and accumulation buffer already in HBM.
  e2e           the same step through the public Python API (raytracingpbr_b200.PathTracer) with
                HOST buffers: scene + camera structs copied host->device and the accumulation
                buffer copied device->host inside the timed region, every step.
  roofline      dominant kernel (k_pathtrace_pool) timed with CUDA events on its launch
                stream: algorithmic HBM bytes / duration against MEASURED_PEAKS.json, plus the
                counted-work FP32 figure that actually bounds this path.
  cpu_baseline  the CPU oracle (oracle/oracle.c; stand-in for "Taichi ti.cpu", which cannot be
                installed here) on a bounded sample of the same workload, all host threads.
  --impl reference   the same CPU oracle as the reference arm (kind "port").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Msamples/s (pixels x spp / s), Cornell Box 1024^2, 8 bounces"
UNIT = "Msamples/s"
W, H, SPP, BOUNCES = 1024, 1024, 64, 8
WORKLOAD = "C1: cornell_box_shortest scene, 1024x1024, 64 spp, max 8 bounces"
BAND = 4                                 # N > 1: rank = (column / BAND) mod N -- fine interleave balances the ranks
# bytes / flops per unit
BYTES_PER_SAMPLE = 16                    # pool kernel: one float4 (radiance, 1) per sample into the scratch buffer
BYTES_PER_PIXEL_PER_LAUNCH = 32          # simple kernel: vec4 f32 accumulator, 16 B read + 16 B write
# dram__bytes_read.sum + dram__bytes_write.sum of ONE k_pathtrace_pool_jit launch at the C1 configuration
# (ncu --set full, profiles/r01h_ncu_full_k_pathtrace_pool_jit_c1.csv: 102.34 MB read + 1182.56 MB written)
NCU_TRAFFIC_BYTES_C1 = 102.339072e6 + 1.182560e9
FLOP_PER_SCENE_EVAL = 8 * 41             # 8 boxes x 41 flop
FLOP_PER_NORMAL = 4 * 41
FLOP_PER_RAY_SHADE = 150


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """SM clock, power and throttle reasons DURING the timed region (B200_PROFILING.md): NVML polled every 10 ms from a
    thread (nvidia_ml_py); falls back to spawning nvidia-smi (~0.15 s per sample) when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index: int):
        self.index, self.rows, self._stop = index, [], threading.Event()   # rows: (sm_mhz, sm_max_mhz, watts, reason bitmask)
        self._nvml = self._handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml, self._handle = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self._nvml = None
        self.source = "nvml" if self._nvml else "nvidia-smi"
        self._t = threading.Thread(target=self._run, daemon=True)

    def _sample_nvml(self):
        n, h = self._nvml, self._handle
        self.rows.append((float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)), float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)),
                          n.nvmlDeviceGetPowerUsage(h) / 1000.0, int(n.nvmlDeviceGetCurrentClocksThrottleReasons(h))))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        parts = [s.strip() for s in out.strip().split(",")]
        if len(parts) >= 7:
            mask = sum(bit for k, (_, bit) in enumerate(self.BITS) if parts[3 + k].lower().startswith("active"))
            self.rows.append((float(parts[0]), float(parts[1]), float(parts[2]), mask))

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                if self._nvml:          # NVML call failed: switch to nvidia-smi for the rest of the run
                    self._nvml, self.source = None, "nvidia-smi"
            self._stop.wait(0.01 if self._nvml else 0.1)

    def start(self):
        self._t.start()

    def stop(self) -> dict:
        self._stop.set()
        self._t.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        sm = sorted(r[0] for r in self.rows)
        reasons = [name for name, bit in self.BITS if any(r[3] & bit for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.rows[0][1], "reasons": reasons,
                "power_w_max": max(r[2] for r in self.rows), "samples": len(self.rows), "source": self.source}


# ------------------------------------------------------------------------------ CPU legs
CPU_BANDS, CPU_COLS, CPU_SPP = 32, 2, 64      # bounded CPU sample: 64 spread columns x 1024 rows x 64 spp = 4.19 M samples (~15 s)


def cpu_sample_columns(width: int, bands: int = CPU_BANDS, cols: int = CPU_COLS):
    """A bounded, spread-out sample of the workload: `bands` groups of `cols` adjacent columns."""
    step = width // bands
    return [(b * step + step // 2, b * step + step // 2 + cols) for b in range(bands)]


def run_cpu_oracle(spp: int, hoisted: bool, threads: int = 0, bands: int = CPU_BANDS, cols: int = CPU_COLS):
    """Time the oracle on the sample columns of the C1 image.  Returns (Msamples/s, samples, s)."""
    from oracle import pyoracle as po   # checker / CPU baseline only (never the product path)
    threads = threads or cpu_threads()
    cfg = po.cornell_shortest_config(W, H, BOUNCES, seed=0)
    objs = po.objects_array(po.cornell_shortest_objects())
    img = np.zeros((W, H, 4), dtype=np.float32)
    columns = [i for (i0, i1) in cpu_sample_columns(W, bands, cols) for i in range(i0, i1)]
    samples = len(columns) * H * spp
    t0 = time.perf_counter()
    po.pathtrace_columns(cfg, objs, spp, columns, img, hoisted=hoisted, nthreads=threads)   # one OpenMP region, all threads
    dt = time.perf_counter() - t0
    return samples / dt / 1e6, samples, dt


def cpu_threads() -> int:
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU legs pass
    this count to the oracle explicitly)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def reference_arm(args, rank: int) -> int:
    """--impl reference: the CPU stand-in for the reference's Taichi ti.cpu path (as-written:
    Euler matrices recomputed per object per march step, cornell_box_shortest.py:43)."""
    if rank != 0:
        return 0
    cores = cpu_threads()
    spp = CPU_SPP
    for _ in range(args.warmup):
        run_cpu_oracle(1, hoisted=False, bands=4)
    vals, t_total, n_total = [], 0.0, 0
    for _ in range(args.steps):
        v, n, dt = run_cpu_oracle(spp, hoisted=False)
        vals.append(v)
        t_total += dt
        n_total += n
    value = n_total / t_total / 1e6
    sample = (f"{CPU_BANDS * CPU_COLS} of 1024 columns ({CPU_BANDS} spread bands x {CPU_COLS}) x 1024 rows x {spp} spp = "
              f"{n_total // args.steps} samples per step")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU oracle port of cornell_box_shortest.py (as-written rotation "
                   "recompute), stand-in for Taichi ti.cpu which is not installable here"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------ other configs (informational)
EXTRA_WORKLOADS = {
    # name: (preset, width, height, spp, max bounces, (exposure, gamma) of the synthetic environment, description)
    "c2": ("bunny_glass", 1024, 1024, 256, 16, (1.8, 2.2), "C2: bunny_sdf_glass scene, frame 0, 1024x1024, 256 spp, max 16 bounces"),
    "c3": ("tokyo_ibl", 1920, 1080, 128, 8, (1.8, 2.2), "C3: tokyo_ibl scene, 1920x1080, 128 spp, max 8 bounces"),
}


def synthetic_env_u8(w=3200, h=1600, seed=11):
    """Stand-in for the reference's 3200 x 1600 .hdr assets (not redistributable, absent on the GPU box):
    a smooth sky gradient with a bright sun lobe plus seeded noise, as uint8 (W, H, 3)."""
    rng = np.random.default_rng(seed)
    u = np.linspace(0, 1, w, dtype=np.float32)[:, None]
    v = np.linspace(0, 1, h, dtype=np.float32)[None, :]
    sky = 60 + 150 * v + 25 * np.sin(6.28 * u)
    sun = 255 * np.exp(-((u - 0.3) ** 2 + (v - 0.8) ** 2) * 400)
    img = np.clip(sky + sun, 0, 255)[..., None] * np.array([0.9, 0.95, 1.0], np.float32)
    img = img + rng.integers(0, 8, size=(w, h, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def extra_workload(args) -> int:
    """Single-GPU timing of another BASELINE.json config (no CPU leg, no counters); one JSON line."""
    from raytracingpbr_b200 import PathTracer, ibl, scenes
    preset, w, h, spp, bounces, env, desc = EXTRA_WORKLOADS[args.workload]
    cfg, objs, cam, tm = getattr(scenes, preset)(w, h, max_bounces=bounces, seed=0)
    with PathTracer(cfg, objs, cam, tm) as pt:
        pt.set_envmap(ibl.process(synthetic_env_u8(), *env))
        ctx = pt.ctx

        def step():
            ctx.flush_l2()
            ctx.refresh()
            ctx.set_sample_base(0)
            ctx.pathtrace(spp)
        for _ in range(args.warmup):
            step()
        ctx.sync()
        ctx.kernel_time()
        sampler = ClockSampler(0)
        sampler.start()
        ctx.timer_start()
        for _ in range(args.steps):
            step()
        ms = ctx.timer_stop()
        clocks = sampler.stop()
        kernel_ms, launches = ctx.kernel_time()
        active, jit = ctx.jit_status()
    line = {"metric": "Msamples/s (pixels x spp / s)", "value": w * h * spp / (ms / args.steps * 1e-3) / 1e6, "unit": UNIT,
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (procedural 3200x1600 environment)",
            "config": {"workload": desc, "jit": jit}, "kernel_ms_per_step": kernel_ms / args.steps, "clocks": clocks}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------ GPU arm
def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--kernel", default="persistent", choices=["persistent", "simple"])
    ap.add_argument("--counters", action="store_true", help="extra untimed pass with work counters (default at N=1)")
    ap.add_argument("--workload", default="c1", choices=list(EXTRA_WORKLOADS) + ["c1", "c4"],
                    help="c1 = the headline (default); c2 / c3 = BASELINE.json configs[2] / configs[3], informational; "
                         "c4 = configs[4]: 4096 x 4096 x 1024 spp tile-sharded over the ranks (strong scaling, run under torchrun)")
    args = ap.parse_args()
    if args.workload in EXTRA_WORKLOADS:
        return extra_workload(args)
    global W, H, SPP, WORKLOAD
    strong = args.workload == "c4"
    if strong:
        W, H, SPP = 4096, 4096, 1024
        WORKLOAD = "C4: cornell_box_shortest scene, 4096x4096, 1024 spp, max 8 bounces, tile-sharded"

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args, rank)

    from raytracingpbr_b200 import PathTracer, _native as N, scenes

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # weak scaling (c1): per-GPU samples fixed (W*H*SPP), tiles sharded by column band; c4: the job is fixed
    spp = SPP if strong else SPP * world
    kernel = N.KERNEL_PERSISTENT if args.kernel == "persistent" else N.KERNEL_SIMPLE
    cfg, objs, cam, tm = scenes.cornell_box_shortest(W, H, max_bounces=BOUNCES, seed=0, kernel=kernel)
    pt = PathTracer(cfg, objs, cam, tm, device=local_rank)
    ctx = pt.ctx
    if world > 1:
        import torch
        ctx.set_shard(rank, world, BAND)
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(N.Context.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        ctx.nccl_init(bytes(uid.cpu().numpy().tobytes()), rank, world)

    def step_device():
        ctx.flush_l2()
        ctx.refresh()
        ctx.set_sample_base(0)
        ctx.pathtrace(spp)
        if world > 1:
            ctx.reduce_tiles(0)

    # ---- leg 1: device-resident -----------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    ctx.sync()
    ctx.kernel_time()                      # reset per-launch event pool
    l0 = ctx.counters()["launches"]
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    ctx.sync()
    if sampler:
        sampler.start()
    ctx.timer_start()
    for _ in range(args.steps):
        step_device()
    ms = ctx.timer_stop()                  # synchronises the stream
    barrier()
    if sampler:
        clocks = sampler.stop()
    kernel_ms, kernel_launches = ctx.kernel_time()
    launches = ctx.counters()["launches"] - l0
    per_rank_kernel_ms = [kernel_ms / max(kernel_launches, 1)]
    if dist is not None:
        import torch
        t = torch.tensor([kernel_ms / max(kernel_launches, 1)], dtype=torch.float64, device="cuda")
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        per_rank_kernel_ms = [float(x.item()) for x in out]
    ms = max_over_ranks(ms)
    kernel_ms_max = max_over_ranks(kernel_ms)
    total_samples = float(W) * H * spp     # whole job, all ranks (each rank: W*H*SPP)
    ms_per_step = ms / args.steps
    value = total_samples / (ms_per_step * 1e-3) / 1e6

    # ---- leg 2: end to end through the public API with host buffers --------------------
    pinned = N.PinnedArray((W, H, 4), np.float32)          # page-locked destination of the per-step device -> host read
    host_img = pinned.array
    h2d = sum(len(bytes(o.to_native())) for o in objs) + len(bytes(cam.to_native()))
    d2h = host_img.nbytes if rank == 0 else 0

    def step_e2e():
        pt.set_scene(objs)                 # host structs -> device parameter block
        pt.set_camera(cam)
        ctx.flush_l2()
        pt.refresh()
        ctx.set_sample_base(0)
        pt.pathtrace(spp)
        if world > 1:
            ctx.reduce_tiles(0)
        if rank == 0:
            ctx.download(N.BUF_IMAGE_BUFFER, host_img)     # device -> host, synchronises
        else:
            ctx.sync()

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    ctx.sync()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = total_samples * args.steps / e2e_s / 1e6
    checksum = float(host_img[..., 3].sum()) if rank == 0 else 0.0
    host_img = None
    pinned.free()

    # ---- roofline of the dominant kernel ---------------------------------------------
    info = ctx.device_info()
    hbm_peak, peak_src, sm_max_mhz = peaks()
    k_ms = kernel_ms_max / max(kernel_launches, 1)
    local_pixels = W * H / world
    alg_bytes = (BYTES_PER_SAMPLE * local_pixels * spp if kernel == N.KERNEL_PERSISTENT
                 else BYTES_PER_PIXEL_PER_LAUNCH * local_pixels)
    launches_per_step = max(kernel_launches, 1) / args.steps      # > 1 when the spp are chunked to the scratch budget
    achieved = alg_bytes / (k_ms * launches_per_step * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "k_pathtrace_pool" if kernel == N.KERNEL_PERSISTENT else "k_pathtrace_simple",
            "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "traffic": NCU_TRAFFIC_BYTES_C1 if (kernel == N.KERNEL_PERSISTENT and world == 1 and not strong) else None,
            "algorithmic_bytes_per_launch": alg_bytes,
            "peak_source": peak_src, "kernel_ms_per_launch": k_ms, "kernel_ms_per_launch_by_rank": per_rank_kernel_ms, "kernel_share_of_step": kernel_ms_max / ms,
            "note": "this path is instruction-issue bound, not HBM bound (16 B written per sample); see fp32"}

    # counted work (untimed extra pass with the counting variant of the kernel)
    fp32 = None
    if rank == 0 and (args.counters or world == 1) and not strong:
        ccfg, _, _, _ = scenes.cornell_box_shortest(W, H, max_bounces=BOUNCES, seed=0, kernel=kernel, count_work=True)
        with PathTracer(ccfg, objs, cam, tm, device=local_rank) as cpt:
            cpt.refresh()
            cpt.pathtrace(4)
            cpt.sync()
            c = cpt.ctx.counters()
        per_sample = {k: c[k] / max(c["samples"], 1) for k in ("scene_evals", "rays", "normals")}
        flop_per_sample = (per_sample["scene_evals"] * FLOP_PER_SCENE_EVAL + per_sample["normals"] * FLOP_PER_NORMAL
                           + per_sample["rays"] * FLOP_PER_RAY_SHADE)
        samples_per_s_gpu = W * H * SPP / (k_ms * 1e-3)
        tflops = flop_per_sample * samples_per_s_gpu / 1e12
        clk = (clocks.get("sm_mhz") or sm_max_mhz) if sampler else sm_max_mhz
        peak_tflops = info["sm_count"] * 128 * 2 * sm_max_mhz * 1e6 / 1e12
        fp32 = {"achieved_tflops": tflops, "peak_tflops": peak_tflops, "frac": tflops / peak_tflops,
                "flop_per_sample": flop_per_sample, "per_sample": per_sample,
                "lane_utilisation_in_march_loop": (c["march_active"] / c["march_iters"]) if c["march_iters"] else None,
                "sm_count": info["sm_count"], "sm_mhz_under_load": clk,
                "peak_basis": "SMs x 128 lanes x 2 flop x clocks.max.sm"}

    # ---- CPU baseline (rank 0, N = 1 only) --------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and not strong:
        v_aw, n_aw, dt_aw = run_cpu_oracle(CPU_SPP, hoisted=False)
        v_h, n_h, dt_h = run_cpu_oracle(CPU_SPP, hoisted=True)
        cpu = {"value": v_aw, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
               "sample": f"{CPU_BANDS * CPU_COLS} of 1024 columns ({CPU_BANDS} spread bands x {CPU_COLS}) x 1024 rows x {CPU_SPP} spp = {n_aw} samples, {dt_aw:.1f} s; "
                         "as-written (rotation matrices recomputed per object per march step like cornell_box_shortest.py:43)",
               "hoisted_value": v_h, "hoisted_sample": f"{n_h} samples, {dt_h:.1f} s (matrices precomputed)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "width": W, "height": H, "spp_per_step": spp, "max_bounces": BOUNCES,
                       "sharding": (f"{world} ranks, {BAND}-column interleaved bands" + ("" if strong else f", spp x {world}")) if world > 1 else "none",
                       "l2": "flushed between steps by a 256 MiB memset on the launch stream (inside the timed region)",
                       "kernel": args.kernel, "blocks_per_sm": info["blocks_per_sm"], "jit": ctx.jit_status()[1]},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "alpha_checksum": checksum},
            "gpu_launches": launches,
            "clocks": clocks if sampler else None,
            "roofline": roof,
        }
        if fp32:
            line["fp32"] = fp32
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    pt.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
