#!/usr/bin/env python
"""bench.py -- headline benchmark of the path-tracing hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|c4]

Headline workload (config.workload): C1 of BASELINE.md -- Cornell Box of examples/cornell_box/cornell_box_shortest.py,
1024 x 1024, 64 spp, max 8 bounces.  One "step" = one pass of the hot path over that batch: refresh() + pathtrace(spp)
(+ the NCCL tile reduce when N > 1).  Metric: Msamples/s = pixels x spp / seconds.

N > 1 (launched by torchrun, one process per GPU): the image is sharded by 4-column bands (rank = (i / 4) mod N); the only
collective is the NCCL sum of per-tile sample sums at the end of the step (tonemap time).  torch.distributed is used for
the barrier, the max-over-ranks reduction of the timings and the NCCL-id broadcast only; the render goes through
raytracingpbr_b200.PathTracer (set_shard / nccl_init / reduce_tiles).

What one JSON line carries
  value / ms_per_step  device-resident WEAK-scaling figure: the spp is scaled by N, so every GPU traces as many samples as at N = 1.
  e2e           the same step through the public Python API with HOST buffers: scene + camera structs copied
                host->device and the accumulation buffer copied device->host inside the timed region, every step.
  strong        BASELINE's multi-GPU case at the C1 size: the FIXED 1024^2 x 64 spp job split over the N ranks -- ms per
                step, Msamples/s, per-rank kernel ms, the tile reduce alone, and image_crc32 of the reduced accumulation
                buffer, which must be the same number at N = 1, 2, 4, 8 (the image does not depend on the GPU count).
  c4            BASELINE configs[4]: 4096^2 x 1024 spp, 8 bounces, tile-sharded over the N ranks (strong); one timed step.
  workloads     (N = 1) configs[2] (bunny_sdf_glass 1024^2 x 256 spp x 16 bounces) and configs[3] (tokyo_ibl 1920 x 1080 x
                128 spp x 8 bounces): Msamples/s, kernel ms and the counted-work fraction of the FP32 peak.
  without_scene_analyses  (N = 1) the same step with the code generator's two exact shortcuts switched off, for transparency.
  roofline      the roof that bounds this path: FP32 instruction issue.  achieved = ALGORITHMIC flops (SURVEY.md 8(d):
                41 flop per box SDF etc. x the reference algorithm's evaluation counts, counted by the counting twin of
                the kernel) / the kernel's average launch duration (CUDA events on its launch stream).
  roofline_hbm  the HBM view north_star asks for: 8(d)'s algorithmic bytes (16 B per pixel per launch) and the design's
                own bytes (16 B per sample of ordered-accumulation scratch) against MEASURED_PEAKS.json.
  cpu_baseline  the CPU oracle (oracle/oracle.c; stand-in for "Taichi ti.cpu", which cannot be installed here) on a
                bounded sample of the same workload, all host threads.
  --impl reference   the same CPU oracle as the reference arm (kind "port"); the per-step sample shrinks for a large --steps so
                that all timed steps fit ~150 s of CPU time (config.subsample says what was used).

stdout carries exactly that one line: file descriptor 1 is pointed at stderr for the run (NCCL prints a version banner there).
"""
from __future__ import annotations

import argparse
import csv
import glob
import json
import os
import subprocess
import sys
import threading
import time
import zlib

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Msamples/s (pixels x spp / s), Cornell Box 1024^2, 8 bounces"
UNIT = "Msamples/s"
W, H, SPP, BOUNCES = 1024, 1024, 64, 8
WORKLOAD = "C1: cornell_box_shortest scene, 1024x1024, 64 spp, max 8 bounces"
BAND = 4                                 # N > 1: rank = (column / BAND) mod N -- fine interleave balances the ranks
# bytes per unit
ALG_BYTES_PER_PIXEL_PER_LAUNCH = 16      # 8(d): with the spp loop inside the kernel, one vec4 f32 accumulator per pixel per launch
DESIGN_BYTES_PER_SAMPLE = 16             # this design: one float4 (radiance, 1) per sample into the ordered-accumulation scratch
# flops per unit (translate 3 + rotate 15 + primitive; sqrt = 1)
FLOP_BOX, FLOP_SPHERE, FLOP_CYLINDER = 41, 28, 39
FLOP_BUNNY_OUTER = 3 + 15 + 15 + 1 + 6 + 2          # translate, rotate, animation matrix + bob, length, compare / subtract
FLOP_BUNNY_MLP = 1250 + 48 * 20                     # 8(d): ~1.25 kflop + 48 sin; a sine = 3-term reduction + two degree-7/8 polynomials ~ 20 flop
FLOP_PER_RAY_SHADE = 150


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def ncu_summary(tag: str):
    """Metrics of the tracked ncu capture profiles/r02*_ncu_full_*<tag>*.csv (tools/ncu_summary.py output), newest first.
    Returns ({metric: float}, file name) or ({}, None): the figures that cannot be measured inside a bench run (DRAM
    bytes, issue-active) are READ from the committed capture and the file is named next to them."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r02*_ncu_full_*{tag}*.csv")), reverse=True)
    for path in files:
        out = {}
        with open(path) as f:
            for row in csv.reader(f):
                if len(row) >= 3:
                    try:
                        v = float(row[2])
                    except ValueError:
                        continue
                    unit = row[1].strip().lower()
                    if row[0].startswith("dram__bytes"):          # ncu picks the unit per capture: normalise to MB
                        v *= {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(unit, 1.0)
                    out[row[0]] = v
        if out:
            return out, os.path.relpath(path, ROOT)
    return {}, None


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
REF_BUDGET_S = 150.0                          # --impl reference: CPU seconds for all timed steps together (fewer columns per step for a large K)


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


def reference_bands(rate_msamples: float, steps: int, spp: int = CPU_SPP) -> int:
    """Column bands of one reference-arm step: CPU_BANDS, halved until `steps` steps at the calibrated rate fit REF_BUDGET_S
    (never fewer than 4 bands = 8 columns)."""
    bands = CPU_BANDS
    if rate_msamples > 0.0:
        fit = rate_msamples * 1e6 * (REF_BUDGET_S / max(steps, 1)) / (H * spp * CPU_COLS)      # bands that fit one step's share
        while bands > 4 and bands > fit:
            bands //= 2
    return bands


def reference_arm(args, rank: int) -> int:
    """--impl reference: the CPU stand-in for the reference's Taichi ti.cpu path (as-written:
    Euler matrices recomputed per object per march step, cornell_box_shortest.py:43)."""
    if rank != 0:
        return 0
    cores = cpu_threads()
    spp = CPU_SPP
    rate = 0.0
    for _ in range(args.warmup):
        rate = max(rate, run_cpu_oracle(1, hoisted=False)[0])        # 64 columns x 1 spp: also calibrates the sample below
    # bounded sample per step: the whole --steps K run stays within ~REF_BUDGET_S of CPU time whatever K is
    bands = reference_bands(rate, args.steps, spp)
    t_total, n_total = 0.0, 0
    for _ in range(args.steps):
        v, n, dt = run_cpu_oracle(spp, hoisted=False, bands=bands)
        t_total += dt
        n_total += n
    value = n_total / t_total / 1e6
    v_h, n_h, dt_h = run_cpu_oracle(spp, hoisted=True, bands=bands)
    ncols = bands * CPU_COLS
    sample = (f"{ncols} of {W} columns ({bands} spread bands x {CPU_COLS}) x {H} rows x {spp} spp = "
              f"{n_total // args.steps} samples per step")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "width": W, "height": H, "spp_per_step": spp, "max_bounces": BOUNCES,
                   "subsample": {"columns": ncols, "of_columns": W, "rows": H, "spp": spp, "samples_per_step": n_total // args.steps,
                                 "why": "the CPU path needs ~4 minutes for the whole 67.1 M-sample step; pixels are independent and "
                                        "the columns are spread over the image, so Msamples/s of the subsample is that of the step"},
                   "note": "CPU oracle port of cornell_box_shortest.py run as written (rotation matrices recomputed per object per "
                           "march step, :43), all host threads: the stand-in for Taichi ti.cpu, which is not installable here"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "hoisted_value": v_h, "hoisted_note": "same port with the rotation matrices precomputed (the fair algorithmic "
                                                               f"baseline): {n_h} samples in {dt_h:.1f} s"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------ torch.distributed plumbing (N > 1 only)
class Ranks:
    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(self.local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist, self.torch = dist, torch

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def _reduce(self, x: float, op):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.MAX if self.dist else None)

    def gather(self, x: float):
        if self.dist is None:
            return [x]
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(v.item()) for v in out]

    def broadcast_bytes(self, payload: bytes | None, n: int) -> bytes:
        if self.dist is None:
            return payload
        t = self.torch.zeros(n, dtype=self.torch.uint8, device="cuda")
        if self.rank == 0:
            t = self.torch.frombuffer(bytearray(payload), dtype=self.torch.uint8).cuda()
        self.dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def sharded_tracer(R: Ranks, width: int, height: int, kernel: int):
    """PathTracer of this rank for the Cornell scene at (width, height): shard + communicator through the product API."""
    from raytracingpbr_b200 import PathTracer, nccl_unique_id, scenes
    cfg, objs, cam, tm = scenes.cornell_box_shortest(width, height, max_bounces=BOUNCES, seed=0, kernel=kernel)
    pt = PathTracer(cfg, objs, cam, tm, device=R.local)
    if R.world > 1:
        pt.set_shard(R.rank, R.world, BAND)
        uid = R.broadcast_bytes(nccl_unique_id() if R.rank == 0 else None, 128)
        pt.nccl_init(uid, R.rank, R.world)
    return pt, objs, cam, tm


def timed_steps(R: Ranks, pt, spp: int, warmup: int, steps: int, sampler=None):
    """`warmup` untimed + `steps` timed device-resident steps (L2 flushed, refresh, pathtrace, tile reduce): barrier +
    synchronize on both sides, CUDA events on the launch stream, max over ranks.  Returns a dict."""
    ctx = pt.ctx

    def step():
        ctx.flush_l2()
        pt.refresh()
        ctx.set_sample_base(0)
        pt.pathtrace(spp)
        if R.world > 1:
            pt.reduce_tiles(0)
    for _ in range(warmup):
        step()
    ctx.sync()
    ctx.kernel_time()                      # reset the per-launch event pool
    l0 = ctx.counters()["launches"]
    R.barrier()
    ctx.sync()
    if sampler:
        sampler.start()
    ctx.timer_start()
    for _ in range(steps):
        step()
    ms = ctx.timer_stop()                  # synchronises the stream
    R.barrier()
    clocks = sampler.stop() if sampler else None
    kernel_ms, kernel_launches = ctx.kernel_time()
    launches = ctx.counters()["launches"] - l0
    per_launch = kernel_ms / max(kernel_launches, 1)
    return {"ms_per_step": R.max(ms) / steps, "kernel_ms_per_step_max": R.max(kernel_ms) / steps, "kernel_launches": kernel_launches,
            "kernel_ms_per_launch_by_rank": R.gather(per_launch), "launches": launches, "clocks": clocks}


def reduce_alone_ms(R: Ranks, pt, spp: int, reps: int = 3) -> float | None:
    """The NCCL tile reduce by itself: every rank finishes its kernel first (sync + barrier), then CUDA events around
    rtpbr_reduce_tiles; max over ranks, mean over reps."""
    if R.world == 1:
        return None
    ctx, tot = pt.ctx, 0.0
    for _ in range(reps):
        pt.refresh()
        ctx.set_sample_base(0)
        pt.pathtrace(min(spp, 4))
        ctx.sync()
        R.barrier()
        ctx.timer_start()
        pt.reduce_tiles(0)
        tot += R.max(ctx.timer_stop())
    return tot / reps


def image_crc(R: Ranks, pt, spp: int):
    """crc32 of the reduced accumulation buffer (rank 0), after one fresh pass; and its alpha sum (= pixels x spp)."""
    from raytracingpbr_b200 import _native as N
    ctx = pt.ctx
    pt.refresh()
    ctx.set_sample_base(0)
    pt.pathtrace(spp)
    if R.world > 1:
        pt.reduce_tiles(0)
    crc, alpha = None, None
    if R.rank == 0:
        img = ctx.download(N.BUF_IMAGE_BUFFER)
        crc = zlib.crc32(img.tobytes()) & 0xFFFFFFFF
        alpha = float(img[..., 3].astype(np.float64).sum())
    else:
        ctx.sync()
    pt.refresh()
    return crc, alpha


# ------------------------------------------------------------------------------ counted work -> FP32 roofline
def scene_eval_flops(objs) -> int:
    """Algorithmic flops of ONE scene evaluation (SURVEY.md 8(d) convention), the neural bunny counted by its cheap
    outer branch (its MLP evaluations are counted separately)."""
    from raytracingpbr_b200 import scenes
    per = {scenes.SHAPE_BOX: FLOP_BOX, scenes.SHAPE_SPHERE: FLOP_SPHERE, scenes.SHAPE_CYLINDER: FLOP_CYLINDER,
           scenes.SHAPE_BUNNY: FLOP_BUNNY_OUTER}
    return sum(per.get(o.type, FLOP_BOX) for o in objs)                      # (8(d)'s per-primitive figures include the min / argmin step)


def counted_work(preset: str, width: int, height: int, bounces: int, spp: int, device: int, env=None):
    """Untimed pass with the counting twin of the kernel (ahead-of-time, the reference's evaluation sequence)."""
    from raytracingpbr_b200 import PathTracer, scenes
    cfg, objs, cam, tm = getattr(scenes, preset)(width, height, max_bounces=bounces, seed=0, count_work=True)
    with PathTracer(cfg, objs, cam, tm, device=device) as cpt:
        if env is not None:
            cpt.set_envmap(env)
        cpt.refresh()
        cpt.pathtrace(spp)
        cpt.sync()
        c = cpt.ctx.counters()
    n = max(c["samples"], 1)
    per = {k: c[k] / n for k in ("scene_evals", "rays", "normals", "mlp_evals")}
    box_like = scene_eval_flops(objs)
    flop = (per["scene_evals"] * box_like + per["normals"] * 4 * (box_like / max(len(objs), 1)) + per["rays"] * FLOP_PER_RAY_SHADE
            + per["mlp_evals"] * FLOP_BUNNY_MLP)
    lane = (c["march_active"] / c["march_iters"]) if c["march_iters"] else None
    return per, flop, lane, box_like


def fp32_roof(flop_per_sample: float, samples_per_launch: float, kernel_ms_per_launch: float, sm_count: int, sm_max_mhz: float):
    peak = sm_count * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    achieved = flop_per_sample * samples_per_launch / (kernel_ms_per_launch * 1e-3) / 1e12
    return achieved, peak


# ------------------------------------------------------------------------------ other configs
EXTRA_WORKLOADS = {
    # name: (preset, width, height, spp, max bounces, (exposure, gamma), asset file, description)
    "c2": ("bunny_glass", 1024, 1024, 256, 16, (1.8, 2.2), "limpopo_golf_course_3k.hdr",
           "C2: bunny_sdf_glass scene, frame 0, 1024x1024, 256 spp, max 16 bounces"),
    "c3": ("tokyo_ibl", 1920, 1080, 128, 8, (1.8, 2.2), "Tokyo_BigSight_3k.hdr",
           "C3: tokyo_ibl scene, 1920x1080, 128 spp, max 8 bounces"),
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


def environment(asset: str, exposure: float, gamma: float):
    """(processed table, description): the real .hdr when $RTPBR_ASSETS holds it, else the procedural stand-in."""
    from raytracingpbr_b200 import ibl
    for d in (os.environ.get("RTPBR_ASSETS"), os.path.join(ROOT, "tests", "assets_local")):
        if d and os.path.exists(os.path.join(d, asset)):
            return ibl.process(ibl.imread(os.path.join(d, asset)), exposure, gamma), f"the reference's own map assets/{asset} (staged copy)"
    return ibl.process(synthetic_env_u8(), exposure, gamma), "synthetic (procedural 3200x1600 environment; the .hdr is not staged)"


def run_extra(name: str, steps: int, warmup: int, device: int = 0, sm_max_mhz: float = 1965.0, with_clocks: bool = False) -> dict:
    """One of BASELINE configs[2] / configs[3] on one GPU: timing + counted-work fraction of the FP32 peak."""
    from raytracingpbr_b200 import PathTracer, scenes
    preset, w, h, spp, bounces, (exposure, gamma), asset, desc = EXTRA_WORKLOADS[name]
    env, env_desc = environment(asset, exposure, gamma)
    cfg, objs, cam, tm = getattr(scenes, preset)(w, h, max_bounces=bounces, seed=0)
    with PathTracer(cfg, objs, cam, tm, device=device) as pt:
        pt.set_envmap(env)
        ctx = pt.ctx

        def step():
            ctx.flush_l2()
            pt.refresh()
            ctx.set_sample_base(0)
            pt.pathtrace(spp)
        for _ in range(warmup):
            step()
        ctx.sync()
        ctx.kernel_time()
        sampler = ClockSampler(device) if with_clocks else None
        if sampler:
            sampler.start()
        ctx.timer_start()
        for _ in range(steps):
            step()
        ms = ctx.timer_stop()
        clocks = sampler.stop() if sampler else None
        kernel_ms, launches = ctx.kernel_time()
        info = ctx.device_info()
        jit = ctx.jit_status()[1]
    small = 4 if name == "c2" else 2                       # counting pass at 1/4 (1/2) of the resolution: per-sample work
    per, flop, lane, eval_flops = counted_work(preset, w // small, h // small, bounces, 2, device, env)
    k_ms = kernel_ms / max(launches, 1)
    achieved, peak = fp32_roof(flop, w * h * spp * steps / max(launches, 1), k_ms, info["sm_count"], sm_max_mhz)
    prof, prof_file = ncu_summary(name)
    out = {"workload": desc, "value": w * h * spp / (ms / steps * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
           "warmup": warmup, "kernel_ms_per_step": kernel_ms / steps, "kernel_launches_per_step": launches / steps, "environment": env_desc,
           "jit": jit, "blocks_per_sm": info["blocks_per_sm"],
           "roofline": {"bound": "fp32-issue", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                        "flop_per_sample": flop, "flop_per_scene_eval": eval_flops, "per_sample": per,
                        "counted_on": f"{w // small}x{h // small} x 2 spp, counting twin (ahead-of-time kernel, the reference's evaluation sequence)",
                        "issue_active_pct": prof.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                        "threads_per_instruction": prof.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
                        "ncu_source": prof_file}}
    if clocks:
        out["clocks"] = clocks
    return out


def extra_workload(args) -> int:
    """--workload c2 | c3: one JSON line for that configuration alone."""
    _, _, sm_max = peaks()
    b = run_extra(args.workload, args.steps, args.warmup, sm_max_mhz=sm_max, with_clocks=True)
    line = {"metric": "Msamples/s (pixels x spp / s)", "value": b["value"], "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": b["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": b["environment"], "config": {"workload": b["workload"], "jit": b["jit"]},
            "kernel_ms_per_step": b["kernel_ms_per_step"], "clocks": b.get("clocks"), "roofline": b["roofline"]}
    print(json.dumps(line), flush=True)
    return 0


def strong_job(R: Ranks, width: int, height: int, spp: int, warmup: int, steps: int, kernel: int, desc: str) -> dict | None:
    """A FIXED width x height x spp job split over the ranks (BASELINE's multi-GPU case).  Rank 0 gets the dict."""
    pt, objs, cam, tm = sharded_tracer(R, width, height, kernel)
    try:
        pt.refresh()                     # untimed short pass, whatever `warmup` is: the specialised kernel of this resolution is
        pt.pathtrace(min(spp, 16))       # built (NVRTC, ~2 s), the per-sample scratch grows to its chunk size (4 GiB at 4096^2) and
        if R.world > 1:                  # NCCL sets up its channels (first collective: ~0.3 s) before the clock starts
            pt.reduce_tiles(0)
        pt.sync()
        t = timed_steps(R, pt, spp, warmup, steps)
        red = reduce_alone_ms(R, pt, spp)
        if width * height * spp <= (1 << 28):
            crc, alpha = image_crc(R, pt, spp)
            crc_note = f"crc of the reduced {spp} spp accumulation buffer"
        else:                                            # big jobs: the CRC of a 2-spp pass of the same sharded pipeline
            crc, alpha = image_crc(R, pt, 2)
            crc_note = "crc of a 2 spp pass (same sharding, same reduce)"
    finally:
        pt.close()
    if R.rank != 0:
        return None
    total = float(width) * height * spp
    return {"workload": desc, "n_gpus": R.world, "scaling": "strong", "width": width, "height": height, "spp": spp, "steps": steps,
            "warmup": warmup, "ms_per_step": t["ms_per_step"], "value": total / (t["ms_per_step"] * 1e-3) / 1e6, "unit": UNIT,
            "kernel_ms_per_step_by_rank": [k * t["kernel_launches"] / steps for k in t["kernel_ms_per_launch_by_rank"]],
            "kernel_launches_per_step": t["kernel_launches"] / steps,
            "reduce_tiles_ms": red, "image_crc32": crc, "crc_note": crc_note, "alpha_checksum": alpha,
            "sharding": f"{R.world} ranks, {BAND}-column interleaved bands" if R.world > 1 else "none"}


# ------------------------------------------------------------------------------ GPU arm
def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-blocks", action="store_true", help="headline only: skip the strong / c4 / workloads blocks")
    ap.add_argument("--kernel", default="persistent", choices=["persistent", "simple"])
    ap.add_argument("--workload", default="c1", choices=list(EXTRA_WORKLOADS) + ["c1", "c4"],
                    help="c1 = the headline (default) with its strong / c4 / workloads blocks; c2 / c3 = BASELINE.json configs[2] / "
                         "configs[3] alone; c4 = configs[4] alone: 4096 x 4096 x 1024 spp tile-sharded over the ranks (strong scaling)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args, int(os.environ.get("RANK", "0")))
    if args.workload in EXTRA_WORKLOADS:
        return extra_workload(args)

    from raytracingpbr_b200 import _native as N
    R = Ranks()
    rank, world = R.rank, R.world
    kernel = N.KERNEL_PERSISTENT if args.kernel == "persistent" else N.KERNEL_SIMPLE
    hbm_peak, peak_src, sm_max_mhz = peaks()

    if args.workload == "c4":              # configs[4] alone, as the line's headline (strong)
        blk = strong_job(R, 4096, 4096, 1024, 1, max(1, min(args.steps, 2)), kernel,
                         "C4: cornell_box_shortest scene, 4096x4096, 1024 spp, max 8 bounces, tile-sharded")
        if rank == 0:
            line = {"metric": METRIC.replace("1024^2", "4096^2"), "value": blk["value"], "unit": UNIT, "n_gpus": world, "steps": blk["steps"],
                    "warmup": blk["warmup"], "ms_per_step": blk["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                    "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": blk["workload"], "sharding": blk["sharding"]},
                    "c4": blk}
            print(json.dumps(line), flush=True)
        R.close()
        return 0

    # ---- leg 1: device-resident, weak scaling: per-GPU samples fixed (W*H*SPP), tiles sharded by column band --------
    spp = SPP * world
    pt, objs, cam, tm = sharded_tracer(R, W, H, kernel)
    ctx = pt.ctx
    sampler = ClockSampler(R.local) if rank == 0 else None
    t = timed_steps(R, pt, spp, args.warmup, args.steps, sampler)
    clocks = t["clocks"]
    ms_per_step = t["ms_per_step"]
    total_samples = float(W) * H * spp     # whole job, all ranks (each rank: W*H*SPP)
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
            pt.reduce_tiles(0)
        if rank == 0:
            ctx.download(N.BUF_IMAGE_BUFFER, host_img)     # device -> host, synchronises
        else:
            ctx.sync()

    step_e2e()
    R.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    ctx.sync()
    R.barrier()
    e2e_s = R.max(time.perf_counter() - t0)
    e2e_value = total_samples * args.steps / e2e_s / 1e6
    checksum = float(host_img[..., 3].astype(np.float64).sum()) if rank == 0 else 0.0
    weak_crc = (zlib.crc32(host_img.tobytes()) & 0xFFFFFFFF) if rank == 0 else None
    host_img = None
    pinned.free()
    info = ctx.device_info()
    jit = ctx.jit_status()[1]
    pt.refresh()

    # ---- rooflines of the dominant kernel ---------------------------------------------
    k_launches = max(t["kernel_launches"], 1)
    k_ms = t["kernel_ms_per_step_max"] * args.steps / k_launches           # average launch duration (slowest rank)
    launches_per_step = k_launches / args.steps                            # > 1 when the spp are chunked to the scratch budget
    local_pixels = W * H / world
    kname = "k_pathtrace_pool_jit" if (kernel == N.KERNEL_PERSISTENT and "active" in jit) else \
            ("k_pathtrace_pool" if kernel == N.KERNEL_PERSISTENT else "k_pathtrace_simple")
    roof = None
    if rank == 0:
        per, flop, lane, eval_flops = counted_work("cornell_box_shortest", W, H, BOUNCES, 4, R.local)
        achieved, peak_tf = fp32_roof(flop, W * H * SPP / launches_per_step, k_ms, info["sm_count"], sm_max_mhz)
        prof, prof_file = ncu_summary("c1")
        traffic = None
        if prof_file and world == 1:
            traffic = (prof.get("dram__bytes_read.sum", 0.0) + prof.get("dram__bytes_write.sum", 0.0)) * 1e6
        roof = {"bound": "fp32-issue", "kernel": kname, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "traffic": traffic,
                "traffic_note": (f"dram__bytes_read.sum + dram__bytes_write.sum of one launch at this configuration, read from {prof_file}"
                                 if traffic else "no tracked ncu capture of this configuration (N > 1 or file absent)"),
                "flop_per_sample": flop, "flop_per_scene_eval": eval_flops, "per_sample": per,
                "work_note": "ALGORITHMIC work: the reference algorithm's evaluation counts (counting twin of the kernel, 4 spp pass) x "
                             "SURVEY 8(d)'s flops; the product kernel reaches the same bits with fewer executed flops (walls as planes, "
                             "provable misses cut short)",
                "lane_utilisation_in_march_loop_of_counting_twin": lane,
                "peak_basis": f"{info['sm_count']} SMs x 128 lanes x 2 flop x {sm_max_mhz:.0f} MHz (clocks.max.sm; MEASURED_PEAKS.json has no FP32 entry)",
                "kernel_ms_per_launch": k_ms, "kernel_ms_per_launch_by_rank": t["kernel_ms_per_launch_by_rank"],
                "kernel_launches_per_step": launches_per_step, "kernel_share_of_step": t["kernel_ms_per_step_max"] / ms_per_step,
                "issue_active_pct": prof.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "threads_per_instruction": prof.get("smsp__thread_inst_executed_per_inst_executed.ratio"), "ncu_source": prof_file}
    alg_bytes = ALG_BYTES_PER_PIXEL_PER_LAUNCH * local_pixels
    design_bytes = (DESIGN_BYTES_PER_SAMPLE * local_pixels * spp / launches_per_step if kernel == N.KERNEL_PERSISTENT
                    else 32 * local_pixels)
    hbm = {"bound": "hbm", "kernel": kname, "achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
           "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": peak_src,
           "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_basis": "SURVEY 8(d): 16 B per pixel per launch (spp loop inside the kernel)",
           "design_bytes_per_launch": design_bytes,
           "design_basis": "16 B per SAMPLE: each path writes float4(radiance, 1) to the ordered-accumulation scratch, k_fold_samples adds them "
                           "in sample order (bit-exact accumulation with path-granular load balance)",
           "design_achieved_GBps": design_bytes / (k_ms * 1e-3) / 1e9,
           "traffic": roof["traffic"] if roof else None,
           "traffic_over_algorithmic": (roof["traffic"] / alg_bytes) if (roof and roof["traffic"]) else None,
           "traffic_over_design": (roof["traffic"] / design_bytes) if (roof and roof["traffic"]) else None,
           "note": "not HBM-bound by three orders of magnitude; reported because north_star asks for it"}
    pt.close()

    # ---- strong scaling at the C1 size, C4, the other configurations ------------------
    strong = c4 = None
    workloads = {}
    if not args.no_blocks and kernel == N.KERNEL_PERSISTENT:
        if world == 1:
            if rank == 0:                  # at N = 1 the strong job IS the headline step
                strong = {"workload": WORKLOAD, "n_gpus": 1, "scaling": "strong", "width": W, "height": H, "spp": SPP, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": ms_per_step, "value": value, "unit": UNIT,
                          "kernel_ms_per_step_by_rank": [t["kernel_ms_per_step_max"]], "kernel_launches_per_step": launches_per_step,
                          "reduce_tiles_ms": None, "image_crc32": weak_crc, "crc_note": f"crc of the {SPP} spp accumulation buffer (e2e leg download)",
                          "alpha_checksum": checksum, "sharding": "none"}
        else:
            strong = strong_job(R, W, H, SPP, args.warmup, args.steps, kernel, WORKLOAD)
        c4 = strong_job(R, 4096, 4096, 1024, 1 if world >= 4 else 0, 1, kernel,
                        "C4: cornell_box_shortest scene, 4096x4096, 1024 spp, max 8 bounces, tile-sharded")
        if world == 1 and rank == 0:
            for name in ("c2", "c3"):
                workloads[name] = run_extra(name, 2 if name == "c2" else 3, 1, R.local, sm_max_mhz)

    # ---- transparency: the same step with both scene analyses off (N = 1) ---------------
    plain = None
    if world == 1 and rank == 0 and not args.no_blocks and kernel == N.KERNEL_PERSISTENT:
        saved = {k: os.environ.get(k) for k in ("RTPBR_JIT_FAST", "RTPBR_JIT_BBOX")}
        os.environ.update(RTPBR_JIT_FAST="0", RTPBR_JIT_BBOX="0")          # read when the specialised kernel is built
        try:
            pt2, _, _, _ = sharded_tracer(R, W, H, kernel)
            t2 = timed_steps(R, pt2, SPP, 3, 3)
            crc2, _ = image_crc(R, pt2, SPP)
            jit2 = pt2.ctx.jit_status()[1]
            pt2.close()
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        plain = {"value": float(W) * H * SPP / (t2["ms_per_step"] * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": t2["ms_per_step"], "steps": 3,
                 "warmup": 3, "image_crc32": crc2, "jit": jit2,
                 "note": "RTPBR_JIT_FAST=0 RTPBR_JIT_BBOX=0: every scene evaluation of the reference's march is executed with the full box "
                         "formulas and missed rays are marched to t > MAX_DIS (the round-1 kernel); the headline kernel reaches the same "
                         "image (same crc) through two exact shortcuts: walls as planes inside a proven region, provable misses cut short"}

    # ---- CPU baseline (rank 0, N = 1 only) --------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v_aw, n_aw, dt_aw = run_cpu_oracle(CPU_SPP, hoisted=False)
        v_h, n_h, dt_h = run_cpu_oracle(CPU_SPP, hoisted=True)
        cpu = {"value": v_aw, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
               "sample": f"{CPU_BANDS * CPU_COLS} of 1024 columns ({CPU_BANDS} spread bands x {CPU_COLS}) x 1024 rows x {CPU_SPP} spp = {n_aw} samples, {dt_aw:.1f} s; "
                         "as-written (rotation matrices recomputed per object per march step like cornell_box_shortest.py:43)",
               "hoisted_value": v_h, "hoisted_sample": f"{n_h} samples, {dt_h:.1f} s (matrices precomputed)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "width": W, "height": H, "spp_per_step": spp, "max_bounces": BOUNCES,
                       "sharding": f"{world} ranks, {BAND}-column interleaved bands, spp x {world}" if world > 1 else "none",
                       "l2": "flushed between steps by a 256 MiB memset on the launch stream (inside the timed region)",
                       "kernel": args.kernel, "blocks_per_sm": info["blocks_per_sm"], "jit": jit},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "alpha_checksum": checksum, "image_crc32": weak_crc},
            "gpu_launches": t["launches"],
            "clocks": clocks,
            "roofline": roof,
            "roofline_hbm": hbm,
        }
        if strong:
            line["strong"] = strong
        if c4:
            line["c4"] = c4
        if workloads:
            line["workloads"] = workloads
        if plain:
            line["without_scene_analyses"] = plain
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    R.close()
    return 0


def _only_json_on_stdout() -> None:
    """The contract is ONE JSON line on stdout.  Libraries write there as well (NCCL prints its version banner to stdout
    when the box sets NCCL_DEBUG), so file descriptor 1 is pointed at stderr for the whole run and Python's sys.stdout keeps
    a private copy of the original descriptor for the line itself."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(keep, "w", buffering=1)


if __name__ == "__main__":
    _only_json_on_stdout()
    sys.exit(main())
