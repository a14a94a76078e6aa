#!/usr/bin/env python
"""Bench harness for the B200 TDBP backend (contract: see README / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config c2] [--scale S]

One "step" = one time-domain backprojection of the synthetic frame named by ``--config``
(default: BASELINE.json configs[1], the NISAR L-band 20 MHz frame 16384 pulses x 12288 bins
-> 8192 x 8192, flat DEM, tsx delay).  Prints ONE JSON line:

  value   pixel.pulses/s, whole job, range-compressed swath already resident in HBM
          (i3b_plan_execute: target solve + accumulation + finalisation on device)
  e2e     same metric through the reference-facing call ``backproject(out, ...)`` with HOST
          buffers: H2D of the swath and D2H of the image inside the timed region
  roofline      FP32 roofline of the accumulation kernel: algorithmic flops (34 + 10 K per
                pixel.pulse) / CUDA-event kernel time, against the FFMA rate
                measured on this device by i3b_measure_peaks
  cpu_baseline  the CPU oracle (oracle/_ref = reference sources compiled here when present,
                else the restated port) on a bounded sub-block of the same frame

N > 1 (torchrun, one process per GPU): the SAME frame is cut into N contiguous azimuth
blocks; rank r focuses block r from its own copy of the swath, no inter-GPU exchange;
time = max over ranks, value = total pixel.pulses / time ("scaling": "strong").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "pixel_pulses_per_s"
UNIT = "pixel*pulses/s"


def algorithmic_flops_per_pp(taps: int) -> float:
    """geometry 16 + index 4 + phase 6 + rotate/accumulate 8 + 10 per tap."""
    return 34.0 + 10.0 * taps


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [c for c, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(smax)),
                "power_w_max": float(max(power)), "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


def make_scene(args):
    from isce3_b200 import synth
    kw = {}
    if args.scale != 1.0:
        base = {"c1": (2048, 4096, 512, 512), "c2": (16384, 12288, 8192, 8192),
                "c4": (16384, 32768, 2048, 8192), "c5": (65536, 8192, 2048, 2048)}[args.config[:2]]
        s = args.scale
        kw = dict(pulses=max(int(base[0] * min(1.0, s * 2)), 512), bins=max(int(base[1] * s), 512),
                  out_lines=max(int(base[2] * s), 64), out_samples=max(int(base[3] * s), 128))
    if args.taps:
        kw["taps"] = args.taps
    return synth.make_scene(args.config, **kw)


def block_bounds(lines, world, rank):
    q, r = divmod(lines, world)
    a0 = rank * q + min(rank, r)
    return a0, a0 + q + (1 if rank < r else 0)


def workload_name(args, sc):
    ig, og = sc.in_geometry, sc.out_geometry
    return (f"{args.config}: {ig.grid_length} pulses x {ig.grid_width} bins -> "
            f"{og.grid_length} x {og.grid_width}, {'raster' if sc.dem.have_raster else 'flat'} DEM, "
            f"{sc.dry_tropo_model}, {sc.kernel.table.size if hasattr(sc.kernel, 'table') else 0}-entry "
            f"tabulated Knab, {int(np.ceil(sc.kernel.width))} taps")


def cpu_sample(sc, oracle, lines_wanted, seconds_target=None):
    """Oracle on a contiguous block of azimuth lines around the frame centre."""
    og = sc.out_geometry
    L = og.grid_length
    n = max(1, min(lines_wanted, L))
    a0 = max(0, L // 2 - n // 2)
    sub = sc.out_subgrid(a0, a0 + n)
    out = np.zeros((n, og.grid_width), np.complex64)
    t = time.perf_counter()
    oracle.backproject(out, sub, sc.rc, sc.in_geometry, sc.dem, sc.fc, sc.ds, sc.kernel,
                       sc.dry_tropo_model, sc.rdr2geo_params, sc.geo2rdr_params)
    dt = time.perf_counter() - t
    return dt, a0, n, out


def run_reference(args):
    """--impl reference: the reference's own CPU backprojection (oracle/_ref) on host cores."""
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)  # torchrun exports 1; the reference uses every core
    from oracle import tdbp
    tdbp.set_threads(cores)
    oracle = tdbp.best()
    sc = make_scene(args)
    og = sc.out_geometry
    # bounded sample: a few azimuth lines x full range width, sized for ~10 s per step
    pulses_per_pixel = min(sc.in_geometry.grid_length, 4400)
    lines = max(1, int(2.0e8 * max(cores, 1) / 8 / (pulses_per_pixel * og.grid_width)))
    times = []
    pp = None
    for i in range(args.warmup_ref + args.steps):
        dt, a0, n, _ = cpu_sample(sc, oracle, lines)
        if pp is None:
            pp = estimate_pp(sc, a0, n)
        if i >= args.warmup_ref:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = pp / (ms * 1e-3)
    sample = f"{n} azimuth lines x {og.grid_width} range pixels at the frame centre per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup_ref, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args, sc), "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": oracle.kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def estimate_pp(sc, a0, n):
    """pixel.pulses of an azimuth block without a GPU: sum(kstop-kstart) from the port
    oracle's counter at 8 range columns of the block's first line (the aperture length of
    Backproject.cpp:176-193 varies smoothly with range and not with azimuth)."""
    import ctypes as C

    from isce3_b200.focus import build_args
    from oracle import tdbp
    port = tdbp.port()
    width = sc.out_geometry.grid_width
    cols = np.unique(np.linspace(0, width - 1, 8).astype(int))
    total = 0.0
    for c in cols:
        g = sc.out_subgrid(a0, a0 + 1, int(c), int(c) + 1)
        out = np.zeros((1, 1), np.complex64)
        fl = build_args(out, g, sc.rc, sc.in_geometry, sc.dem, sc.fc, sc.ds, sc.kernel,
                        sc.dry_tropo_model, sc.rdr2geo_params, sc.geo2rdr_params)
        pp = C.c_double(0)
        port._backproject_pp(C.byref(fl.args), C.byref(pp))
        total += pp.value
    return total / len(cols) * width * n


def run_ours(args):
    rank, world, local = dist_env()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    from isce3_b200 import focus
    from isce3_b200.focus import BackprojectPlan, backproject, last_stats, measure_peaks

    t_gen = time.perf_counter()
    sc = make_scene(args)
    t_gen = time.perf_counter() - t_gen
    og = sc.out_geometry
    a0, a1 = block_bounds(og.grid_length, world, rank)
    sub = sc.out_subgrid(a0, a1) if world > 1 else og
    shape = (a1 - a0, og.grid_width)
    taps = int(np.ceil(sc.kernel.width))
    common = (sc.in_geometry, sc.dem, sc.fc, sc.ds, sc.kernel, sc.dry_tropo_model,
              sc.rdr2geo_params, sc.geo2rdr_params)

    def barrier():
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peaks = measure_peaks(local)

    # host buffers for the end-to-end arm: pinned when torch is importable
    rc_host = sc.rc
    pinned_note = "pageable"
    try:
        import torch
        pin = torch.empty(sc.rc.shape, dtype=torch.complex64, pin_memory=True)
        pin_np = pin.numpy()
        pin_np[...] = sc.rc
        rc_host = pin_np
        pinned_note = "pinned"
        out_pin = torch.empty(shape, dtype=torch.complex64, pin_memory=True)
        out_host = out_pin.numpy()
    except Exception:
        out_host = np.empty(shape, np.complex64)

    # ---- resident arm -------------------------------------------------------------
    plan = BackprojectPlan(sub, sc.rc, *common, batch=args.batch, devices=[local])
    for _ in range(args.warmup):
        plan.execute()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    kernel_ms, solve_ms, launches = 0.0, 0.0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        plan.execute()
        st = plan.stats()
        kernel_ms += st["ms_accumulate"]
        solve_ms += st["ms_target_solve"]
        launches += st["total_launches"]
    barrier()
    elapsed = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    st = plan.stats()
    elapsed = max_over_ranks(elapsed)
    pp_rank = st["pixel_pulses"]
    pp_total = sum_over_ranks(pp_rank)
    ms_per_step = 1e3 * elapsed / args.steps
    value = pp_total / (ms_per_step * 1e-3)
    kernel_ms_step = kernel_ms / args.steps
    img = plan.download()
    plan.close()

    # ---- end-to-end arm: host buffers through the reference-shaped call ---------------
    for _ in range(min(args.warmup, 3)):
        backproject(out_host, sub, rc_host, *common, batch=args.batch, devices=[local])
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    e2e_launches = 0
    for _ in range(args.steps):
        backproject(out_host, sub, rc_host, *common, batch=args.batch, devices=[local])
        s2 = last_stats()
        h2d, d2h = s2["h2d_bytes"], s2["d2h_bytes"]
        e2e_launches += s2["total_launches"]
    barrier()
    e2e_elapsed = max_over_ranks(time.perf_counter() - t0)
    e2e_value = pp_total / (e2e_elapsed / args.steps)
    same = float(np.nanmax(np.abs(out_host - img))) if img.size else 0.0

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    flops_pp = algorithmic_flops_per_pp(taps)
    achieved_tflops = flops_pp * pp_rank / (kernel_ms_step * 1e-3) / 1e12 if kernel_ms_step > 0 else 0.0
    nominal_fp32 = 148 * 128 * 2 * 1965e6 / 1e12
    peaks_file = {}
    try:
        peaks_file = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    ig = sc.in_geometry
    alg_bytes = 8.0 * (st["pulse_last"] - st["pulse_first"]) * ig.grid_width + 8.0 * shape[0] * shape[1] \
        + 40.0 * shape[0] * shape[1]
    roofline = {
        "bound": "fp32", "kernel": f"accumulate_fast_kernel<{taps}>" if st["used_fast_kernel"] else "accumulate_generic_kernel",
        "achieved": achieved_tflops, "peak": peaks["fp32_tflops"], "unit": "TFLOP/s",
        "frac": achieved_tflops / peaks["fp32_tflops"] if peaks["fp32_tflops"] else None,
        "peak_source": "i3b_measure_peaks FFMA microbenchmark on this device (MEASURED_PEAKS.json has no FP32 peak)",
        "frac_of_nominal_74.5": achieved_tflops / nominal_fp32,
        "flops_per_pixel_pulse": flops_pp, "kernel_ms_per_step": kernel_ms_step,
        "kernel_share_of_step": kernel_ms_step / ms_per_step,
        "pp_per_s_kernel": pp_rank / (kernel_ms_step * 1e-3) if kernel_ms_step > 0 else None,
        "sfu_frac": (3.0 * pp_rank / (kernel_ms_step * 1e-3) / 1e9) / peaks["sfu_gops"] if kernel_ms_step > 0 else None,
        "hbm_algorithmic_gbs": alg_bytes / (kernel_ms_step * 1e-3) / 1e9 if kernel_ms_step > 0 else None,
        "hbm_peak_gbs": peaks_file.get("hbm_gbs"),
        "measured_peaks": peaks, "traffic": None,
    }
    # DRAM traffic of the dominant kernel: one ncu capture of this workload, committed under
    # profiles/ (per launch, like `achieved`); only quoted for the workload it was taken on
    try:
        tr = json.loads((ROOT / "profiles" / "r01_traffic.json").read_text())
        if world == 1 and args.config == "c2" and args.scale == 1.0 and not args.taps:
            roofline["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            roofline["traffic_unit"] = "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)"
            roofline["traffic_source"] = tr["source"]
            roofline["algorithmic_bytes_per_launch"] = alg_bytes + 32.0 * shape[0] * shape[1]
    except Exception:
        pass

    cpu = None
    if not args.no_cpu:
        from oracle import tdbp
        cores = os.cpu_count() or 1
        tdbp.set_threads(cores)  # torchrun exports OMP_NUM_THREADS=1 to its workers
        oracle = tdbp.best()
        ppx = pp_rank / max(shape[0] * shape[1], 1)
        lines = max(1, int(2.5e8 * cores / 8 / max(ppx * og.grid_width, 1)))
        dt, b0, n, ref = cpu_sample(sc, oracle, lines)
        pp_cpu = ppx * n * og.grid_width
        lo = b0 - a0
        parity = None
        if world == 1:
            g = img[lo:lo + n]
            m = np.isfinite(ref)
            parity = float(np.linalg.norm((g - ref)[m]) / max(np.linalg.norm(ref[m]), 1e-30))
        cpu = {"value": pp_cpu / dt, "unit": UNIT, "cores": cores, "kind": oracle.kind,
               "sample": f"{n} azimuth lines x {og.grid_width} range pixels at the frame centre "
                         f"({pp_cpu:.3g} pixel*pulses, {dt:.1f} s)",
               "rel_rms_gpu_vs_cpu_on_sample": parity}

    # ---- same-GPU comparator: the reference's own CUDA backprojection (reduced harness) ---
    ref_cuda = None
    if not args.no_ref_cuda and world == 1:
        try:
            from oracle import tdbp
            if tdbp.have_ref_cuda() and not sc.dem.have_raster:
                rcu = tdbp.ref_cuda()
                small = sc.out_subgrid(0, min(8, shape[0]))
                rcu.backproject(np.zeros((min(8, shape[0]), shape[1]), np.complex64), small, rc_host,
                                *common, batch=args.batch)  # module load / first-call costs
                ref_out = np.empty(shape, np.complex64)
                t0 = time.perf_counter()
                rcu.backproject(ref_out, sub, rc_host, *common, batch=args.batch)
                dt = time.perf_counter() - t0
                m = np.isfinite(ref_out) & np.isfinite(out_host)
                ref_cuda = {
                    "value": pp_rank / dt, "unit": UNIT, "ms": 1e3 * dt, "kind": rcu.kind,
                    "what": "isce3::cuda::focus::backproject (cuda/focus/Backproject.cu, unmodified) "
                            "on the same frame, same host buffers, one call, wall time",
                    "e2e_speedup_of_this_repo": e2e_value / (pp_rank / dt),
                    "rel_rms_ours_vs_reference_cuda": float(
                        np.linalg.norm((out_host - ref_out)[m]) / max(np.linalg.norm(ref_out[m]), 1e-30)),
                }
            elif sc.dem.have_raster:
                ref_cuda = {"unavailable": "reduced harness supports constant-height DEMs only"}
            else:
                ref_cuda = {"unavailable": "oracle/_ref/libtdbp_refcuda.so not built"}
        except Exception as exc:  # the comparator must never take the bench down
            ref_cuda = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64+f32", "data": "synthetic",
        "config": {"workload": workload_name(args, sc), "sharding": "contiguous azimuth blocks of one frame, no exchange",
                   "l2": "inputs (swath + per-pixel tables) far exceed the 126 MB L2",
                   "pixel_pulses_per_step": pp_total, "batch": args.batch,
                   "scene_generation_s": t_gen},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "host_memory": pinned_note,
                "ms_per_step": 1e3 * e2e_elapsed / args.steps,
                "max_abs_diff_vs_resident": same},
        "gpu_launches": int(launches),
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "reference_cuda": ref_cuda,
        "target_solve_ms_per_step": solve_ms / args.steps,
        # CUDA-event time of the two kernels that make up a resident step (rank 0); the
        # difference to ms_per_step (wall, barrier to barrier) is finalisation + host overhead
        "device_ms_per_step": kernel_ms_step + solve_ms / args.steps,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the named config (dev runs)")
    ap.add_argument("--taps", type=int, default=0)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true",
                    help="skip the reference-CUDA comparator (one ~8 s call at C2)")
    args = ap.parse_args()
    args.warmup_ref = min(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
