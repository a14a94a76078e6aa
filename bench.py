#!/usr/bin/env python
"""Bench harness for the B200 TDBP backend (contract: see README / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config c2] [--scale S] [--no-extras]

One "step" = one time-domain backprojection of the synthetic frame named by ``--config``
(default: BASELINE.json configs[1], the NISAR L-band 20 MHz frame 16384 pulses x 12288 bins
-> 8192 x 8192, flat DEM, tsx delay).  Prints ONE JSON line:

  value   pixel.pulses/s, whole job, range-compressed swath already resident in HBM
          (i3b_plan_execute: target solve + accumulation + finalisation on device)
  e2e     same metric through the reference-facing call ``backproject(out, ...)`` with HOST
          buffers: H2D of the swath and D2H of the image inside the timed region (pinned
          host memory; ``e2e_pageable`` repeats it from ordinary pageable numpy arrays, which is
          what the workflow hands over, focus.py:1857-1859)
  roofline      FP32 roofline of the accumulation kernel: algorithmic flops (34 + 10 K per
                pixel.pulse) / CUDA-event kernel time, against the FFMA rate
                measured on this device by i3b_measure_peaks; ``roofline_general`` is the same
                frame through the general (non-baked coefficient) kernel
  cpu_baseline  the CPU oracle (oracle/_ref = reference sources compiled here when present,
                else the restated port) on a bounded sub-block of the same frame, all host
                cores and one thread
  configs       every other BASELINE.json shape (c1, c4, c5 with 8/16/32 taps) at this N: a few
                steps each -- value, e2e, roofline fraction, parity against the CPU oracle
  inlib_multi_gpu  (N > 1) rank 0 alone calls backproject(out, ..., devices=[0..N-1]) on the
                whole frame from ONE pageable input into ONE pageable output: the in-library
                multi-GPU driver, next to the torchrun number

N > 1 (torchrun, one process per GPU): the SAME frame is cut into N contiguous azimuth
blocks; rank r focuses block r from its own copy of the swath, no inter-GPU exchange;
time = max over ranks, value = total pixel.pulses / time ("scaling": "strong").
"""
from __future__ import annotations

import argparse
import gc
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
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1965e6 / 1e12

BASE_SHAPES = {"c1": (2048, 4096, 512, 512), "c2": (16384, 12288, 8192, 8192),
               "c4": (16384, 32768, 2048, 8192), "c5": (65536, 8192, 2048, 2048)}


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


def bind_near_gpu(local):
    """One rank per GPU: run this rank (and allocate its host buffers, first touch) on the CPUs of
    the GPU's own NUMA node, as a multi-GPU launcher does (numactl / --cpu-bind).  Returns
    (cpus bound to, cpus allowed before) or (None, allowed) when nothing could be learnt."""
    allowed = os.sched_getaffinity(0)
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        bus_id = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(allowed) + 64) // 64)
        near = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1} & allowed
        if near and near != allowed:
            os.sched_setaffinity(0, near)
            return sorted(near), allowed
    except Exception:  # noqa: BLE001 -- binding is an optimisation, never a requirement
        pass
    return None, allowed


def make_scene(config, scale=1.0, taps=0, **extra):
    from testkit import synth
    kw = dict(extra)
    if scale != 1.0:
        base = BASE_SHAPES[config[:2]]
        s = scale
        kw.update(pulses=max(int(base[0] * min(1.0, s * 2)), 512), bins=max(int(base[1] * s), 512),
                  out_lines=max(int(base[2] * s), 64), out_samples=max(int(base[3] * s), 128))
    if taps:
        kw["taps"] = taps
    return synth.make_scene(config, **kw)


def block_bounds(lines, world, rank):
    q, r = divmod(lines, world)
    a0 = rank * q + min(rank, r)
    return a0, a0 + q + (1 if rank < r else 0)


def workload_name(config, sc):
    ig, og = sc.in_geometry, sc.out_geometry
    return (f"{config}: {ig.grid_length} pulses x {ig.grid_width} bins -> "
            f"{og.grid_length} x {og.grid_width}, {'raster' if sc.dem.have_raster else 'flat'} DEM, "
            f"{sc.dry_tropo_model}, {sc.kernel.table.size if hasattr(sc.kernel, 'table') else 0}-entry "
            f"tabulated Knab, {int(np.ceil(sc.kernel.width))} taps")


def cpu_sample(sc, oracle, lines_wanted, cols=None):
    """Oracle on a contiguous block of azimuth lines around the frame centre."""
    og = sc.out_geometry
    L = og.grid_length
    n = max(1, min(lines_wanted, L))
    a0 = max(0, L // 2 - n // 2)
    sub = sc.out_subgrid(a0, a0 + n) if cols is None else sc.out_subgrid(a0, a0 + n, 0, cols)
    width = og.grid_width if cols is None else cols
    out = np.zeros((n, width), np.complex64)
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
    sc = make_scene(args.config, args.scale, args.taps)
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
        "config": {"workload": workload_name(args.config, sc), "sample": sample,
                   "pixel_pulses_counted_by": "sum(kstop - kstart) from the oracle's own aperture "
                                              "bounds at 8 range columns of the sample's first line, "
                                              "times lines x width (apertures vary smoothly with "
                                              "range and not with azimuth)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": oracle.kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
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


class Comm:
    """torch.distributed plumbing of the bench (NCCL for the timing reductions, a gloo group
    for host-only barriers); a no-op at N = 1."""

    def __init__(self, rank, world, local):
        self.rank, self.world, self.local = rank, world, local
        self.dist = None
        self.cpu_group = None
        if world > 1:
            import torch
            import torch.distributed as dist_mod
            torch.cuda.set_device(local)
            dist_mod.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
            self.dist = dist_mod
            self.cpu_group = dist_mod.new_group(backend="gloo")

    def barrier(self):
        if self.dist is not None:
            import torch
            torch.cuda.synchronize()
            self.dist.barrier()

    def host_barrier(self):
        """No GPU work on any rank (an NCCL barrier would spin kernels on the waiting GPUs)."""
        if self.dist is not None:
            self.dist.barrier(group=self.cpu_group)

    def _reduce(self, x, op):
        if self.dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX) if self.dist else x

    def sum(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM) if self.dist else x

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def host_buffers(sc, shape, want_pinned=True):
    """(rc_host, out_host, kind): pinned through torch when available, else pageable numpy."""
    if want_pinned:
        try:
            import torch
            pin = torch.empty(sc.rc.shape, dtype=torch.complex64, pin_memory=True)
            rc_host = pin.numpy()
            rc_host[...] = sc.rc
            out_pin = torch.empty(shape, dtype=torch.complex64, pin_memory=True)
            return rc_host, out_pin.numpy(), "pinned", (pin, out_pin)
        except Exception:
            pass
    return sc.rc, np.empty(shape, np.complex64), "pageable", None


def measure(comm, sc, steps, warmup, batch, sampler=None, e2e_steps=None, pinned=True):
    """Resident and end-to-end timing of one frame at this N.  Returns a dict; the image of
    this rank's block is under "img" (popped by the caller)."""
    from isce3_b200.focus import BackprojectPlan, backproject, last_stats
    og = sc.out_geometry
    a0, a1 = block_bounds(og.grid_length, comm.world, comm.rank)
    sub = sc.out_subgrid(a0, a1) if comm.world > 1 else og
    shape = (a1 - a0, og.grid_width)
    common = (sc.in_geometry, sc.dem, sc.fc, sc.ds, sc.kernel, sc.dry_tropo_model,
              sc.rdr2geo_params, sc.geo2rdr_params)
    rc_host, out_host, host_kind, keep = host_buffers(sc, shape, pinned)

    plan = BackprojectPlan(sub, sc.rc, *common, batch=batch, devices=[comm.local])
    for _ in range(warmup):
        plan.execute()
    if sampler is not None:
        sampler.start()
    comm.barrier()
    kernel_ms = solve_ms = 0.0
    launches = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        plan.execute()
        st = plan.stats()
        kernel_ms += st["ms_accumulate"]
        solve_ms += st["ms_target_solve"]
        launches += st["total_launches"]
    comm.barrier()
    elapsed = comm.max(time.perf_counter() - t0)
    clocks = sampler.stop() if sampler is not None else None
    st = plan.stats()
    pp_rank = st["pixel_pulses"]
    pp_total = comm.sum(pp_rank)
    ms_per_step = 1e3 * elapsed / steps
    img = plan.download()
    plan.close()

    # ---- end-to-end arm: host buffers through the reference-shaped call ---------------
    e2e_steps = e2e_steps or steps
    for _ in range(min(warmup, 3)):
        backproject(out_host, sub, rc_host, *common, batch=batch, devices=[comm.local])
    comm.barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(e2e_steps):
        backproject(out_host, sub, rc_host, *common, batch=batch, devices=[comm.local])
        s2 = last_stats()
        h2d, d2h = s2["h2d_bytes"], s2["d2h_bytes"]
    comm.barrier()
    e2e_elapsed = comm.max(time.perf_counter() - t0)
    same = float(np.nanmax(np.abs(out_host - img))) if img.size else 0.0
    taps = int(np.ceil(sc.kernel.width))
    kms = kernel_ms / steps
    return {
        "value": pp_total / (ms_per_step * 1e-3), "ms_per_step": ms_per_step, "pp_total": pp_total,
        "pp_rank": pp_rank, "kernel_ms_step": kms, "solve_ms_step": solve_ms / steps,
        "launches": launches, "clocks": clocks, "stats": st, "taps": taps, "shape": shape,
        "block": (a0, a1), "sub": sub, "common": common, "rc_host": rc_host, "out_host": out_host,
        "host_kind": host_kind, "_keep": keep, "img": img,
        "e2e": {"value": pp_total / (e2e_elapsed / e2e_steps), "unit": UNIT,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "host_memory": host_kind, "ms_per_step": 1e3 * e2e_elapsed / e2e_steps,
                "max_abs_diff_vs_resident": same},
        "achieved_tflops": algorithmic_flops_per_pp(taps) * pp_rank / (kms * 1e-3) / 1e12 if kms > 0 else 0.0,
    }


def cpu_parity(comm, sc, m, oracle, cores, budget_pp=2.5e8):
    """CPU oracle on a bounded sample inside this rank's block; returns (cpu dict, parity)."""
    og = sc.out_geometry
    a0, a1 = m["block"]
    shape = m["shape"]
    ppx = m["pp_rank"] / max(shape[0] * shape[1], 1)
    lines = max(1, int(budget_pp * cores / 8 / max(ppx * og.grid_width, 1)))
    lines = min(lines, shape[0])
    # sample centred in this rank's block
    c = (a0 + a1) // 2
    b0 = max(a0, min(c - lines // 2, a1 - lines))
    sub = sc.out_subgrid(b0, b0 + lines)
    ref = np.zeros((lines, og.grid_width), np.complex64)
    t = time.perf_counter()
    oracle.backproject(ref, sub, sc.rc, sc.in_geometry, sc.dem, sc.fc, sc.ds, sc.kernel,
                       sc.dry_tropo_model, sc.rdr2geo_params, sc.geo2rdr_params)
    dt = time.perf_counter() - t
    g = m["img"][b0 - a0:b0 - a0 + lines]
    ok = np.isfinite(ref)
    parity = float(np.linalg.norm((g - ref)[ok]) / max(np.linalg.norm(ref[ok]), 1e-30))
    pp_cpu = ppx * lines * og.grid_width
    return {"value": pp_cpu / dt, "unit": UNIT, "cores": cores, "kind": oracle.kind,
            "sample": f"{lines} azimuth lines x {og.grid_width} range pixels at the centre of rank 0's block "
                      f"({pp_cpu:.3g} pixel*pulses, {dt:.1f} s)",
            "rel_rms_gpu_vs_cpu_on_sample": parity}, parity


def run_ours(args):
    rank, world, local = dist_env()
    comm = Comm(rank, world, local)
    from isce3_b200 import focus
    from isce3_b200.focus import backproject, last_stats, measure_peaks, release_device_memory

    # opt in to the device-buffer cache, as a workflow focusing block after block would
    # (default: every call hands its device memory back to the driver)
    focus.keep_device_memory(-1)
    peaks = measure_peaks(local)
    cores = os.cpu_count() or 1
    oracle = None
    if rank == 0 and not args.no_cpu:
        from oracle import tdbp
        tdbp.set_threads(cores)  # torchrun exports OMP_NUM_THREADS=1 to its workers
        oracle = tdbp.best()

    bound_cpus, all_cpus = bind_near_gpu(local) if world > 1 else (None, os.sched_getaffinity(0))
    t_gen = time.perf_counter()
    sc = make_scene(args.config, args.scale, args.taps)
    t_gen = time.perf_counter() - t_gen
    og = sc.out_geometry
    in_width = sc.in_geometry.grid_width
    sampler = ClockSampler(local) if rank == 0 else None
    m = measure(comm, sc, args.steps, args.warmup, args.batch, sampler)
    st, taps, shape = m["stats"], m["taps"], m["shape"]
    pp_rank, kernel_ms_step = m["pp_rank"], m["kernel_ms_step"]
    rc_host, out_host, sub, common = m["rc_host"], m["out_host"], m["sub"], m["common"]

    # ---- the general (non-baked coefficient) kernel on the same frame ------------------
    roofline_general = None
    if st["used_fast_kernel"] and st["fast_variant"] >= 0:
        os.environ["I3B_FAST_NO_IMM"] = "1"
        try:
            g = measure_resident_only(comm, sc, 2, 1, args.batch)
        finally:
            del os.environ["I3B_FAST_NO_IMM"]
        if rank == 0:
            ach = algorithmic_flops_per_pp(taps) * g["pp_rank"] / (g["kernel_ms_step"] * 1e-3) / 1e12
            roofline_general = {"kernel": f"accumulate_fast_kernel<{taps}, CoefBank> (coefficients from shared memory)",
                                "fast_variant": g["fast_variant"], "achieved": ach, "unit": "TFLOP/s",
                                "frac": ach / peaks["fp32_tflops"], "kernel_ms_per_step": g["kernel_ms_step"],
                                "value": g["value"]}

    # ---- end to end from pageable host memory (what the workflow passes) -----------------
    e2e_pageable = None
    if not args.no_extras:
        out_pg = np.empty(shape, np.complex64)
        h_pg = np.empty(shape, np.float32)
        backproject(out_pg, sub, sc.rc, *common, batch=args.batch, devices=[local], height=h_pg)
        comm.barrier()
        t0 = time.perf_counter()
        n_pg = 2
        for _ in range(n_pg):
            backproject(out_pg, sub, sc.rc, *common, batch=args.batch, devices=[local], height=h_pg)
        comm.barrier()
        dt = comm.max(time.perf_counter() - t0)
        e2e_pageable = {"value": m["pp_total"] / (dt / n_pg), "unit": UNIT, "ms_per_step": 1e3 * dt / n_pg,
                        "host_memory": "pageable in, pageable out and height (numpy arrays)",
                        "bit_identical_to_resident": bool(np.array_equal(out_pg, m["img"]))}
        del out_pg, h_pg

    # ---- in-library multi-GPU driver: one process, one pageable in / out ------------------
    inlib = None
    if world > 1 and not args.no_extras:
        release_device_memory()
        comm.host_barrier()
        if rank == 0:
            os.sched_setaffinity(0, all_cpus)  # one process drives every device from here
            full = np.empty((og.grid_length, og.grid_width), np.complex64)
            devs = list(range(world))
            full_args = (og, sc.rc) + common
            backproject(full, *full_args, batch=args.batch, devices=devs)
            t0 = time.perf_counter()
            n_il = 3
            for _ in range(n_il):
                backproject(full, *full_args, batch=args.batch, devices=devs)
            dt = (time.perf_counter() - t0) / n_il
            s2 = last_stats()
            a0, a1 = m["block"]
            inlib = {"value": s2["pixel_pulses"] / dt, "unit": UNIT, "ms_per_step": 1e3 * dt,
                     "n_devices": s2["n_devices"], "h2d_bytes_per_step": int(s2["h2d_bytes"]),
                     "d2h_bytes_per_step": int(s2["d2h_bytes"]),
                     "what": "rank 0 alone: backproject(out, ..., devices=[0..N-1]) on the whole frame, "
                             "ONE pageable input array, ONE pageable output array, one host thread per device",
                     "vs_torchrun_e2e": (s2["pixel_pulses"] / dt) / m["e2e"]["value"],
                     "vs_torchrun_e2e_pageable": ((s2["pixel_pulses"] / dt) / e2e_pageable["value"]
                                                  if e2e_pageable else None),
                     "rank0_block_bit_identical_to_torchrun": bool(np.array_equal(full[a0:a1], m["img"]))}
            del full
            # the same single call from page-locked arrays (what the torchrun `e2e` arm uses)
            try:
                import torch
                full_pin = torch.empty((og.grid_length, og.grid_width), dtype=torch.complex64, pin_memory=True).numpy()
                pin_args = (og, rc_host) + common
                backproject(full_pin, *pin_args, batch=args.batch, devices=devs)
                t0 = time.perf_counter()
                for _ in range(n_il):
                    backproject(full_pin, *pin_args, batch=args.batch, devices=devs)
                dtp = (time.perf_counter() - t0) / n_il
                s3 = last_stats()
                inlib["pinned"] = {"value": s3["pixel_pulses"] / dtp, "unit": UNIT, "ms_per_step": 1e3 * dtp,
                                   "what": "the same single call with page-locked input and output arrays",
                                   "vs_torchrun_e2e": (s3["pixel_pulses"] / dtp) / m["e2e"]["value"],
                                   "rank0_block_bit_identical_to_torchrun":
                                       bool(np.array_equal(full_pin[a0:a1], m["img"]))}
                del full_pin
            except Exception as exc:  # noqa: BLE001 -- an extra arm must not take the line down
                inlib["pinned"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
            release_device_memory()
        comm.host_barrier()

    cpu = None
    if rank == 0 and oracle is not None:
        os.sched_setaffinity(0, all_cpus)  # the CPU baseline uses every host core
        cpu, _ = cpu_parity(comm, sc, m, oracle, cores)
        # one-thread figure on a proportionally smaller sample
        from oracle import tdbp
        tdbp.set_threads(1)
        ppx = pp_rank / max(shape[0] * shape[1], 1)
        cols = max(64, min(og.grid_width, int(6e7 / max(ppx, 1))))
        dt1, _, _, _ = cpu_sample(sc, oracle, 1, cols)
        cpu["value_1_thread"] = ppx * cols / dt1
        cpu["sample_1_thread"] = f"1 azimuth line x {cols} range pixels ({ppx * cols:.3g} pixel*pulses, {dt1:.1f} s)"
        tdbp.set_threads(cores)
    if bound_cpus:
        os.sched_setaffinity(0, set(bound_cpus))  # back next to the GPU for the remaining arms

    # ---- same-GPU comparator: the reference's own CUDA backprojection (reduced harness) ---
    ref_cuda = None
    if rank == 0 and not args.no_ref_cuda and world == 1:
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
                ok = np.isfinite(ref_out) & np.isfinite(out_host)
                ref_cuda = {
                    "value": pp_rank / dt, "unit": UNIT, "ms": 1e3 * dt, "kind": rcu.kind,
                    "what": "isce3::cuda::focus::backproject (cuda/focus/Backproject.cu, unmodified) "
                            "on the same frame, same host buffers, one call, wall time",
                    "e2e_speedup_of_this_repo": m["e2e"]["value"] / (pp_rank / dt),
                    "rel_rms_ours_vs_reference_cuda": float(
                        np.linalg.norm((out_host - ref_out)[ok]) / max(np.linalg.norm(ref_out[ok]), 1e-30)),
                }
                del ref_out
            elif sc.dem.have_raster:
                ref_cuda = {"unavailable": "reduced harness supports constant-height DEMs only"}
            else:
                ref_cuda = {"unavailable": "oracle/_ref/libtdbp_refcuda.so not built"}
        except Exception as exc:  # the comparator must never take the bench down
            ref_cuda = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}

    # ---- slowest supported path: the generic kernel (kernels / shapes the fast path refuses) ---
    worst = None
    if rank == 0 and not args.no_extras:
        try:
            worst = measure_generic(sc, local, args.batch)
            if ref_cuda and "value" in ref_cuda:
                worst["reference_cuda_pp_s_same_box"] = ref_cuda["value"]
        except Exception as exc:
            worst = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}

    main_workload = workload_name(args.config, sc)
    main_img_free = m.pop("img")
    del main_img_free, rc_host, out_host
    m["_keep"] = None
    m["rc_host"] = m["out_host"] = None

    # ---- every other BASELINE.json shape at this N (a few steps each) ----------------------
    configs = {}
    if not args.no_extras and args.scale == 1.0 and args.config == "c2" and not args.taps:
        del sc
        gc.collect()
        release_device_memory()
        for name, cfg, taps_list in (("c1", "c1", [0]), ("c4", "c4", [0]), ("c5", "c5", [8, 16, 32])):
            try:
                scx = make_scene(cfg, 1.0, taps_list[0], **({"n_targets": 81} if cfg == "c4" else {}))
            except Exception as exc:
                configs[name] = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
                continue
            for tp in taps_list:
                key = name if not tp else f"{name}k{tp}"
                try:
                    if tp:
                        from testkit import synth
                        scx.kernel = synth.knab_table_kernel(tp, 0.8, 2048)
                    mx = measure(comm, scx, 2, 1, args.batch, None, pinned=True)
                    entry = None
                    if rank == 0:
                        entry = {"workload": workload_name(cfg, scx), "value": mx["value"], "unit": UNIT,
                                 "ms_per_step": mx["ms_per_step"],
                                 "e2e": {k: mx["e2e"][k] for k in ("value", "ms_per_step", "h2d_bytes_per_step",
                                                                   "d2h_bytes_per_step", "host_memory")},
                                 "e2e_bit_identical_to_resident": mx["e2e"]["max_abs_diff_vs_resident"] == 0.0,
                                 "roofline_frac": mx["achieved_tflops"] / peaks["fp32_tflops"],
                                 "achieved_tflops": mx["achieved_tflops"],
                                 "kernel_ms_per_step": mx["kernel_ms_step"],
                                 "target_solve_ms_per_step": mx["solve_ms_step"],
                                 "fast_variant": mx["stats"]["fast_variant"],
                                 "used_fast_kernel": mx["stats"]["used_fast_kernel"],
                                 "pixel_pulses_per_step": mx["pp_total"]}
                        if oracle is not None:
                            c, parity = cpu_parity(comm, scx, mx, oracle, cores, budget_pp=1.0e8)
                            entry["rel_rms_vs_cpu_reference_on_sample"] = parity
                            entry["cpu_reference_pp_s"] = c["value"]
                            entry["cpu_sample"] = c["sample"]
                        configs[key] = entry
                    mx.clear()
                except Exception as exc:
                    if rank == 0:
                        configs[key] = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
            del scx
            gc.collect()
            release_device_memory()

    if rank != 0:
        comm.close()
        return 0

    achieved_tflops = m["achieved_tflops"]
    peaks_file = {}
    try:
        peaks_file = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    npulse_used = st["pulse_last"] - st["pulse_first"]
    alg_bytes = 8.0 * npulse_used * in_width + 8.0 * shape[0] * shape[1] + 40.0 * shape[0] * shape[1]
    roofline = {
        "bound": "fp32", "kernel": f"accumulate_fast_kernel<{taps}>" if st["used_fast_kernel"] else "accumulate_generic_kernel",
        "fast_variant": st["fast_variant"],
        "achieved": achieved_tflops, "peak": peaks["fp32_tflops"], "unit": "TFLOP/s",
        "frac": achieved_tflops / peaks["fp32_tflops"] if peaks["fp32_tflops"] else None,
        "peak_source": "i3b_measure_peaks FFMA microbenchmark on this device (MEASURED_PEAKS.json has no FP32 peak)",
        "frac_of_nominal_74.5": achieved_tflops / NOMINAL_FP32_TFLOPS,
        "flops_per_pixel_pulse": algorithmic_flops_per_pp(taps), "kernel_ms_per_step": kernel_ms_step,
        "kernel_share_of_step": kernel_ms_step / m["ms_per_step"],
        "pp_per_s_kernel": pp_rank / (kernel_ms_step * 1e-3) if kernel_ms_step > 0 else None,
        "sfu_frac": (3.0 * pp_rank / (kernel_ms_step * 1e-3) / 1e9) / peaks["sfu_gops"] if kernel_ms_step > 0 else None,
        "hbm_algorithmic_gbs": alg_bytes / (kernel_ms_step * 1e-3) / 1e9 if kernel_ms_step > 0 else None,
        "hbm_peak_gbs": peaks_file.get("hbm_gbs"),
        "measured_peaks": peaks, "traffic": None,
    }
    # DRAM traffic of the dominant kernel: one ncu capture of this workload, committed under
    # profiles/ (per launch, like `achieved`); only quoted for the workload it was taken on
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            tr = json.loads((ROOT / "profiles" / name).read_text())
            if world == 1 and args.config == "c2" and args.scale == 1.0 and not args.taps:
                roofline["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
                roofline["traffic_unit"] = "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)"
                roofline["traffic_source"] = tr["source"]
                roofline["algorithmic_bytes_per_launch"] = alg_bytes + 32.0 * shape[0] * shape[1]
            break
        except Exception:
            continue
    # roofline fractions of the wide / narrow kernels measured in `configs`
    by_taps = {f"k{taps}": roofline["frac"]}
    for key, entry in configs.items():
        if key.startswith("c5k") and "roofline_frac" in entry:
            by_taps[key[2:]] = entry["roofline_frac"]
    roofline["frac_by_taps"] = by_taps

    line = {
        "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64+f32", "data": "synthetic",
        "config": {"workload": main_workload, "sharding": "contiguous azimuth blocks of one frame, no exchange",
                   "l2": "inputs (swath + per-pixel tables) far exceed the 126 MB L2",
                   "pixel_pulses_per_step": m["pp_total"], "batch": args.batch,
                   "device_memory": "per-process buffer cache enabled (i3b_set_device_memory_pool(-1); "
                                    "the library's default frees every device allocation before a call returns)",
                   "scene_generation_s": t_gen,
                   "cpu_binding": (f"each rank bound to the {len(bound_cpus)} CPUs of its GPU's NUMA node "
                                   "(NVML cpu affinity) before its host buffers are allocated"
                                   if bound_cpus else "none")},
        "e2e": m["e2e"], "e2e_pageable": e2e_pageable,
        "gpu_launches": int(m["launches"]),
        "clocks": m["clocks"], "roofline": roofline, "roofline_general": roofline_general,
        "cpu_baseline": cpu, "reference_cuda": ref_cuda,
        "worst_supported_path": worst, "inlib_multi_gpu": inlib, "configs": configs,
        "target_solve_ms_per_step": m["solve_ms_step"],
        # CUDA-event time of the two kernels that make up a resident step (rank 0); the
        # difference to ms_per_step (wall, barrier to barrier) is finalisation + host overhead
        "device_ms_per_step": kernel_ms_step + m["solve_ms_step"],
    }
    emit(line)
    comm.close()
    return 0


def measure_resident_only(comm, sc, steps, warmup, batch):
    from isce3_b200.focus import BackprojectPlan
    og = sc.out_geometry
    a0, a1 = block_bounds(og.grid_length, comm.world, comm.rank)
    sub = sc.out_subgrid(a0, a1) if comm.world > 1 else og
    plan = BackprojectPlan(sub, sc.rc, sc.in_geometry, sc.dem, sc.fc, sc.ds, sc.kernel, sc.dry_tropo_model,
                           sc.rdr2geo_params, sc.geo2rdr_params, batch=batch, devices=[comm.local])
    for _ in range(warmup):
        plan.execute()
    comm.barrier()
    kms = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        plan.execute()
        st = plan.stats()
        kms += st["ms_accumulate"]
    comm.barrier()
    el = comm.max(time.perf_counter() - t0)
    pp_total = comm.sum(st["pixel_pulses"])
    plan.close()
    return {"value": pp_total / (el / steps), "kernel_ms_step": kms / steps, "pp_rank": st["pixel_pulses"],
            "fast_variant": st["fast_variant"]}


def measure_generic(sc, local, batch):
    """The generic accumulation kernel (exact kernel evaluation, any kernel type / tap count /
    output spacing) on a block of the same frame: the slowest path a supported call can take."""
    from isce3_b200.focus import BackprojectPlan
    og = sc.out_geometry
    n = min(og.grid_length, 256)
    a0 = max(0, og.grid_length // 2 - n // 2)
    sub = sc.out_subgrid(a0, a0 + n)
    plan = BackprojectPlan(sub, sc.rc, sc.in_geometry, sc.dem, sc.fc, sc.ds, sc.kernel, sc.dry_tropo_model,
                           sc.rdr2geo_params, sc.geo2rdr_params, batch=batch, devices=[local],
                           force_generic=True)
    plan.execute()
    plan.execute()
    st = plan.stats()
    plan.close()
    return {"kernel": "accumulate_generic_kernel (force_generic)", "value": st["pixel_pulses"] / (st["ms_accumulate"] * 1e-3),
            "unit": UNIT, "kernel_ms": st["ms_accumulate"],
            "sample": f"{n} azimuth lines x {og.grid_width} range pixels of the frame ({st['pixel_pulses']:.3g} pixel*pulses)"}


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
    ap.add_argument("--no-extras", action="store_true",
                    help="main workload only: no other configs, pageable / in-library / generic arms")
    args = ap.parse_args()
    args.warmup_ref = min(args.warmup, 1)
    # stdout carries ONE line, the JSON result: anything a library prints there while the bench
    # runs (NCCL announces its version on stdout at communicator creation) goes to stderr
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


_RESULT_FD = None


def emit(line: dict):
    """The bench's one line of output, on the process's original stdout."""
    text = json.dumps(line) + "\n"
    if _RESULT_FD is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, text.encode())


if __name__ == "__main__":
    sys.exit(main())
