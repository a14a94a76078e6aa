"""Full-size parity check of the BASELINE.json configurations on a B200 (test / bench tooling).

For each configuration the GPU focuses the whole frame through ``backproject()``; the CPU
oracle (oracle/_ref when present) focuses

  * a chip around EVERY point target  -> peak phase, IRF peak location, PSLR, ISLR (both axes)
  * a few full-width azimuth lines    -> NaN masks, relative RMS error, height layer

Gate (BASELINE.json): rel. RMS <= 1e-4, peak phase <= 1 mrad, peak location <= 0.01 sample,
PSLR / ISLR <= 0.05 dB, height <= 1e-3 m.  Used by tests/test_gpu_configs.py (the driver-run
suite) and scripts/validate_configs.py (markdown table for profiles/).
"""
import json
import time

import numpy as np

from isce3_b200 import core
from isce3_b200.focus import backproject, last_stats
from oracle import tdbp

from . import irf as point_target
from . import synth

CONFIGS = {
    "c1": dict(name="c1"),
    "c2": dict(name="c2"),
    "c4": dict(name="c4", doppler_lut=True, n_targets=81),
    "c5k8": dict(name="c5", taps=8),
    "c5k16": dict(name="c5", taps=16),
    "c5k32": dict(name="c5", taps=32),
}
CHIP = 64


def run(tag, lines_for_block=4, scene=None, kernel=None):
    """Validate one configuration; ``scene`` reuses an already generated scene (the three
    C5 kernel widths share their echoes), ``kernel`` replaces its interpolation kernel."""
    kw = dict(CONFIGS[tag])
    t = time.time()
    sc = scene if scene is not None else synth.make_scene(kw.pop("name"), **kw)
    if kernel is not None:
        sc.kernel = kernel
    t_gen = time.time() - t
    og = sc.out_geometry
    L, W = og.grid_length, og.grid_width
    out = np.empty((L, W), np.complex64)
    h = np.empty((L, W), np.float32)
    backproject(out, *sc.backproject_args(), height=h)  # warm-up (module load, buffer cache)
    t = time.time()
    err = backproject(out, *sc.backproject_args(), height=h)
    t_gpu = time.time() - t
    st = last_stats()
    oracle = tdbp.best()
    common = sc.backproject_args()[1:]
    r = np.asarray(og.slant_range)
    carrier = np.exp(-1j * 4 * np.pi / (core.speed_of_light / sc.fc) * r)

    worst = dict(phase=0.0, az_off=0.0, rg_off=0.0, pslr=0.0, islr=0.0, chip_rel=0.0)
    n_t = 0
    t = time.time()
    for tg in sc.targets:
        i, j = int(round(tg.az_index)), int(round(tg.rg_index))
        a0, c0 = i - CHIP // 2, j - CHIP // 2
        if a0 < 0 or c0 < 0 or a0 + CHIP > L or c0 + CHIP > W:
            continue
        ref = np.zeros((CHIP, CHIP), np.complex64)
        oracle.backproject(ref, sc.out_subgrid(a0, a0 + CHIP, c0, c0 + CHIP), *common)
        gpu = out[a0:a0 + CHIP, c0:c0 + CHIP]
        car = carrier[None, c0:c0 + CHIP]
        ig, _ = point_target.analyze_point_target(gpu * car, CHIP // 2, CHIP // 2, nov=32, chipsize=32)
        ic, _ = point_target.analyze_point_target(ref * car, CHIP // 2, CHIP // 2, nov=32, chipsize=32)
        worst["phase"] = max(worst["phase"], float(abs(np.angle(gpu[CHIP // 2, CHIP // 2] * np.conj(ref[CHIP // 2, CHIP // 2])))))
        worst["az_off"] = max(worst["az_off"], abs(ig["azimuth"]["offset"] - ic["azimuth"]["offset"]))
        worst["rg_off"] = max(worst["rg_off"], abs(ig["range"]["offset"] - ic["range"]["offset"]))
        for ax in ("azimuth", "range"):
            worst["pslr"] = max(worst["pslr"], abs(ig[ax]["PSLR"] - ic[ax]["PSLR"]))
            worst["islr"] = max(worst["islr"], abs(ig[ax]["ISLR"] - ic[ax]["ISLR"]))
        worst["chip_rel"] = max(worst["chip_rel"], float(np.linalg.norm(gpu - ref) / np.linalg.norm(ref)))
        n_t += 1
    t_chips = time.time() - t

    # a few full-width lines around the frame centre
    b0 = max(0, L // 2 - lines_for_block // 2)
    n = min(lines_for_block, L)
    ref = np.zeros((n, W), np.complex64)
    href = np.zeros((n, W), np.float32)
    t = time.time()
    oerr = oracle.backproject(ref, sc.out_subgrid(b0, b0 + n), *common, height=href)
    t_block = time.time() - t
    g = out[b0:b0 + n]
    nan_equal = bool(np.array_equal(np.isnan(g.real), np.isnan(ref.real)))
    m = np.isfinite(ref.real)
    rel = float(np.linalg.norm((g - ref)[m]) / max(np.linalg.norm(ref[m]), 1e-30))
    hdiff = float(np.nanmax(np.abs(h[b0:b0 + n] - href)))
    ok = (rel <= 1e-4 and nan_equal and worst["phase"] <= 1e-3 and worst["az_off"] <= 0.01 and
          worst["rg_off"] <= 0.01 and worst["pslr"] <= 0.05 and worst["islr"] <= 0.05 and hdiff <= 1e-3)
    rec = {
        "config": tag, "pulses": sc.in_geometry.grid_length, "bins": sc.in_geometry.grid_width,
        "out_lines": L, "out_samples": W, "taps": st["taps"], "dem": "raster" if sc.dem.have_raster else "flat",
        "tropo": sc.dry_tropo_model, "oracle": oracle.kind, "gpu_status": bool(err), "oracle_status": bool(oerr),
        "used_fast_kernel": st["used_fast_kernel"], "fast_variant": st["fast_variant"],
        "pixel_pulses": st["pixel_pulses"], "gpu_call_s": t_gpu, "gpu_pp_per_s": st["pixel_pulses"] / t_gpu,
        "targets_checked": n_t, "block_lines": n, "rel_rms_block": rel, "nan_masks_equal": nan_equal,
        "height_max_abs_diff_m": hdiff, "worst_chip_rel_rms": worst["chip_rel"],
        "worst_peak_phase_rad": worst["phase"], "worst_az_peak_offset_diff": worst["az_off"],
        "worst_rg_peak_offset_diff": worst["rg_off"], "worst_pslr_diff_db": worst["pslr"],
        "worst_islr_diff_db": worst["islr"], "pass": bool(ok),
        "seconds": {"scene": t_gen, "oracle_chips": t_chips, "oracle_block": t_block},
    }
    rec = {k: (float(v) if isinstance(v, (np.floating,)) else int(v) if isinstance(v, (np.integer,)) else v)
           for k, v in rec.items()}
    print(json.dumps(rec), flush=True)
    return rec


