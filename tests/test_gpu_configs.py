"""Every BASELINE.json configuration at FULL size, through ``backproject()`` on the GPU,
against the CPU oracle where it is affordable: a chip around every point target (peak phase,
IRF peak location, PSLR, ISLR in both axes) and a few full-width azimuth lines (NaN masks,
relative RMS error, height layer).  Gate: BASELINE.json tolerances (testkit/validate.py).

  c1     2048 pulses x 4096 bins -> 512 x 512, single target
  c2     16384 x 12288 -> 8192 x 8192, flat DEM, tsx, noise
  c4     16384 x 32768 -> 2048 x 8192, raster DEM (biquintic), tsx, input Doppler LUT,
         9 x 9 point-target array
  c5     65536 x 8192 -> 2048 x 2048 airborne geometry, 8 / 16 / 32-tap kernels
"""
import numpy as np
import pytest

from testkit import synth, validate

pytestmark = pytest.mark.gpu


def _assert_pass(rec):
    detail = {k: rec[k] for k in ("rel_rms_block", "worst_chip_rel_rms", "worst_peak_phase_rad",
                                  "worst_az_peak_offset_diff", "worst_rg_peak_offset_diff",
                                  "worst_pslr_diff_db", "worst_islr_diff_db", "height_max_abs_diff_m",
                                  "nan_masks_equal", "used_fast_kernel")}
    assert rec["pass"], detail
    assert rec["used_fast_kernel"] == 1
    assert rec["gpu_status"] == rec["oracle_status"]


def test_c1_full_size():
    rec = validate.run("c1")
    assert rec["targets_checked"] == 1
    _assert_pass(rec)


def test_c2_full_size():
    rec = validate.run("c2")
    assert rec["targets_checked"] == 3 and rec["pixel_pulses"] > 2.5e11
    _assert_pass(rec)


def test_c4_full_size_9x9_targets_raster_dem_doppler_lut_tsx():
    rec = validate.run("c4")
    assert rec["targets_checked"] == 81 and rec["dem"] == "raster" and rec["tropo"] == "tsx"
    _assert_pass(rec)


def test_c5_full_size_kernel_length_sweep():
    sc = synth.make_scene("c5")
    for taps in (8, 16, 32):
        rec = validate.run(f"c5k{taps}", scene=sc, kernel=synth.knab_table_kernel(taps, 0.8, 2048))
        assert rec["taps"] == taps and rec["targets_checked"] == 3
        _assert_pass(rec)
