"""GPU parity tests: the CUDA path, called through the reference-shaped operator (and so
through the C-ABI), against the CPU oracle on identical seeded inputs.

Gate (BASELINE.json north_star): RMS relative complex error <= 1e-4 over finite pixels with
identical NaN masks, peak phase error <= 1 mrad at each target's peak pixel, point-target
IRF peak location within 0.01 sample and PSLR / ISLR within 0.05 dB, height layer within
1e-3 m.  The oracle is oracle/_ref (the reference's own sources compiled in the build
container) when that library travelled with the repo, else the restated port.
"""
import math

import numpy as np
import pytest

from isce3_b200 import core, focus
from testkit import synth, irf as point_target
from isce3_b200.container import RadarGeometry
from isce3_b200.core import LookSide, LUT2d, OrbitInterpMethod
from isce3_b200.focus import BackprojectPlan, backproject, last_stats

pytestmark = pytest.mark.gpu

RMS_TOL = 1e-4
PHASE_TOL = 1e-3  # rad


def shape_of(sc):
    return (sc.out_geometry.grid_length, sc.out_geometry.grid_width)


def run_gpu(sc, **kw):
    out = np.full(shape_of(sc), 7 + 7j, np.complex64)
    h = np.full(shape_of(sc), -1.0, np.float32)
    err = backproject(out, *sc.backproject_args(), height=h, **kw)
    return err, out, h, last_stats()


def run_cpu(oracle, sc):
    out = np.zeros(shape_of(sc), np.complex64)
    h = np.zeros(shape_of(sc), np.float32)
    err = oracle.backproject(out, *sc.backproject_args(), height=h)
    return err, out, h


def check(gpu, cpu, sc=None, rms_tol=RMS_TOL):
    eg, og, hg, _ = gpu
    ec, oc, hc = cpu
    assert eg == ec
    assert np.array_equal(np.isnan(og.real), np.isnan(oc.real)), "NaN masks differ"
    m = np.isfinite(oc.real)
    if m.any() and np.linalg.norm(oc[m]) > 0:
        rel = np.linalg.norm((og - oc)[m]) / np.linalg.norm(oc[m])
        assert rel <= rms_tol, f"relative RMS error {rel:.3e}"
    assert np.array_equal(np.isnan(hg), np.isnan(hc))
    hm = np.isfinite(hc)
    if hm.any():
        assert np.max(np.abs(hg[hm] - hc[hm])) <= 1e-3
    if sc is not None:
        for tg in sc.targets:
            i, j = int(round(tg.az_index)), int(round(tg.rg_index))
            if 0 <= i < og.shape[0] and 0 <= j < og.shape[1] and abs(oc[i, j]) > 0:
                dphi = abs(np.angle(og[i, j] * np.conj(oc[i, j])))
                assert dphi <= PHASE_TOL, f"peak phase error {dphi:.3e} rad"


@pytest.mark.parametrize("generic", [False, True])
def test_c1_reduced_fast_and_generic(oracle, generic):
    sc = synth.make_scene("c1", pulses=1024, bins=2048, out_lines=40, out_samples=200)
    gpu = run_gpu(sc, force_generic=generic, batch=300)
    assert gpu[3]["used_fast_kernel"] == (0 if generic else 1)
    assert gpu[3]["pixel_pulses"] == 40 * 200 * 1024
    check(gpu, run_cpu(oracle, sc), sc)


def test_c2_like_tsx_noise_full_aperture(oracle):
    sc = synth.make_scene("c2", pulses=6144, bins=1536, out_lines=24, out_samples=256, n_targets=1)
    gpu = run_gpu(sc)
    assert gpu[3]["used_fast_kernel"] == 1
    check(gpu, run_cpu(oracle, sc), sc)


def test_specialised_and_general_fast_kernels_agree(oracle, monkeypatch):
    """The workflow kernel (Knab(9, 1/1.2) on a 2048-point table, focus.py:794-803) runs the
    build-time specialised kernel (coefficients as FFMA2 immediates); I3B_FAST_NO_IMM=1 forces
    the general constant-bank kernel.  Same polynomials, same arithmetic: same image."""
    sc = synth.make_scene("c2", pulses=4096, bins=1024, out_lines=20, out_samples=300, n_targets=1)
    imm = run_gpu(sc)
    assert imm[3]["used_fast_kernel"] == 1 and imm[3]["fast_variant"] == 0
    monkeypatch.setenv("I3B_FAST_NO_IMM", "1")
    bank = run_gpu(sc)
    assert bank[3]["used_fast_kernel"] == 1 and bank[3]["fast_variant"] == -1
    assert np.linalg.norm(imm[1] - bank[1]) <= 2e-6 * np.linalg.norm(bank[1])
    cpu = run_cpu(oracle, sc)
    check(imm, cpu, sc)
    check(bank, cpu, sc)


def test_irf_metrics_match_oracle_nov128_chip64(oracle):
    """The IRF gate exactly as SURVEY.md 8d / the reference test specify it
    (tests/python/extensions/pybind/focus/backproject.py:141-170): nov = 128 on a 64-pixel
    chip; peak location within 0.01 sample, PSLR and ISLR within 0.05 dB, both axes."""
    sc = synth.make_scene("c2", pulses=5120, bins=1024, out_lines=72, out_samples=72, n_targets=1,
                          noise_db=False)
    _, og, _, st = run_gpu(sc)
    _, oc, _ = run_cpu(oracle, sc)
    assert st["used_fast_kernel"] == 1
    r = np.asarray(sc.out_geometry.slant_range)
    carrier = np.exp(-1j * 4 * np.pi / (core.speed_of_light / sc.fc) * r)[None, :]
    tg = sc.targets[0]
    ig, _ = point_target.analyze_point_target(og * carrier, tg.az_index, tg.rg_index, nov=128, chipsize=64)
    ic, _ = point_target.analyze_point_target(oc * carrier, tg.az_index, tg.rg_index, nov=128, chipsize=64)
    for axis in ("azimuth", "range"):
        assert abs(ig[axis]["offset"] - ic[axis]["offset"]) <= 0.01
        assert abs(ig[axis]["PSLR"] - ic[axis]["PSLR"]) <= 0.05
        assert abs(ig[axis]["ISLR"] - ic[axis]["ISLR"]) <= 0.05
        assert abs(ic[axis]["offset"]) < 0.05
    assert abs(ig["phase"] - ic["phase"]) <= PHASE_TOL
    # the reference test's own thresholds: position within resolution / 128, range width <= c / 2B
    dr = sc.out_geometry.radar_grid.range_pixel_spacing
    assert dr * ig["range"]["resolution"] <= core.speed_of_light / (2 * sc.range_bandwidth)
    for axis in ("azimuth", "range"):
        assert abs(ig[axis]["offset"]) <= ig[axis]["resolution"] / 128 + 0.01


def test_irf_metrics_match_oracle(oracle):
    """Point-target IRF of the GPU image vs the oracle image: peak location within 0.01
    sample, PSLR and ISLR within 0.05 dB, in both axes (nov = 32 on a 32-pixel chip)."""
    sc = synth.make_scene("c2", pulses=5120, bins=1024, out_lines=72, out_samples=72, n_targets=1,
                          noise_db=False)
    _, og, _, st = run_gpu(sc)
    _, oc, _ = run_cpu(oracle, sc)
    assert st["used_fast_kernel"] == 1
    r = np.asarray(sc.out_geometry.slant_range)
    carrier = np.exp(-1j * 4 * np.pi / (core.speed_of_light / sc.fc) * r)[None, :]
    tg = sc.targets[0]
    ig, _ = point_target.analyze_point_target(og * carrier, tg.az_index, tg.rg_index, nov=32, chipsize=32)
    ic, _ = point_target.analyze_point_target(oc * carrier, tg.az_index, tg.rg_index, nov=32, chipsize=32)
    for axis in ("azimuth", "range"):
        assert abs(ig[axis]["offset"] - ic[axis]["offset"]) <= 0.01
        assert abs(ig[axis]["PSLR"] - ic[axis]["PSLR"]) <= 0.05
        assert abs(ig[axis]["ISLR"] - ic[axis]["ISLR"]) <= 0.05
        assert abs(ic[axis]["offset"]) < 0.05  # and the target focuses where it was placed
    assert abs(ig["phase"] - ic["phase"]) <= PHASE_TOL
    # reference test thresholds (tests/python/extensions/pybind/focus/backproject.py:161-170)
    dr = sc.out_geometry.radar_grid.range_pixel_spacing
    assert dr * ig["range"]["resolution"] <= core.speed_of_light / (2 * sc.range_bandwidth)


@pytest.mark.parametrize("taps", [3, 4, 5, 6, 7, 10, 11, 12, 13])
def test_every_instantiated_tap_count(oracle, taps):
    """Kernel widths other than the workflow default: each has a fast instantiation."""
    sc = synth.make_scene("c1", pulses=640, bins=1024, out_lines=8, out_samples=136, taps=taps,
                          noise_db=-20.0)
    gpu = run_gpu(sc)
    assert gpu[3]["taps"] == taps and gpu[3]["used_fast_kernel"] == 1
    check(gpu, run_cpu(oracle, sc), sc)


@pytest.mark.parametrize("taps", [8, 16, 32])
def test_airborne_kernel_widths(oracle, taps):
    sc = synth.make_scene("c5", pulses=6144, bins=1536, out_lines=12, out_samples=200, n_targets=1,
                          taps=taps)
    gpu = run_gpu(sc)
    assert gpu[3]["used_fast_kernel"] == 1 and gpu[3]["taps"] == taps
    check(gpu, run_cpu(oracle, sc), sc)


def test_raster_dem_doppler_lut_tsx(oracle):
    """C4-like: EPSG:4326 raster DEM with biquintic sampling, tsx delay, Doppler LUT on the
    input geometry (bilinear, no bounds error)."""
    sc = synth.make_scene("c4", pulses=2048, bins=2048, out_lines=20, out_samples=150, n_targets=4,
                          doppler_lut=True)
    gpu = run_gpu(sc)
    cpu = run_cpu(oracle, sc)
    assert np.nanmax(cpu[2]) - np.nanmin(cpu[2]) > 1.0  # the height layer really varies
    check(gpu, cpu, sc)


@pytest.mark.parametrize("hmax", [2000.0, 2500.0])
def test_raster_dem_relief_up_to_the_verge_of_layover(oracle, hmax):
    """Raster DEMs with 2000 m and 2500 m of relief (largest gradient 0.51 and 0.64 against
    tan(incidence) = 0.73: steep, nothing lays over yet): the height layer and the image are the
    reference's.  (These are the two cases that separated a narrow look bracket from the
    reference's interval when one was tried, see rdr2geo_bracket.)"""
    import dataclasses
    sc = synth.make_scene("c4", pulses=2048, bins=2048, out_lines=24, out_samples=180, n_targets=2)
    d = sc.dem
    half = 0.5 * abs(d.delta_x) * (d.width - 1)
    lon_c, lat_c = d.x_start + d.delta_x * (d.width - 1) / 2, d.y_start + d.delta_y * (d.length - 1) / 2
    dem = synth.synthetic_dem(lon_c, lat_c, half, posting_deg=abs(d.delta_x), hmax=hmax)
    sc = dataclasses.replace(sc, dem=dem)
    gpu = run_gpu(sc)
    cpu = run_cpu(oracle, sc)
    assert np.nanmax(cpu[2]) - np.nanmin(cpu[2]) > 1.0
    check(gpu, cpu, sc)


def _doppler_lut(geom, method, lo, hi, b_error=False, shape=(7, 9), range_margin=2.0e4, wiggle=25.0):
    """Smooth 2-D Doppler LUT covering ``geom``'s grid (and the whole orbit in time)."""
    g, orbit = geom.radar_grid, geom.orbit
    y = np.linspace(orbit.start_time, orbit.end_time, shape[0])
    x = np.linspace(g.starting_range - range_margin,
                    g.starting_range + g.width * g.range_pixel_spacing + range_margin, shape[1])
    yy, xx = np.meshgrid(y, x, indexing="ij")
    data = lo + (hi - lo) * (xx - x[0]) / (x[-1] - x[0]) + \
        wiggle * np.sin(2 * np.pi * (yy - y[0]) / (y[-1] - y[0]))
    return LUT2d(x[0], y[0], x[1] - x[0], y[1] - y[0], data, method, b_error)


@pytest.mark.parametrize("method", ["bilinear", "bicubic", "biquintic"])
def test_output_grid_doppler_lut_with_data(oracle, method):
    """Squinted OUTPUT grid: the output geometry's Doppler LUT holds data (non-zero sin(squint)
    in rdr2geo), sampled with each 2-D method; the input geometry has a different, bilinear LUT,
    so geo2rdr does not simply invert rdr2geo (its root is ~0.3 s from the output line's time)."""
    sc = synth.make_scene("c2", pulses=3072, bins=1536, out_lines=20, out_samples=200, n_targets=1,
                          doppler_lut=True)
    sc.out_geometry = RadarGeometry(sc.out_geometry.radar_grid, sc.out_geometry.orbit,
                                    _doppler_lut(sc.out_geometry, method, -180.0, 220.0))
    gpu = run_gpu(sc)
    cpu = run_cpu(oracle, sc)
    assert gpu[3]["used_fast_kernel"] == 1
    check(gpu, cpu)
    # the squint really moved the targets: compare with the zero-Doppler output grid
    sc0 = synth.make_scene("c2", pulses=3072, bins=1536, out_lines=20, out_samples=200, n_targets=1,
                           doppler_lut=True)
    plain = run_gpu(sc0)
    assert np.linalg.norm(plain[1] - gpu[1]) > 0.5 * np.linalg.norm(plain[1])


def test_biquintic_input_doppler_lut(oracle):
    sc = synth.make_scene("c2", pulses=2048, bins=1024, out_lines=16, out_samples=160, n_targets=1)
    sc.in_geometry = RadarGeometry(sc.in_geometry.radar_grid, sc.in_geometry.orbit,
                                   _doppler_lut(sc.in_geometry, "biquintic", -60.0, 90.0))
    check(run_gpu(sc), run_cpu(oracle, sc))


def test_sinc_doppler_luts(oracle):
    """2-D sinc interpolation (core/Sinc2dInterpolator.cpp: 8 x 8 taps from an 8192-row table)
    of both Doppler LUTs; the reference returns 0 within half a kernel of a LUT's edges, so the
    LUTs are large enough for the solutions to be interior (the root finder's probes still
    visit the zero band, like on the CPU)."""
    sc = synth.make_scene("c2", pulses=2048, bins=1024, out_lines=16, out_samples=160, n_targets=1)
    sc.in_geometry = RadarGeometry(sc.in_geometry.radar_grid, sc.in_geometry.orbit,
                                   _doppler_lut(sc.in_geometry, "sinc", -60.0, 90.0, shape=(24, 24)))
    sc.out_geometry = RadarGeometry(sc.out_geometry.radar_grid, sc.out_geometry.orbit,
                                    _doppler_lut(sc.out_geometry, "sinc", 40.0, -25.0, shape=(24, 24)))
    gpu = run_gpu(sc)
    cpu = run_cpu(oracle, sc)
    check(gpu, cpu)
    assert np.isfinite(gpu[1]).any()


def test_lut_bounds_error_is_reported_and_lookups_are_clamped(oracle):
    """LUT2d(bounds_error=True) that does not cover the output swath: the CPU reference raises
    through its error channel (LUT2d.cpp:143-150), the reference CUDA path silently returns
    ref_value.  Here: the lookup is clamped exactly like bounds_error=False (same image) and the
    call reports the soft code OutOfBoundsLookup (returns True)."""
    sc = synth.make_scene("c2", pulses=2048, bins=1024, out_lines=12, out_samples=160, n_targets=1)
    g = sc.out_geometry.radar_grid
    images = {}
    for b_error in (False, True):
        lut = _doppler_lut(sc.out_geometry, "bilinear", -40.0, 60.0, b_error=b_error,
                           range_margin=-0.25 * g.width * g.range_pixel_spacing)
        assert not lut.contains(g.sensing_start, g.starting_range)
        sc.out_geometry = RadarGeometry(g, sc.out_geometry.orbit, lut)
        images[b_error] = run_gpu(sc)
    assert images[False][0] is False and images[True][0] is True
    np.testing.assert_array_equal(images[True][1], images[False][1])
    cpu = run_cpu(oracle, sc)  # (the oracle's LUT shim clamps and does not raise)
    check((cpu[0],) + images[True][1:], cpu)


def test_input_geometry_with_its_own_orbit_and_doppler(oracle):
    """Input and output geometry do NOT share orbit object, sampling or Doppler model: the
    input orbit is the same trajectory resampled every 5 s (Legendre), the output one every
    10 s (Hermite), both LUTs hold different data.  geo2rdr's shortcut ("the output line time
    is the root") must fail its own test and fall back to the bracketed search."""
    sc = synth.make_scene("c2", pulses=3072, bins=1536, out_lines=16, out_samples=180, n_targets=1)
    o = sc.out_geometry.orbit
    t = np.arange(o.start_time, o.end_time + 1e-9, 5.0)
    pos, vel = synth.interpolate_orbit_many(o, t)
    own = core.Orbit.from_arrays(float(t[0]), 5.0, pos, vel, o.reference_epoch)
    own.interp_method = OrbitInterpMethod.LEGENDRE
    sc.in_geometry = RadarGeometry(sc.in_geometry.radar_grid, own,
                                   _doppler_lut(sc.in_geometry, "bilinear", 120.0, 260.0))
    sc.out_geometry = RadarGeometry(sc.out_geometry.radar_grid, o,
                                    _doppler_lut(sc.out_geometry, "bilinear", -150.0, -30.0))
    gpu = run_gpu(sc)
    check(gpu, run_cpu(oracle, sc))
    assert gpu[3]["pixel_pulses"] > 0


@pytest.mark.parametrize("dem_epsg", ["utm", 3413, 6933])
def test_raster_dem_in_projected_coordinates(oracle, dem_epsg):
    """Raster DEM in UTM / polar stereographic / EASE-2 coordinates
    (DEMInterpolator::interpolateLonLat -> createProj(epsg)->forward,
    geometry/DEMInterpolator.cpp:592-611, core/Projections.cpp:373-402)."""
    sc = synth.make_scene("c4", pulses=1024, bins=1024, out_lines=12, out_samples=140, n_targets=1,
                          dem_epsg=dem_epsg)
    assert sc.dem.have_raster and sc.dem.epsg_code != 4326
    gpu = run_gpu(sc)
    cpu = run_cpu(oracle, sc)
    assert np.nanmax(cpu[2]) - np.nanmin(cpu[2]) > 1.0  # the height layer really varies
    check(gpu, cpu, sc)


@pytest.mark.parametrize("method", ["bilinear", "bicubic", "nearest", "sinc"])
def test_other_dem_interpolators(oracle, method):
    sc = synth.make_scene("c4", pulses=512, bins=1024, out_lines=8, out_samples=64, n_targets=1)
    sc.dem.interp_method = core.parse_interp_method(method)
    check(run_gpu(sc), run_cpu(oracle, sc), sc)


def test_right_looking_legendre_orbit(oracle):
    sc = synth.make_scene("c1", pulses=768, bins=1024, out_lines=16, out_samples=130,
                          look_side=LookSide.Right)
    for g in (sc.in_geometry, sc.out_geometry):
        g.orbit.interp_method = OrbitInterpMethod.LEGENDRE
    check(run_gpu(sc), run_cpu(oracle, sc), sc)


def test_general_output_spacing(oracle):
    """Output range spacing and PRF different from the input's: exercises the independent
    window path of the fast kernel (pixel pairs that do not share a sample window)."""
    sc = synth.make_scene("c1", pulses=1024, bins=2048, out_lines=24, out_samples=140,
                          out_range_spacing_ratio=1.37, out_prf_ratio=0.71)
    gpu = run_gpu(sc)
    assert gpu[3]["used_fast_kernel"] == 1
    check(gpu, run_cpu(oracle, sc), sc)
    sc = synth.make_scene("c1", pulses=1024, bins=2048, out_lines=24, out_samples=140,
                          out_range_spacing_ratio=0.5)
    check(run_gpu(sc), run_cpu(oracle, sc), sc)


def test_swath_edges_are_zero_padded_like_the_cpu(oracle):
    """Output grid wider than the input swath: windows hanging over the first/last range
    bin use zeros (CPU semantics, core/detail/Interp1d.h:54-80), pixels far outside sum to 0."""
    sc = synth.make_scene("c1", pulses=512, bins=256, out_lines=8, out_samples=300, noise_db=-20.0)
    gpu = run_gpu(sc)
    cpu = run_cpu(oracle, sc)
    check(gpu, cpu, sc)
    assert np.all(cpu[1][:, :10] == 0) and np.all(gpu[1][:, :10] == 0)
    assert np.any(cpu[1][:, 20:30] != 0)


def test_clipped_apertures_and_chunking_invariance(oracle):
    """Output lines beyond the pulse span have clipped apertures (kstart/kstop clamps,
    Backproject.cpp:190-193); results do not depend on the H2D slab size."""
    sc = synth.make_scene("c2", pulses=3000, bins=1024, out_lines=5000, out_samples=4, n_targets=1,
                          out_prf_ratio=1.0)
    # keep every 250th line only: rebuild the output grid with a lower PRF
    g = sc.out_geometry.radar_grid
    g.prf = g.prf / 250.0
    g.length = 20
    sc.out_geometry = RadarGeometry(g, sc.out_geometry.orbit, LUT2d())
    cpu = run_cpu(oracle, sc)
    a = run_gpu(sc, batch=97)
    b = run_gpu(sc, batch=100000)
    check(a, cpu)
    check(b, cpu)
    np.testing.assert_array_equal(a[1], b[1])  # bit-identical whatever the slab size
    pp = a[3]["pixel_pulses"]
    assert 0 < pp < 20 * 4 * 3000


def test_non_finite_samples_only_poison_the_apertures_that_hold_them(oracle):
    """A NaN pulse near the start of the aperture of the first output lines: those lines are
    NaN (they integrate it), later lines -- whose apertures start after it but which share
    pulse tiles with the first ones on the GPU -- must stay finite, exactly like the CPU."""
    sc = synth.make_scene("c2", pulses=6144, bins=1024, out_lines=16, out_samples=130, n_targets=1)
    probe = np.zeros(shape_of(sc), np.complex64)
    oracle.backproject(probe, *sc.backproject_args())
    assert np.isfinite(probe).all()
    # find the first pulse used by line 0 through the oracle's own NaN response
    lo, hi = 0, sc.rc.shape[0] - 1
    base = sc.rc.copy()
    while lo < hi:  # bisection on "line 0 integrates pulse k"
        mid = (lo + hi) // 2
        sc.rc = base.copy()
        sc.rc[:mid + 1] = np.nan
        o = np.zeros(shape_of(sc), np.complex64)
        oracle.backproject(o, *sc.backproject_args())
        if np.isnan(o[0]).any():
            hi = mid
        else:
            lo = mid + 1
    kstart0 = lo
    sc.rc = base.copy()
    sc.rc[kstart0 + 6] = np.nan
    gpu = run_gpu(sc)
    cpu = run_cpu(oracle, sc)
    nan_rows = np.isnan(cpu[1]).all(axis=1)
    assert nan_rows.any() and not nan_rows.all()
    check(gpu, cpu)


def test_fused_output_encoding_matches_the_workflow_writer():
    """range_cor / mantissa_nbits extensions: same result as the host post-processing of
    BackgroundWriter.write (nisar/workflows/focus.py:899-925): z *= range_cor[None, :], then
    truncate_mantissa(z, nbits) (isce3/core/types.py:116-171)."""
    sc = synth.make_scene("c2", pulses=2048, bins=1024, out_lines=16, out_samples=200, n_targets=1)
    _, plain, _, _ = run_gpu(sc)
    rng = np.random.default_rng(11)
    cor = (2.5 * np.exp(2j * np.pi * rng.uniform(size=200))).astype(np.complex64)
    out = np.zeros(shape_of(sc), np.complex64)
    backproject(out, *sc.backproject_args(), range_cor=cor)
    want = plain * cor[None, :]
    assert np.max(np.abs(out - want)) <= 4e-7 * np.max(np.abs(want))
    nbits = 10
    backproject(out, *sc.backproject_args(), range_cor=cor, mantissa_nbits=nbits)
    mask = np.uint32((0xFFFFFFFF << (23 - nbits)) & 0xFFFFFFFF)
    bits = out.view(np.uint32)
    assert np.all(bits & ~mask == 0), "low mantissa bits must be zero"
    trunc = want.copy()
    trunc.view(np.uint32)[...] &= mask
    # identical except where the fused multiply rounded the last kept bit differently
    assert np.max(np.abs(out - trunc)) <= 2.0 ** -nbits * np.max(np.abs(want))
    assert np.mean(out == trunc) > 0.99
    with pytest.raises(focus.InvalidArgument):
        backproject(out, *sc.backproject_args(), range_cor=cor[:-1])
    with pytest.raises(focus.InvalidArgument):
        backproject(out, *sc.backproject_args(), mantissa_nbits=24)


def test_one_launch_per_slab_equals_one_launch(oracle):
    """Many accumulation launches over narrow pulse slabs (I3B_LAUNCH_PER_SLAB=1 is read at
    first use, so this runs in a fresh process): every pixel still integrates exactly its
    own aperture -- per-tile pulse spans, aperture-edge tiles and cross-launch FP64 sums --
    for an output width that is not a multiple of the warp or tile size."""
    import subprocess
    import sys
    code = """
import numpy as np, sys
from testkit import synth
from isce3_b200.focus import backproject, last_stats
sc = synth.make_scene("c2", pulses=3072, bins=1024, out_lines=37, out_samples=203, n_targets=1)
shape = (37, 203)
a = np.zeros(shape, np.complex64); b = np.zeros(shape, np.complex64)
backproject(a, *sc.backproject_args(), batch=53)
na = last_stats()["accumulate_launches"]
backproject(b, *sc.backproject_args(), batch=100000)
nb = last_stats()["accumulate_launches"]
# pulse tiles / geometry segments sit on absolute pulse indices and launches end on tile
# boundaries: the two images are bit-identical
print("RESULT", na, nb, float(np.max(np.abs(a - b))))
assert na > 20 and nb == 1, (na, nb)
assert np.array_equal(a, b)
"""
    import os
    env = dict(os.environ, I3B_LAUNCH_PER_SLAB="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env,
                       cwd=str(__import__("pathlib").Path(__file__).resolve().parent.parent))
    assert r.returncode == 0, r.stdout + r.stderr


def test_failed_pixels_are_nan_and_flagged(oracle):
    """geo2rdr bracket that excludes part of the image: those pixels are (NaN, NaN), the call
    returns True (FailedToConverge); rdr2geo failure also NaNs the height layer."""
    sc = synth.make_scene("c1", pulses=512, bins=1024, out_lines=12, out_samples=260)
    t = sc.out_geometry.sensing_time
    sc.geo2rdr_params = {"time_start": float(t[5]) + 1e-4, "time_end": None}
    gpu = run_gpu(sc)
    cpu = run_cpu(oracle, sc)
    assert gpu[0] is True and cpu[0] is True
    assert np.isnan(gpu[1][:5]).all() and np.isfinite(gpu[1][7:]).all()
    assert np.isfinite(gpu[2]).all()  # rdr2geo converged everywhere: heights are valid
    check(gpu, cpu)
    sc.geo2rdr_params = {}
    sc.rdr2geo_params = {"look_min": 0.0, "look_max": 0.3}  # target look angle is ~0.6 rad
    gpu = run_gpu(sc)
    cpu = run_cpu(oracle, sc)
    assert gpu[0] is True and np.isnan(gpu[1]).all() and np.isnan(gpu[2]).all()
    check(gpu, cpu)


@pytest.mark.parametrize("kernel", ["cheby", "knab", "linear", "bartlett", "tab5"])
def test_other_kernel_types(oracle, kernel):
    sc = synth.make_scene("c1", pulses=512, bins=1024, out_lines=8, out_samples=130, noise_db=-20.0)
    k = {"cheby": lambda: core.ChebyKernelF32(core.KnabKernel(9.0, 0.8), 16),
         "knab": lambda: core.KnabKernelF32(9.0, 0.8),
         "linear": core.LinearKernelF32,
         "bartlett": lambda: core.BartlettKernelF32(5.0),
         "tab5": lambda: core.TabulatedKernelF32(core.KnabKernel(5.0, 0.8), 512)}[kernel]()
    sc.kernel = k
    gpu = run_gpu(sc)
    # piecewise-linear kernels have no per-tap polynomial form; the others take the fast
    # kernel whenever the host-side fit meets its residual bound (tab5: a coarse 512-point table)
    from isce3_b200 import _capi
    import ctypes
    fl = _capi.Flattened()
    fit = _capi.TapPolyFit()
    _capi.load_library().i3b_fit_tap_polynomials(ctypes.byref(_capi.flatten_kernel(k, fl)), ctypes.byref(fit))
    assert fit.supported == (0 if kernel in ("linear", "bartlett") else fit.supported)
    assert gpu[3]["used_fast_kernel"] == fit.supported
    if kernel in ("cheby", "knab"):
        assert fit.supported == 1
    # Cheby/Knab<float> kernels are evaluated in float by the reference: allow their rounding
    check(gpu, run_cpu(oracle, sc), sc, rms_tol=2e-4 if kernel == "knab" else RMS_TOL)


def test_error_paths_raise_like_the_reference():
    sc = synth.make_scene("c1", pulses=128, bins=256, out_lines=4, out_samples=8)
    out = np.zeros((4, 8), np.complex64)
    other = core.DateTime(1999, 1, 1)
    g = sc.out_geometry.radar_grid.copy()
    g.ref_epoch = other
    orb = core.Orbit.from_arrays(sc.out_geometry.orbit.time.first, sc.out_geometry.orbit.time.spacing,
                                 sc.out_geometry.orbit.position, sc.out_geometry.orbit.velocity, other)
    args = list(sc.backproject_args())
    args[0] = RadarGeometry(g, orb, LUT2d())
    with pytest.raises(RuntimeError, match="reference epoch"):
        backproject(out, *args)
    # output line time outside the orbit: the CPU reference throws OutOfRange
    g2 = sc.out_geometry.radar_grid.copy()
    g2.sensing_start = sc.out_geometry.orbit.end_time + 100.0
    args = list(sc.backproject_args())
    args[0] = RadarGeometry(g2, sc.out_geometry.orbit, LUT2d())
    with pytest.raises(IndexError):
        backproject(out, *args)


def test_plan_resident_matches_one_shot_and_is_repeatable(oracle):
    sc = synth.make_scene("c2", pulses=2048, bins=1024, out_lines=16, out_samples=300, n_targets=1)
    one = run_gpu(sc)
    with BackprojectPlan(*sc.backproject_args()) as plan:
        plan.execute()
        a = plan.download()
        plan.execute()
        b = plan.download()
        st = plan.stats()
    np.testing.assert_array_equal(a, b)
    # the one-shot call integrates slab by slab: same tiles, same order of FP64 additions
    np.testing.assert_array_equal(a, one[1])
    assert st["accumulate_launches"] == 1 and st["used_fast_kernel"] == 1
    check((one[0], a, one[2], st), run_cpu(oracle, sc), sc)


def test_azimuth_sharding_over_device_list(oracle):
    """devices=[0, 0, 0]: three azimuth-block shards (host threads) on the same GPU give the
    same image as one shard -- the multi-GPU path without needing several GPUs."""
    sc = synth.make_scene("c2", pulses=2048, bins=1024, out_lines=23, out_samples=140, n_targets=1)
    one = run_gpu(sc)
    many = run_gpu(sc, devices=[0, 0, 0])
    assert many[3]["n_devices"] == 3
    assert many[3]["pixel_pulses"] == one[3]["pixel_pulses"]
    np.testing.assert_array_equal(many[1], one[1])
    np.testing.assert_array_equal(many[2], one[2])
    check(many, run_cpu(oracle, sc), sc)


def test_image_is_bit_reproducible():
    """Fixed inputs -> one image, bit for bit: whatever the H2D slab size (`batch`), however
    the host link's timing cut the pulses into launches (repeated one-shot calls), resident
    plan or one-shot call, one shard or several (the reference CUDA path is deterministic
    per batch, cuda/focus/Backproject.cu:663-688)."""
    sc = synth.make_scene("c2", pulses=4096, bins=1024, out_lines=41, out_samples=300, n_targets=2)
    ref = run_gpu(sc, batch=1024)
    assert ref[3]["used_fast_kernel"] == 1
    for kw in (dict(batch=53), dict(batch=100000), dict(batch=1024), dict(batch=300, devices=[0, 0, 0]),
               dict(batch=17, devices=[0, 0])):
        other = run_gpu(sc, **kw)
        np.testing.assert_array_equal(other[1], ref[1], err_msg=str(kw))
        np.testing.assert_array_equal(other[2], ref[2], err_msg=str(kw))
    with BackprojectPlan(*sc.backproject_args()) as plan:
        plan.execute()
        np.testing.assert_array_equal(plan.download(), ref[1])
    # general output spacing (independent windows) and a wide kernel (chunked MAC)
    for kw in (dict(name="c1", pulses=1024, bins=2048, out_lines=24, out_samples=140,
                    out_range_spacing_ratio=1.37, out_prf_ratio=0.71),
               dict(name="c5", pulses=4096, bins=1536, out_lines=12, out_samples=200, n_targets=1, taps=16)):
        kw = dict(kw)
        sc = synth.make_scene(kw.pop("name"), **kw)
        a = run_gpu(sc, batch=100000)
        b = run_gpu(sc, batch=61, devices=[0, 0])
        np.testing.assert_array_equal(a[1], b[1])


def test_azimuth_sharding_over_two_gpus(oracle):
    """devices=[0, 1]: one azimuth block per GPU, each fed only the pulses its block needs,
    results gathered on the host with no inter-GPU exchange (skipped on a 1-GPU box)."""
    from isce3_b200 import _capi
    if _capi.load_library().i3b_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sc = synth.make_scene("c2", pulses=3072, bins=1024, out_lines=50, out_samples=260, n_targets=1)
    one = run_gpu(sc)
    two = run_gpu(sc, devices=[0, 1])
    assert two[3]["n_devices"] == 2
    assert two[3]["pixel_pulses"] == one[3]["pixel_pulses"]
    np.testing.assert_array_equal(two[1], one[1])
    np.testing.assert_array_equal(two[2], one[2])
    check(two, run_cpu(oracle, sc), sc)


def test_callers_current_device_is_left_alone():
    """The library selects devices internally (shards, plan destruction, buffer cache) but
    hands the calling thread back on the device it came with (the reference never touches
    the current device; needs 2 GPUs to be observable)."""
    from isce3_b200 import _capi
    lib = _capi.load_library()
    if lib.i3b_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sc = synth.make_scene("c2", pulses=1024, bins=1024, out_lines=16, out_samples=130, n_targets=1)
    before = lib.i3b_current_device()
    run_gpu(sc, devices=[1])
    assert lib.i3b_current_device() == before
    run_gpu(sc, devices=[0, 1])
    assert lib.i3b_current_device() == before
    plan = BackprojectPlan(*sc.backproject_args(), devices=[1])
    plan.execute()
    assert lib.i3b_current_device() == before
    plan.close()
    assert lib.i3b_current_device() == before
    lib.i3b_release_device_memory()
    assert lib.i3b_current_device() == before


def test_plan_download_needs_an_execute():
    sc = synth.make_scene("c2", pulses=512, bins=512, out_lines=8, out_samples=130, n_targets=1)
    with BackprojectPlan(*sc.backproject_args()) as plan:
        with pytest.raises(RuntimeError, match="before any"):
            plan.download()


def test_reference_cuda_comparator_agrees(oracle):
    """The reference's own CUDA backprojection (oracle/_ref/libtdbp_refcuda.so, reduced
    harness) on the same inputs: it is the same-GPU baseline of bench.py, so check that it
    really computes the same image as the CPU reference and as this repo's kernels."""
    from oracle import tdbp
    if not tdbp.have_ref_cuda():
        pytest.skip("oracle/_ref/libtdbp_refcuda.so not built")
    sc = synth.make_scene("c2", pulses=4096, bins=1024, out_lines=24, out_samples=200, n_targets=1)
    ref = np.zeros(shape_of(sc), np.complex64)
    href = np.zeros(shape_of(sc), np.float32)
    assert tdbp.ref_cuda().backproject(ref, *sc.backproject_args(), height=href) is False
    gpu = run_gpu(sc)
    cpu = run_cpu(oracle, sc)
    assert np.linalg.norm(ref - cpu[1]) <= 1e-5 * np.linalg.norm(cpu[1])
    assert np.linalg.norm(ref - gpu[1]) <= RMS_TOL * np.linalg.norm(ref)
    assert np.max(np.abs(href - cpu[2])) <= 1e-3


def test_isce3_side_adapter_round_trip(oracle):
    """integration/isce3/cuda/focus/BackprojectB200.cpp (the file a maintainer adds to isce3),
    compiled against the reference's headers: isce3 objects -> adapter -> i3b_backproject.
    Same image as this repo's own host side, raster DEM / Doppler LUT / Cheby kernel included,
    and the reference's exceptions come back for bad arguments."""
    from oracle import tdbp
    if not tdbp.have_adapter():
        pytest.skip("oracle/_ref/libtdbp_adapter.so not built")
    ad = tdbp.adapter()
    for kw in (dict(name="c2", pulses=2048, bins=1024, out_lines=12, out_samples=150, n_targets=1),
               dict(name="c4", pulses=1024, bins=1024, out_lines=10, out_samples=130, n_targets=1,
                    doppler_lut=True)):
        kw = dict(kw)
        sc = synth.make_scene(kw.pop("name"), **kw)
        direct = run_gpu(sc)
        out = np.zeros(shape_of(sc), np.complex64)
        h = np.zeros(shape_of(sc), np.float32)
        err = ad.backproject(out, *sc.backproject_args(), height=h)
        assert err == direct[0]
        np.testing.assert_array_equal(out, direct[1])
        np.testing.assert_array_equal(h, direct[2])
    # a Kernel<float> subclass the adapter does not know (the test wrapper replays Chebyshev
    # coefficients through its own subclass) is refused like the reference CUDA path refuses
    # it (cuda/focus/Backproject.cu:750-752)
    sc.kernel = core.ChebyKernelF32(core.KnabKernel(9.0, 0.8), 16)
    with pytest.raises(RuntimeError, match="not implemented"):
        ad.backproject(out, *sc.backproject_args())
    sc.kernel = core.KnabKernelF32(9.0, 0.8)
    direct = run_gpu(sc)
    ad.backproject(out, *sc.backproject_args())
    np.testing.assert_array_equal(out, direct[1])


def test_full_c1_properties():
    """BASELINE.json configs[0] at full size (2048 x 4096 -> 512 x 512), checked through
    size-independent properties: the target focuses at its pixel with gain = #pulses,
    linearity in the input data, and the fast kernel agrees with the generic (FP64) one."""
    sc = synth.make_scene("c1")
    e, out, h, st = run_gpu(sc)
    assert e is False and st["used_fast_kernel"] == 1
    assert st["pixel_pulses"] == 512 * 512 * 2048
    tg = sc.targets[0]
    peak = np.unravel_index(np.argmax(np.abs(out)), out.shape)
    assert peak == (int(tg.az_index), int(tg.rg_index))
    assert abs(abs(out[peak]) - 2048) < 0.02 * 2048
    assert np.allclose(h, 0.0, atol=1e-3)
    _, gen, _, st2 = run_gpu(sc, force_generic=True)
    assert st2["used_fast_kernel"] == 0
    assert np.linalg.norm(out - gen) <= RMS_TOL * np.linalg.norm(gen)
    rng = np.random.default_rng(5)
    other = (rng.standard_normal(sc.rc.shape, dtype=np.float32) * 0.05).astype(np.complex64)
    base = sc.rc
    sc.rc = other
    _, o2, _, _ = run_gpu(sc)
    sc.rc = (2.0 * base - 3.0 * other).astype(np.complex64)
    _, o3, _, _ = run_gpu(sc)
    assert np.linalg.norm(o3 - (2.0 * out - 3.0 * o2)) <= 2e-5 * np.linalg.norm(o3)


def test_empty_and_tiny_grids():
    sc = synth.make_scene("c1", pulses=256, bins=512, out_lines=1, out_samples=1)
    e, out, _, st = run_gpu(sc)
    assert e is False and np.isfinite(out).all() and st["pixel_pulses"] == 256
    g = sc.out_geometry.radar_grid.copy()
    g.length = 0
    sc.out_geometry = RadarGeometry(g, sc.out_geometry.orbit, LUT2d())
    out = np.zeros((0, 1), np.complex64)
    assert backproject(out, *sc.backproject_args()) is False


def test_blocks_api_equals_per_block_calls(oracle):
    """i3b_blocks_*: the swath and every table uploaded once, output blocks focused from there
    (the workflow's per-block loop, focus.py:1988-2007).  Full-width blocks cut on tile rows
    reproduce the whole-image call bit for bit; arbitrary blocks (range splits, ragged sizes)
    equal their own per-block backproject() call bit for bit, and the range_cor phasors follow
    the block's columns."""
    from isce3_b200.focus import BlockFocuser
    sc = synth.make_scene("c2", pulses=3072, bins=1024, out_lines=44, out_samples=300, n_targets=2)
    grid = sc.out_geometry.radar_grid
    whole = run_gpu(sc)
    common = sc.backproject_args()
    with BlockFocuser(*common, devices=[0, 0]) as bf:
        # (a) full-width blocks of 12 lines (a multiple of the 4-line tile rows)
        blocks, image = [], np.zeros(shape_of(sc), np.complex64)
        heights = np.zeros(shape_of(sc), np.float32)
        for a0 in range(0, 44, 12):
            a1 = min(a0 + 12, 44)
            blocks.append((grid[a0:a1, :], image[a0:a1], heights[a0:a1]))
        assert bf.run(blocks) is False
        st = bf.stats()
        assert st["pixel_pulses"] == whole[3]["pixel_pulses"] and st["used_fast_kernel"] == 1
        np.testing.assert_array_equal(image, whole[1])
        np.testing.assert_array_equal(heights, whole[2])
        # (b) ragged blocks with range splits: each equals its own one-shot call
        cuts = [(0, 17, 0, 130), (0, 17, 130, 300), (17, 44, 0, 77), (17, 44, 77, 300)]
        outs = [np.zeros((a1 - a0, r1 - r0), np.complex64) for a0, a1, r0, r1 in cuts]
        assert bf.run([(grid[a0:a1, r0:r1], o) for (a0, a1, r0, r1), o in zip(cuts, outs)]) is False
        for (a0, a1, r0, r1), o in zip(cuts, outs):
            single = np.zeros_like(o)
            backproject(single, sc.out_subgrid(a0, a1, r0, r1), *common[1:])
            np.testing.assert_array_equal(o, single)
            assert np.linalg.norm(o - whole[1][a0:a1, r0:r1]) <= 1e-5 * np.linalg.norm(o)
    cpu = run_cpu(oracle, sc)
    check((whole[0], image, heights, whole[3]), cpu, sc)
    # (c) per-column output phasors follow the block's columns
    rng = np.random.default_rng(3)
    cor = np.exp(2j * np.pi * rng.uniform(size=300)).astype(np.complex64)
    with BlockFocuser(*common, range_cor=cor) as bf:
        o = np.zeros((10, 100), np.complex64)
        bf.run([(grid[8:18, 150:250], o)])
    want = np.zeros((10, 100), np.complex64)
    backproject(want, sc.out_subgrid(8, 18, 150, 250), *common[1:], range_cor=cor[150:250])
    np.testing.assert_array_equal(o, want)


def test_non_uniform_pulse_times(oracles):
    """ABI 3 extension ``pulse_times`` (dithered PRF without the resampling pre-pass of
    nisar/workflows/focus.py:973-1061).  Pinned three ways: (i) explicit times that ARE the
    uniform grid give the bit-identical image of a call without them; (ii) on a dithered pulse
    train the GPU equals the restated port oracle run with the same times (the reference API
    cannot express them); (iii) physics: the point target focuses at its pixel with gain =
    #pulses and the IRF of the uniform acquisition, while ignoring the times defocuses it."""
    port, _ = oracles
    kw = dict(pulses=4096, bins=1024, out_lines=40, out_samples=72, n_targets=1, noise_db=False)
    uni = synth.make_scene("c2", **kw)
    g = uni.in_geometry.radar_grid
    t_uniform = g.sensing_start + np.arange(g.length) / g.prf
    plain = run_gpu(uni)
    same = run_gpu(uni, pulse_times=t_uniform)
    np.testing.assert_array_equal(same[1], plain[1])
    assert same[3]["used_fast_kernel"] == 1

    dit = synth.make_scene("c2", prf_dither=2.0, **kw)
    assert np.max(np.abs(dit.pulse_times - t_uniform)) * g.prf > 1.5
    gpu = run_gpu(dit, pulse_times=dit.pulse_times)
    ref = np.zeros(shape_of(dit), np.complex64)
    href = np.zeros(shape_of(dit), np.float32)
    err = port.backproject(ref, *dit.backproject_args(), height=href, pulse_times=dit.pulse_times)
    check(gpu, (err, ref, href), dit)
    tg = dit.targets[0]
    i, j = int(tg.az_index), int(tg.rg_index)
    n_int = gpu[3]["pixel_pulses"] / ref.size
    assert np.unravel_index(np.argmax(np.abs(gpu[1])), ref.shape) == (i, j)
    assert abs(abs(gpu[1][i, j]) - n_int) < 0.02 * n_int
    ignored = run_gpu(dit)  # same echoes, pulse times not given: defocused
    assert abs(ignored[1][i, j]) < 0.6 * abs(gpu[1][i, j])
    r = np.asarray(dit.out_geometry.slant_range)
    carrier = np.exp(-1j * 4 * np.pi / (core.speed_of_light / dit.fc) * r)[None, :]
    a, _ = point_target.analyze_point_target(gpu[1] * carrier, i, j, nov=32, chipsize=32)
    b, _ = point_target.analyze_point_target(plain[1] * carrier, i, j, nov=32, chipsize=32)
    for axis in ("azimuth", "range"):
        assert abs(a[axis]["offset"] - b[axis]["offset"]) <= 0.01
        assert abs(a[axis]["resolution"] - b[axis]["resolution"]) <= 0.02 * b[axis]["resolution"]
    # argument checks
    out = np.zeros(shape_of(dit), np.complex64)
    with pytest.raises(focus.InvalidArgument):
        backproject(out, *dit.backproject_args(), pulse_times=dit.pulse_times[:-1])
    bad = dit.pulse_times.copy()
    bad[100] = bad[99]
    with pytest.raises(focus.InvalidArgument, match="strictly increasing"):
        backproject(out, *dit.backproject_args(), pulse_times=bad)


def test_compiled_pybind11_binding_gives_the_same_image(oracle):
    """isce3_b200.ext._backproject -- the compiled binding with the reference's call shape
    (pybind_isce3/cuda/focus/Backproject.cpp:25-117) -- against the ctypes route and the oracle,
    raster DEM / Doppler LUT / Chebyshev kernel included; geometry failures come back as True."""
    import isce3_b200.ext.isce3 as isce
    for kw in (dict(name="c2", pulses=2048, bins=1024, out_lines=12, out_samples=150, n_targets=1),
               dict(name="c4", pulses=1024, bins=1024, out_lines=10, out_samples=130, n_targets=1,
                    doppler_lut=True)):
        kw = dict(kw)
        sc = synth.make_scene(kw.pop("name"), **kw)
        direct = run_gpu(sc)
        out = np.zeros(shape_of(sc), np.complex64)
        h = np.zeros(shape_of(sc), np.float32)
        err = isce.cuda.focus.backproject(out, *sc.backproject_args(), height=h)
        assert err is direct[0]
        np.testing.assert_array_equal(out, direct[1])
        np.testing.assert_array_equal(h, direct[2])
    sc.kernel = core.ChebyKernelF32(core.KnabKernel(9.0, 0.8), 16)
    direct = run_gpu(sc)
    isce.cuda.focus.backproject(out, *sc.backproject_args())
    np.testing.assert_array_equal(out, direct[1])
    check((direct[0], out, direct[2], direct[3]), run_cpu(oracle, sc), sc)
    sc.rdr2geo_params = {"look_min": 0.0, "look_max": 0.3}
    assert isce.cuda.focus.backproject(out, *sc.backproject_args()) is True
    assert np.isnan(out).all()


def test_device_memory_is_returned_by_default_and_kept_on_request():
    """Ownership contract of the reference (all device memory freed before return,
    cuda/focus/Backproject.cu:691-695) is the DEFAULT; the buffer cache is opt-in."""
    import ctypes
    from isce3_b200 import _capi
    from isce3_b200.focus import keep_device_memory, release_device_memory
    rt = None
    for name in ("libcudart.so", "libcudart.so.12"):
        try:
            rt = ctypes.CDLL(name)
            break
        except OSError:
            continue
    if rt is None:
        pytest.skip("no libcudart to query free memory with")

    def free_bytes():
        f, t = ctypes.c_size_t(), ctypes.c_size_t()
        assert rt.cudaMemGetInfo(ctypes.byref(f), ctypes.byref(t)) == 0
        return f.value
    sc = synth.make_scene("c2", pulses=2048, bins=2048, out_lines=512, out_samples=1024, n_targets=1)
    keep_device_memory(0)
    release_device_memory()
    run_gpu(sc)
    base = free_bytes()
    run_gpu(sc)
    assert abs(free_bytes() - base) < (8 << 20)  # nothing of the call stays allocated
    try:
        keep_device_memory(-1)
        run_gpu(sc)
        kept = base - free_bytes()
        assert kept > (30 << 20)  # per-pixel tables + swath stay cached for the next call
        release_device_memory()
        assert abs(free_bytes() - base) < (8 << 20)
    finally:
        keep_device_memory(0)
