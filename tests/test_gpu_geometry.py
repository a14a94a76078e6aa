"""Batch geometry API (i3b_rdr2geo_bracket_batch / i3b_geo2rdr_bracket_batch) against the
reference's own rdr2geo_bracket / geo2rdr_bracket (oracle/_ref: geometry/rdr2geo_roots.cpp,
geo2rdr_roots.cpp compiled unchanged; else the restated port), point by point.

Tolerances follow the solvers' own stopping rules: rdr2geo stops at tol_height = 1e-5 m of
height error, so two correct solvers agree to ~1e-4 m in position; geo2rdr stops at
tol_aztime = 1e-7 s (7.5e-4 m along track at orbital speed)."""
import numpy as np
import pytest

from isce3_b200 import core
from isce3_b200 import geometry as geo
from isce3_b200.core import LookSide, LUT2d
from testkit import synth

pytestmark = pytest.mark.gpu


def _points(sc, n, seed):
    rng = np.random.default_rng(seed)
    g = sc.out_geometry.radar_grid
    t = g.sensing_start + rng.uniform(0, g.length - 1, n) / g.prf
    r = g.starting_range + rng.uniform(0, g.width - 1, n) * g.range_pixel_spacing
    return t, r


@pytest.mark.parametrize("case", ["flat", "raster", "raster_utm"])
def test_rdr2geo_bracket_batch_matches_reference(oracle, case):
    kw = dict(pulses=1024, bins=2048, out_lines=64, out_samples=512, n_targets=1)
    if case == "flat":
        sc = synth.make_scene("c2", **kw)
    else:
        sc = synth.make_scene("c4", dem_epsg="utm" if case == "raster_utm" else None, **kw)
    orbit, dem = sc.out_geometry.orbit, sc.dem
    wvl = core.speed_of_light / sc.fc
    side = sc.out_geometry.look_side
    t, r = _points(sc, 300, 3)
    fd = np.random.default_rng(4).uniform(-300.0, 300.0, t.size)
    for dop in (None, fd):
        xyz, status = geo.rdr2geo_bracket(t, r, dop, orbit, dem, wvl, side)
        assert np.all(status == 0) and np.isfinite(xyz).all()
        for i in range(0, t.size, 7):
            ok, ref = oracle.rdr2geo_bracket(t[i], r[i], 0.0 if dop is None else dop[i], orbit, dem, wvl, side)
            assert ok
            assert np.linalg.norm(xyz[i] - ref) <= 2e-4, (i, xyz[i] - ref)
    # look-angle bracket that excludes the solution: per-point failure codes, NaN positions
    xyz, status = geo.rdr2geo_bracket(t[:16], r[:16], None, orbit, dem, wvl, side, look_min=0.0, look_max=0.3)
    assert np.all(status != 0) and np.isnan(xyz).all()


def test_geo2rdr_bracket_batch_matches_reference(oracle):
    sc = synth.make_scene("c2", pulses=1024, bins=2048, out_lines=64, out_samples=512, n_targets=1,
                          doppler_lut=True)
    orbit = sc.in_geometry.orbit
    wvl = core.speed_of_light / sc.fc
    side = sc.in_geometry.look_side
    t, r = _points(sc, 200, 5)
    xyz, status = geo.rdr2geo_bracket(t, r, None, orbit, sc.dem, wvl, side)
    assert np.all(status == 0)
    for dop in (LUT2d(), sc.in_geometry.doppler):
        tt, rr, st = geo.geo2rdr_bracket(xyz, orbit, dop, wvl, side)
        assert np.all(st == 0)
        for i in range(0, t.size, 5):
            ok, tref, rref = oracle.geo2rdr_bracket(xyz[i], orbit, dop, wvl, side)
            assert ok
            assert abs(tt[i] - tref) <= 3e-7 and abs(rr[i] - rref) <= 2e-3
    # zero-Doppler geo2rdr inverts zero-Doppler rdr2geo
    tt, rr, _ = geo.geo2rdr_bracket(xyz, orbit, LUT2d(), wvl, side)
    assert np.max(np.abs(tt - t)) <= 3e-7 and np.max(np.abs(rr - r)) <= 2e-3
    # wrong look side is reported per point (Geo2Rdr.icc:234-236)
    other = LookSide.Right if side == LookSide.Left else LookSide.Left
    _, _, st = geo.geo2rdr_bracket(xyz[:8], orbit, LUT2d(), wvl, other)
    assert np.all(st == 7)
    # a time bracket that excludes the root
    _, _, st = geo.geo2rdr_bracket(xyz[:8], orbit, LUT2d(), wvl, side, time_start=float(t.max()) + 1.0)
    assert np.all(st != 0)
