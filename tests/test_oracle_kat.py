"""Pins the CPU oracle: the reference's own data-free known-answer tests,
run against BOTH oracle builds -- the restated port (oracle/tdbp_oracle.cpp) and, when it
was built, oracle/_ref (the reference's sources compiled unchanged) -- plus port == _ref."""
import math

import numpy as np
import pytest

from isce3_b200 import core
from testkit import synth
from isce3_b200.core import LookSide, LUT2d, Orbit, OrbitInterpMethod
from isce3_b200.geometry import DEMInterpolator

C = core.speed_of_light


def both(oracles):
    port, ref = oracles
    return [o for o in (port, ref) if o is not None]


# ---- tests/cxx/isce3/focus/bistatic-delay.cpp:27-52 ----------------------------------------
def test_bistatic_delay_kat(oracles):
    p0, v = np.array([0.0, 0.0, 700e3]), np.array([0.0, 8000.0, 0.0])
    x = np.array([50e3, 20e3, 0.0])
    for o in both(oracles):
        for t in range(0, 11):
            p = p0 + v * t
            tau = o.bistatic_delay(p, v, x)
            r1 = np.linalg.norm(x - p)
            r2 = np.linalg.norm(p + v * tau - x)
            assert tau == pytest.approx((r1 + r2) / C, rel=4e-16, abs=0)


# ---- tests/cxx/isce3/core/orbit/orbit.cpp:414-551 -------------------------------------------
def _orbit(fn_pos, fn_vel, n=11, dt=10.0, method=OrbitInterpMethod.HERMITE):
    t = np.arange(n) * dt
    return Orbit.from_arrays(0.0, dt, np.array([fn_pos(ti) for ti in t]),
                             np.array([fn_vel(ti) for ti in t]), interp_method=method)


@pytest.mark.parametrize("method", [OrbitInterpMethod.HERMITE, OrbitInterpMethod.LEGENDRE])
def test_orbit_linear_and_circular_kat(oracles, method):
    x0, v0 = np.array([0.0, 0.0, 7e6]), np.array([7000.0, 1000.0, -100.0])
    lin = _orbit(lambda t: x0 + v0 * t, lambda t: v0, method=method)
    th0, om, R = 2 * math.pi / 8, 2 * math.pi / 7000.0, 8e6
    cpos = lambda t: R * np.array([math.cos(th0 + om * t), math.sin(th0 + om * t), 0.0])
    cvel = lambda t: R * om * np.array([-math.sin(th0 + om * t), math.cos(th0 + om * t), 0.0])
    circ = _orbit(cpos, cvel, method=method)
    for o in both(oracles):
        for t in (23.3, 36.7, 54.5, 89.3):
            st, p, v = o.orbit_interpolate(lin, t, 2)
            assert st == 0
            np.testing.assert_allclose(p, x0 + v0 * t, atol=1e-8)
            np.testing.assert_allclose(v, v0, atol=1e-8)
            st, p, v = o.orbit_interpolate(circ, t, 2)
            np.testing.assert_allclose(p, cpos(t), atol=1e-8 * 100 if method == 0 else 1e-8)
            np.testing.assert_allclose(v, cvel(t), atol=1e-7)


def test_orbit_border_modes(oracles):
    orb = _orbit(lambda t: np.array([t, 2 * t, 3 * t]), lambda t: np.array([1.0, 2.0, 3.0]))
    for o in both(oracles):
        st, p, v = o.orbit_interpolate(orb, -5.0, 2)  # FillNaN
        assert st == 2 and np.all(np.isnan(p)) and np.all(np.isnan(v))
        st, p, v = o.orbit_interpolate(orb, -5.0, 1)  # Extrapolate
        assert st == 0
        np.testing.assert_allclose(p, [-5.0, -10.0, -15.0], atol=1e-9)
        st, _, _ = o.orbit_interpolate(orb, -5.0, 0)  # Error -> OutOfRange
        assert st == -5


def test_host_orbit_matches_oracle(oracles):
    sc_orbit = synth.circular_orbit(7.1e6, 7500.0, 100.0, 160.0, 10.0)
    for o in both(oracles):
        for t in (101.3, 128.0, 155.55):
            _, p, v = o.orbit_interpolate(sc_orbit, t, 0)
            ph, vh = sc_orbit.interpolate(t)
            np.testing.assert_allclose(ph, p, rtol=0, atol=1e-6)
            np.testing.assert_allclose(vh, v, rtol=0, atol=1e-9)
    pm, vm = synth.interpolate_orbit_many(sc_orbit, np.array([101.3, 128.0, 155.55]))
    np.testing.assert_allclose(pm[1], sc_orbit.interpolate(128.0)[0], atol=1e-6)
    np.testing.assert_allclose(vm[2], sc_orbit.interpolate(155.55)[1], atol=1e-9)


# ---- tests/cxx/isce3/math/root_find1d.cpp:198-225 ---------------------------------------------
def test_brent_kat(oracles):
    problems = [
        (lambda x: x ** 3 - 2 * x - 5, 2.0, 3.0, 2.0945514815423265),
        (lambda x: math.cos(x) - x, 0.0, 1.0, 0.7390851332151607),
        (lambda x: math.exp(x) - 3 * x * x, 0.0, 1.0, 0.9100075724887090),
        (lambda x: x * math.exp(-x) - 0.1, 0.0, 1.0, 0.11183255915896297),
    ]
    for o in both(oracles):
        for f, a, b, root in problems:
            st, x = o.brent(a, b, f, 1e-12)
            assert st == 0 and abs(x - root) < 1e-11
        st, _ = o.brent(0.0, 1.0, lambda x: float("nan"), 1e-12)
        assert st != 0  # NaN-returning function cannot converge
        st, _ = o.brent(1.0, 2.0, lambda x: x, 1e-12)
        assert st == 11  # InvalidInterval: no sign change


# ---- tests/cxx/isce3/core/interp1d.cpp:97-141,270-309 -------------------------------------------
class _TestSignal:
    """Band-limited random signal: sum of sincs (interp1d.cpp:97-141)."""

    def __init__(self, n, bw, seed):
        rng = np.random.default_rng(seed)
        self.n, self.bw = n, bw
        self.w = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / math.sqrt(2 * n) * math.sqrt(n)
        self.t = np.arange(n)

    def eval(self, t):
        return np.sum(self.w * np.sinc(self.bw * (t - self.t)))

    def samples(self):
        return np.array([self.eval(t) for t in range(self.n)]).astype(np.complex64)


@pytest.mark.parametrize("make_kernel", [
    lambda: core.TabulatedKernelF32(core.KnabKernel(9.0, 0.8), 2048),
    lambda: core.ChebyKernelF32(core.KnabKernel(9.0, 0.8), 16),
    lambda: core.KnabKernelF32(9.0, 0.8),
])
def test_interp1d_thresholds(oracles, make_kernel):
    n, bw = 512, 0.8
    sig = _TestSignal(n, bw, 1234)
    z = sig.samples()
    rng = np.random.default_rng(99)
    pad = 16
    t = pad + (n - 2 * pad) * rng.random(300)
    truth = np.array([sig.eval(ti) for ti in t])
    kernel = make_kernel()
    for o in both(oracles):
        got = o.interp1d(kernel, z, t)
        corr = abs(np.vdot(truth, got)) / (np.linalg.norm(truth) * np.linalg.norm(got))
        dphi = np.degrees(np.angle(got * truth.conj()))
        ratio_db = 20 * np.log10(np.abs(got) / np.abs(truth))
        assert corr >= 0.998                    # min_cor
        assert np.std(dphi) <= 5.0              # max_phs
        assert abs(np.mean(ratio_db)) <= 0.5 and np.std(ratio_db) <= 0.5  # max_bias / spread
        # zero offset reproduces the samples (interp1d.cpp: exact to 1e-3 deg)
        t_int = np.arange(pad, n - pad, 7).astype(float)
        got0 = o.interp1d(kernel, z, t_int)
        np.testing.assert_allclose(got0, z[t_int.astype(int)], rtol=0, atol=2e-5 * np.abs(z).max())


def test_interp1d_edges_are_zero_padded(oracles):
    """CPU semantics: partial windows at the swath edges use zeros (detail/Interp1d.h:54-80),
    unlike the reference CUDA path which returns 0 (cuda/core/Interp1d.icc:18-20)."""
    kernel = core.TabulatedKernelF32(core.KnabKernel(9.0, 0.8), 2048)
    z = np.ones(32, np.complex64)
    for o in both(oracles):
        inside = o.interp1d(kernel, z, [16.3])[0]
        edge = o.interp1d(kernel, z, [1.3])[0]
        outside = o.interp1d(kernel, z, [-7.0, 40.0])
        assert abs(inside) > 0.9
        assert 0.3 < abs(edge) < abs(inside) + 0.2 and edge != 0
        assert np.all(outside == 0)


def test_kernel_tables_match_host_types(oracles):
    """TabulatedKernelF32 / ChebyKernelF32 built on the host (isce3_b200.core) hold exactly the
    table / coefficients the reference constructors produce (Kernels.icc:114-137,156-189)."""
    for o in both(oracles):
        for w, bw in ((9.0, 20.0 / 24.0), (8.0, 0.8), (16.0, 0.8)):
            tab = core.TabulatedKernelF32(core.KnabKernel(w, bw), 2048)
            np.testing.assert_array_equal(tab.table, o.tabulate_knab(w, bw, 2048))
            ch = core.ChebyKernelF32(core.KnabKernel(w, bw), 16)
            np.testing.assert_allclose(ch.coeffs, o.cheby_knab(w, bw, 16), rtol=0, atol=2e-7)
        t = np.linspace(-5, 5, 1001)
        tab = core.TabulatedKernelF32(core.KnabKernel(9.0, 0.8), 2048)
        np.testing.assert_array_equal(tab(t), o.kernel_eval(tab, t))
        lin = core.LinearKernelF32()
        np.testing.assert_array_equal(lin(t), o.kernel_eval(lin, t))


def test_azimuth_kernel_golden():
    """tests/python/extensions/pybind/core/kernels.py:15"""
    assert core.AzimuthKernel(1.0)(0.1) == pytest.approx(0.946, abs=1e-3)


# ---- tests/cxx/isce3/geometry/geometry/geometry_equator.cpp:42-160 ------------------------------
def _equator_orbit():
    hsat, omega = 700e3, math.radians(0.1)
    a = core.earth_semi_major_axis
    t = np.arange(10) * 10.0
    lon = omega * t
    r = a + hsat
    pos = np.stack([r * np.cos(lon), r * np.sin(lon), 0 * lon], -1)
    vel = np.stack([-omega * r * np.sin(lon), omega * r * np.cos(lon), 0 * lon], -1)
    return Orbit.from_arrays(0.0, 10.0, pos, vel), hsat, omega


@pytest.mark.parametrize("side", [LookSide.Left, LookSide.Right])
def test_rdr2geo_geo2rdr_equator(oracles, side):
    orbit, hsat, omega = _equator_orbit()
    a, e2 = core.earth_semi_major_axis, core.earth_eccentricity_squared
    b = a * math.sqrt(1 - e2)
    wvl = 0.24
    for o in both(oracles):
        for t in (15.0, 25.0, 55.0):
            for look_deg in (10.0, 25.0, 40.0):
                # closed form: target in the plane x-z rotated by lon(t), on the ellipsoid
                # geocentric latitude psi solved from the law of cosines along the ellipse
                sat_lon = omega * t
                # march: find psi (geocentric lat) with look angle look_deg by bisection
                def range_for(psi):
                    rt = a * b / math.hypot(b * math.cos(psi), a * math.sin(psi))
                    x, z = rt * math.cos(psi), rt * math.sin(psi)
                    return math.hypot(a + hsat - x, z), x, z
                lo, hi = 0.0, math.radians(20)
                for _ in range(100):
                    mid = 0.5 * (lo + hi)
                    rg, x, z = range_for(mid)
                    ang = math.degrees(math.atan2(z, a + hsat - x))
                    if ang < look_deg:
                        lo = mid
                    else:
                        hi = mid
                rng, x, z = range_for(0.5 * (lo + hi))
                sgn = 1.0 if side == LookSide.Left else -1.0  # left of +y velocity at lon 0 is +z
                want = np.array([x * math.cos(sat_lon), x * math.sin(sat_lon), sgn * z])
                ok, xyz = o.rdr2geo_bracket(t, rng, 0.0, orbit, DEMInterpolator(0.0), wvl, side)
                assert ok == 1
                np.testing.assert_allclose(xyz, want, rtol=0, atol=2e-3)
                llh = o.xyz_to_llh(xyz)
                assert abs(llh[2]) < 1e-4            # on the DEM (height 0)
                assert abs(llh[0] - sat_lon) < 1e-8  # zero Doppler: same longitude as the radar
                ok, t2, r2 = o.geo2rdr_bracket(xyz, orbit, LUT2d(), wvl, side)
                assert ok == 1 and abs(t2 - t) < 1e-6 and abs(r2 - rng) < 1e-4
                # wrong look side is rejected (Geo2Rdr.icc:234-236)
                other = LookSide.Right if side == LookSide.Left else LookSide.Left
                ok, _, _ = o.geo2rdr_bracket(xyz, orbit, LUT2d(), wvl, other)
                assert ok == 0


def test_ellipsoid_roundtrip_and_tropo(oracles):
    for o in both(oracles):
        for llh in ([0.3, -0.4, 120.0], [-2.0, 1.2, 8000.0], [3.0, 0.0, -50.0]):
            x = o.llh_to_xyz(llh)
            for got in (o.xyz_to_llh(x), synth.ecef_to_llh(x)):
                assert np.all(np.abs(got - llh) <= np.array([1e-12, 1e-12, 1e-8]))
            np.testing.assert_allclose(synth.llh_to_ecef(*llh), x, rtol=0, atol=1e-6)
        # zenith path at sea level: 2*ZPD/c (DryTroposphereModel.icc:10-29)
        llh = np.array([0.1, 0.2, 0.0])
        n = np.array([math.cos(0.2) * math.cos(0.1), math.cos(0.2) * math.sin(0.1), math.sin(0.2)])
        p = o.llh_to_xyz(llh) + 700e3 * n
        assert o.dry_tropo_tsx(p, llh) == pytest.approx(2 * 2.3 / C, rel=1e-9)
        llh[2] = 6000.0
        p = o.llh_to_xyz(llh) + 700e3 * n
        assert o.dry_tropo_tsx(p, llh) == pytest.approx(2 * 2.3 / C / math.e, rel=1e-9)
        assert synth.dry_tropo_delay_tsx(p, llh) == pytest.approx(o.dry_tropo_tsx(p, llh), rel=1e-12)


# ---- samplers (restated for both oracle builds: analytic properties) -------------------------------
def test_lut2d_and_dem_samplers(oracles):
    y, x = np.mgrid[0:12, 0:15].astype(float)
    plane = 3.0 + 0.5 * x - 0.25 * y
    lut = LUT2d(100.0, 10.0, 2.0, 0.5, plane, "bilinear", False)
    for o in both(oracles):
        assert o.lut2d_eval(lut, 10.0 + 0.5 * 3.3, 100.0 + 2.0 * 4.7) == pytest.approx(3 + 0.5 * 4.7 - 0.25 * 3.3)
        assert o.lut2d_eval(lut, -100.0, 1e9) == pytest.approx(plane[0, -1])  # clamped
        assert o.lut2d_eval(LUT2d(), 1.0, 2.0) == 0.0
        assert lut.eval(10.0 + 0.5 * 3.3, 100.0 + 2.0 * 4.7) == pytest.approx(
            o.lut2d_eval(lut, 10.0 + 0.5 * 3.3, 100.0 + 2.0 * 4.7))
        for method in ("bilinear", "bicubic", "biquintic"):
            # every method reproduces a plane exactly away from the borders
            h = (100.0 + 2.0 * x + 3.0 * y).astype(np.float32)
            dem = DEMInterpolator.from_array(h, -60.0, 10.0, 0.01, -0.01, 4326, method)
            lon, lat = math.radians(-60.0 + 0.01 * 6.25), math.radians(10.0 - 0.01 * 5.5)
            assert o.dem_interp(dem, lon, lat) == pytest.approx(100 + 2 * 6.25 + 3 * 5.5, abs=2e-3)
            # outside the [2, n-1) margin -> reference height (DEMInterpolator.cpp:649-653)
            assert o.dem_interp(dem, math.radians(-60.0 + 0.01 * 0.5), lat) == pytest.approx(dem.ref_height)
        assert o.dem_interp(DEMInterpolator(123.0), 0.1, 0.2) == 123.0


# ---- port == reference sources on the whole path -----------------------------------------------------
@pytest.mark.parametrize("name,kw", [
    ("c1", dict(pulses=384, bins=768, out_lines=12, out_samples=40)),
    ("c2", dict(pulses=512, bins=768, out_lines=6, out_samples=24, n_targets=1)),
    ("c5", dict(pulses=2048, bins=1024, out_lines=6, out_samples=20, n_targets=1, taps=8)),
    ("c4", dict(pulses=512, bins=1024, out_lines=5, out_samples=16, n_targets=1)),
])
def test_port_equals_reference_sources(oracles, name, kw):
    port, ref = oracles
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference at build time)")
    sc = synth.make_scene(name, **kw)
    shape = (sc.out_geometry.grid_length, sc.out_geometry.grid_width)
    a, b = np.zeros(shape, np.complex64), np.zeros(shape, np.complex64)
    ha, hb = np.zeros(shape, np.float32), np.zeros(shape, np.float32)
    ea = port.backproject(a, *sc.backproject_args(), height=ha)
    eb = ref.backproject(b, *sc.backproject_args(), height=hb)
    assert ea == eb
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = np.isfinite(b)
    assert np.linalg.norm((a - b)[m]) <= 1e-6 * np.linalg.norm(b[m])
    np.testing.assert_allclose(ha, hb, rtol=0, atol=1e-5)


# ---- map projections of raster DEMs (core/Projections.cpp) ---------------------------------

PROJ_POINTS = [(-118.3, 34.1), (11.2, -3.4), (-45.0, 72.0), (100.0, -75.0), (179.2, 60.0), (3.0, 0.0)]


def _epsg_cases(lon_deg, lat_deg):
    import math
    from isce3_b200.projections import utm_epsg_for
    return [utm_epsg_for(math.radians(lon_deg), math.radians(lat_deg)), 3413, 3031, 6933, 4326]


def test_projection_known_answers(oracles):
    """Textbook anchors: a point on a UTM central meridian at the equator maps to
    (500000, 0) north / (500000, 10000000) south; the pole is the origin of the polar
    stereographic grids; EASE-2 (6933) is linear in longitude."""
    import math
    for o in [o for o in oracles if o is not None]:
        st, x, y = o.project_forward(32631, math.radians(3.0), 0.0)
        assert st == 0 and abs(x - 500000.0) < 1e-6 and abs(y) < 1e-6
        st, x, y = o.project_forward(32731, math.radians(3.0), 0.0)
        assert st == 0 and abs(x - 500000.0) < 1e-6 and abs(y - 1.0e7) < 1e-6
        st, x, y = o.project_forward(3413, 0.3, math.pi / 2)
        assert st == 0 and abs(x) < 1e-6 and abs(y) < 1e-6
        st, x1, _ = o.project_forward(6933, 0.5, 0.2)
        st, x2, _ = o.project_forward(6933, 1.0, 0.2)
        assert abs(x2 - 2 * x1) < 1e-6
        # UTM refuses points a quarter of the globe away from the zone (Projections.cpp:199-212)
        assert o.project_forward(32631, math.radians(3.0 + 89.9), 0.0)[0] == 1


def test_projection_port_python_and_reference_agree(oracles):
    """Restated forward projections (oracle port, host mirror isce3_b200.projections) against
    the reference's own Projections.cpp (oracle/_ref) on scattered points."""
    import math
    from isce3_b200.projections import make_projection
    for lon_d, lat_d in PROJ_POINTS:
        lon, lat = math.radians(lon_d), math.radians(lat_d)
        for epsg in _epsg_cases(lon_d, lat_d):
            if epsg == 3413 and lat_d < -60 or epsg == 3031 and lat_d > 60:
                continue
            vals = [o.project_forward(epsg, lon, lat) for o in oracles if o is not None]
            px, py = make_projection(epsg).forward(lon, lat)
            for st, x, y in vals:
                assert st == 0
                assert abs(x - px) <= 1e-6 * max(1.0, abs(px)) and abs(y - py) <= 1e-6 * max(1.0, abs(py))
            if len(vals) == 2:
                assert abs(vals[0][1] - vals[1][1]) <= 1e-7 * max(1.0, abs(vals[1][1]))
                assert abs(vals[0][2] - vals[1][2]) <= 1e-7 * max(1.0, abs(vals[1][2]))


def test_projected_dem_sampling_matches_between_oracles(oracles):
    """DEMInterpolator.interpolateLonLat on a UTM raster: port (restated projection + sampler)
    vs reference projection + sampler."""
    import math
    from testkit import synth
    from isce3_b200.projections import utm_epsg_for
    lon, lat = math.radians(-118.3), math.radians(34.1)
    dem = synth.synthetic_dem_projected(utm_epsg_for(lon, lat), lon, lat, 20e3, posting_m=100.0)
    rng = np.random.default_rng(3)
    for _ in range(50):
        lo = lon + rng.uniform(-0.15, 0.15) * math.pi / 180
        la = lat + rng.uniform(-0.15, 0.15) * math.pi / 180
        h = [o.dem_interp(dem, lo, la) for o in oracles if o is not None]
        assert 0.0 <= h[0] <= 2000.0
        if len(h) == 2:
            assert abs(h[0] - h[1]) <= 1e-4
