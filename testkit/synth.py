"""Deterministic synthetic TDBP scenes (bench + parity inputs).

The reference's only backproject fixture (tests/data/point-target-sim-rc.h5) is
stripped from the mount, so scenes are generated here from the
recipe in SURVEY.md 8(d): analytic orbits, point targets placed by a zero-Doppler
range/height solve, range-compressed echoes

    rc[k, i] += A * sinc(B/fs * (i - u_k)) * exp(-j 2 pi fc tau_k),
    tau_k = bistaticDelay(pos_k, vel_k, x) [+ dryTropoDelayTSX],   u_k = (tau_k - tau0)/dtau

(cxx/isce3/focus/BistaticDelay.icc:10-17, DryTroposphereModel.icc:10-29) evaluated in
float64 and stored as complex64, optional complex Gaussian noise with a fixed seed.

Named configurations follow BASELINE.json ``configs``:
  c1       single point target, 2048 pulses x 4096 bins -> 512 x 512
  c2       NISAR 20 MHz-like frame, 16384 x 12288 -> 8192 x 8192, flat DEM, tsx
  c4       80 MHz-like swath, raster DEM (EPSG:4326, biquintic), tsx, 9 x 9 targets
  c5       airborne, curved track, 65536 pulses, Knab width 8/16/32
plus ``scale`` < 1 variants that keep the geometry and shrink the array sizes.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

from isce3_b200 import core
from isce3_b200.container import RadarGeometry
from isce3_b200.core import DateTime, LookSide, LUT2d, Orbit, speed_of_light as C0
from isce3_b200.geometry import DEMInterpolator
from isce3_b200.product import RadarGridParameters

A_WGS84 = core.earth_semi_major_axis
E2_WGS84 = core.earth_eccentricity_squared
GM = 3.986004418e14


# ---- small geodesy helpers (host, float64) --------------------------------

def llh_to_ecef(lon, lat, h):
    """cxx/isce3/core/Ellipsoid.h:177-190"""
    re = A_WGS84 / np.sqrt(1.0 - E2_WGS84 * np.sin(lat) ** 2)
    return np.stack([(re + h) * np.cos(lat) * np.cos(lon), (re + h) * np.cos(lat) * np.sin(lon),
                     (re * (1.0 - E2_WGS84) + h) * np.sin(lat)], axis=-1)


def ecef_to_llh(x):
    """Vermeille (2002) closed form as cxx/isce3/core/Ellipsoid.h:196-224."""
    x = np.asarray(x, dtype=np.float64)
    e2, e4, a2 = E2_WGS84, E2_WGS84 ** 2, A_WGS84 ** 2
    p = (x[..., 0] ** 2 + x[..., 1] ** 2) / a2
    q = (1.0 - e2) * x[..., 2] ** 2 / a2
    r = (p + q - e4) / 6.0
    s = e4 * p * q / (4.0 * r ** 3)
    t = np.cbrt(1.0 + s + np.sqrt(s * (2.0 + s)))
    u = r * (1.0 + t + 1.0 / t)
    rv = np.sqrt(u * u + e4 * q)
    w = e2 * (u + rv - q) / (2.0 * rv)
    k = np.sqrt(u + rv + w * w) - w
    d = k * np.sqrt(x[..., 0] ** 2 + x[..., 1] ** 2) / (k + e2)
    lat = np.arctan2(x[..., 2], d)
    lon = np.arctan2(x[..., 1], x[..., 0])
    h = (k + e2 - 1.0) * np.sqrt(d * d + x[..., 2] ** 2) / k
    return np.stack([lon, lat, h], axis=-1)


def interpolate_orbit_many(orbit: Orbit, t):
    """Vectorised cubic Hermite (core/detail/InterpolateOrbit.icc:15-109) for uniform
    state-vector spacing; returns (pos[n,3], vel[n,3])."""
    if orbit.interp_method != core.OrbitInterpMethod.HERMITE:
        pv = [orbit.interpolate(ti) for ti in np.atleast_1d(t)]
        return np.array([p for p, _ in pv]), np.array([v for _, v in pv])
    t = np.atleast_1d(np.asarray(t, dtype=np.float64))
    n, t0, dt = orbit.size, orbit.time.first, orbit.time.spacing
    search = np.floor((t - t0) / dt + 1).astype(np.int64)
    search = np.where(t < t0, 0, np.where(t > orbit.time.last, n, search))
    idx = np.clip(search - 2, 0, n - 4)
    tt = t0 + (idx[:, None] + np.arange(4)[None, :]) * dt  # [m,4]
    P = orbit.position[idx[:, None] + np.arange(4)[None, :]]  # [m,4,3]
    V = orbit.velocity[idx[:, None] + np.arange(4)[None, :]]
    d = t[:, None] - tt
    f1 = d
    gsum = np.zeros_like(tt)
    h = np.ones_like(tt)
    hdot = np.zeros_like(tt)
    for i in range(4):
        for j in range(4):
            if j == i:
                continue
            gsum[:, i] += 1.0 / (tt[:, i] - tt[:, j])
            h[:, i] *= d[:, j] / (tt[:, i] - tt[:, j])
            prod = 1.0 / (tt[:, i] - tt[:, j])
            for k in range(4):
                if k != i and k != j:
                    prod = prod * d[:, k] / (tt[:, i] - tt[:, k])
            hdot[:, i] += prod
    f0 = 1.0 - 2.0 * gsum * d
    g1 = h + 2.0 * hdot * d
    g0 = 2.0 * (f0 * hdot - gsum * h)
    pos = ((h * h)[..., None] * (P * f0[..., None] + V * f1[..., None])).sum(1)
    vel = (h[..., None] * (P * g0[..., None] + V * g1[..., None])).sum(1)
    return pos, vel


def bistatic_delay(p, v, x):
    """cxx/isce3/focus/BistaticDelay.icc:10-17, vectorised over pulses."""
    r = x - p
    return 2.0 * ((r * v).sum(-1) - C0 * np.linalg.norm(r, axis=-1)) / ((v * v).sum(-1) - C0 * C0)


def dry_tropo_delay_tsx(p, llh):
    """cxx/isce3/focus/DryTroposphereModel.icc:10-29"""
    x = llh_to_ecef(llh[0], llh[1], llh[2])
    r_hat = (p - x) / np.linalg.norm(p - x)
    n_hat = np.array([math.cos(llh[1]) * math.cos(llh[0]), math.cos(llh[1]) * math.sin(llh[0]),
                      math.sin(llh[1])])
    return 2.0 * 2.3 * math.exp(-llh[2] / 6000.0) / (C0 * float(r_hat @ n_hat))


# ---- orbits ------------------------------------------------------------------

def circular_orbit(radius, speed, t_start, t_end, dt_sv, inclination_deg=98.4,
                   lon_node_deg=-60.0, cross_track_amp=0.0, cross_track_period=1.0,
                   epoch=None, theta0_deg=10.0):
    """Circular track of given radius/speed in an inclined plane (ECEF held fixed);
    optional sinusoidal out-of-plane deviation (airborne 'curved' track)."""
    inc, lon0 = math.radians(inclination_deg), math.radians(lon_node_deg)
    e1 = np.array([math.cos(lon0), math.sin(lon0), 0.0])
    e2 = np.array([-math.sin(lon0) * math.cos(inc), math.cos(lon0) * math.cos(inc), math.sin(inc)])
    e3 = np.cross(e1, e2)
    n0 = int(math.floor(t_start / dt_sv)) - 3
    n1 = int(math.ceil(t_end / dt_sv)) + 3
    t = np.arange(n0, n1 + 1) * float(dt_sv)
    w = speed / radius
    th = math.radians(theta0_deg) + w * t
    ph = 2.0 * math.pi * t / cross_track_period
    pos = radius * (np.cos(th)[:, None] * e1 + np.sin(th)[:, None] * e2) + \
        cross_track_amp * np.sin(ph)[:, None] * e3
    vel = radius * w * (-np.sin(th)[:, None] * e1 + np.cos(th)[:, None] * e2) + \
        cross_track_amp * (2.0 * math.pi / cross_track_period) * np.cos(ph)[:, None] * e3
    return Orbit.from_arrays(t[0], float(dt_sv), pos, vel, epoch or DateTime(2025, 1, 1))


# ---- target placement ------------------------------------------------------------

def zero_doppler_target(orbit: Orbit, t, r, side: LookSide, height_fn):
    """ECEF point at slant range r in the plane normal to the velocity at time t whose
    height above the ellipsoid equals height_fn(lon, lat): bisection on the
    pseudo-look angle, same construction as geometry/detail/Rdr2Geo.icc:184-214."""
    P, V = orbit.interpolate(t)
    a = V / np.linalg.norm(V)
    right = np.cross(a, P)
    right /= np.linalg.norm(right)
    down = np.cross(a, right)
    hvec = right if side == LookSide.Right else -right

    def xyz(look):
        return P + r * math.sin(look) * hvec + r * math.cos(look) * down

    def f(look):
        llh = ecef_to_llh(xyz(look))
        return llh[2] - height_fn(llh[0], llh[1])

    lo, hi = 0.0, math.pi / 2
    flo = f(lo)
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        fm = f(mid)
        if (fm < 0) == (flo < 0):
            lo, flo = mid, fm
        else:
            hi = mid
        if hi - lo < 1e-15:
            break
    return xyz(0.5 * (lo + hi))


@dataclasses.dataclass
class Target:
    az_index: float  # position in the OUTPUT grid (line, sample)
    rg_index: float
    xyz: np.ndarray
    amplitude: complex = 1.0


@dataclasses.dataclass
class Scene:
    name: str
    in_geometry: RadarGeometry
    out_geometry: RadarGeometry
    rc: np.ndarray
    dem: DEMInterpolator
    fc: float
    ds: float
    kernel: object
    dry_tropo_model: str
    targets: list
    range_bandwidth: float
    range_sample_rate: float
    rdr2geo_params: dict = dataclasses.field(default_factory=dict)
    geo2rdr_params: dict = dataclasses.field(default_factory=dict)
    pulse_times: object = None  # float64 [pulses] when the pulse train is non-uniform

    def backproject_args(self):
        """Positional arguments 2..9 of backproject (after `out`)."""
        return (self.out_geometry, self.rc, self.in_geometry, self.dem, self.fc, self.ds,
                self.kernel, self.dry_tropo_model, self.rdr2geo_params, self.geo2rdr_params)

    def out_subgrid(self, a0, a1, r0=None, r1=None) -> RadarGeometry:
        g = self.out_geometry.radar_grid[a0:a1, slice(r0, r1)]
        return RadarGeometry(g, self.out_geometry.orbit, self.out_geometry.doppler)


def simulate_echoes(rc, in_grid: RadarGridParameters, orbit: Orbit, targets, fc, bandwidth, fs,
                    tau_atm=None, halfwidth=48, pulse_window=None, pulse_times=None):
    """Add point-target echoes into rc (complex64 [pulses, bins]) in place.  ``pulse_times``:
    explicit (non-uniform) azimuth time of every pulse; default the uniform grid of ``in_grid``."""
    npulse, nr = rc.shape
    tk = in_grid.sensing_start + np.arange(npulse) / in_grid.prf
    if pulse_times is not None:
        tk = np.asarray(pulse_times, dtype=np.float64)
    pos, vel = interpolate_orbit_many(orbit, tk)
    tau0 = 2.0 * in_grid.starting_range / C0
    dtau = 2.0 * in_grid.range_pixel_spacing / C0
    offs = np.arange(-halfwidth, halfwidth + 1)
    for m, tg in enumerate(targets):
        tau = bistatic_delay(pos, vel, tg.xyz[None, :])
        if tau_atm is not None:
            tau = tau + tau_atm[m]
        u = (tau - tau0) / dtau
        k = np.arange(npulse)
        if pulse_window is not None:
            k0, k1 = pulse_window[m]
            k = k[max(k0, 0):min(k1, npulse)]
        i0 = np.rint(u[k]).astype(np.int64)
        idx = i0[:, None] + offs[None, :]
        ok = (idx >= 0) & (idx < nr)
        env = np.sinc((bandwidth / fs) * (idx - u[k][:, None]))
        cyc = fc * tau[k]
        ph = np.exp(-2j * np.pi * (cyc - np.rint(cyc)))
        val = (tg.amplitude * env * ph[:, None]).astype(np.complex64)
        kk = np.broadcast_to(k[:, None], idx.shape)
        np.add.at(rc, (kk[ok], idx[ok]), val[ok])
    return rc


def add_noise(rc, sigma, seed=1234, block=1 << 22):
    """Complex Gaussian noise of std `sigma` per component-pair (fixed seed)."""
    rng = np.random.default_rng(seed)
    flat = rc.reshape(-1).view(np.float32)
    s = np.float32(sigma / math.sqrt(2.0))
    for a in range(0, flat.size, block):
        b = min(a + block, flat.size)
        flat[a:b] += s * rng.standard_normal(b - a, dtype=np.float32)
    return rc


def knab_table_kernel(width, bandwidth, n=2048):
    """What the workflow passes: TabulatedKernelF32(KnabKernel(w, bw), 2048)
    (python/packages/nisar/workflows/focus.py:796-808)."""
    return core.TabulatedKernelF32(core.KnabKernel(float(width), float(bandwidth)), n)


def synthetic_dem(lon_c, lat_c, half_extent_deg, posting_deg=1.0 / 3600, hmin=0.0, hmax=2000.0,
                  method="biquintic"):
    """Smooth relief: sum of 2-D sinusoids in [hmin, hmax], EPSG:4326, north-up."""
    n = int(round(2 * half_extent_deg / posting_deg)) + 1
    lon = lon_c - half_extent_deg + np.arange(n) * posting_deg
    lat = lat_c + half_extent_deg - np.arange(n) * posting_deg
    h = dem_height_fn(hmin, hmax)(np.radians(lon)[None, :], np.radians(lat)[:, None])
    return DEMInterpolator.from_array(h.astype(np.float32), lon[0], lat[0], posting_deg,
                                      -posting_deg, 4326, method)


def synthetic_dem_projected(epsg, lon_c, lat_c, half_extent_m, posting_m=30.0, hmin=0.0, hmax=2000.0,
                            method="biquintic"):
    """Same kind of smooth relief on a grid of projected coordinates (UTM / polar
    stereographic / EASE-2), north-up, centred on (lon_c, lat_c) [radians]."""
    from isce3_b200.projections import make_projection
    xc, yc = make_projection(epsg).forward(lon_c, lat_c)
    n = int(round(2 * half_extent_m / posting_m)) + 1
    x = xc - half_extent_m + np.arange(n) * posting_m
    y = yc + half_extent_m - np.arange(n) * posting_m
    X, Y = x[None, :] - xc, y[:, None] - yc
    z = (np.sin(2 * np.pi * X / 31e3) * np.cos(2 * np.pi * Y / 23e3) +
         0.5 * np.sin(2 * np.pi * (X + Y) / 11e3) + 0.25 * np.cos(2 * np.pi * (X - 2 * Y) / 7e3))
    h = hmin + (hmax - hmin) * (z + 1.75) / 3.5
    return DEMInterpolator.from_array(h.astype(np.float32), x[0], y[0], posting_m, -posting_m,
                                      epsg, method)


def dem_height_fn(hmin=0.0, hmax=2000.0):
    def f(lon, lat):
        lon_d, lat_d = np.degrees(lon), np.degrees(lat)
        z = (np.sin(2 * np.pi * lon_d / 0.31) * np.cos(2 * np.pi * lat_d / 0.23) +
             0.5 * np.sin(2 * np.pi * (lon_d + lat_d) / 0.11) +
             0.25 * np.cos(2 * np.pi * (lon_d - 2 * lat_d) / 0.07))
        return hmin + (hmax - hmin) * (z + 1.75) / 3.5
    return f


def _dem_sample_fn(dem: DEMInterpolator):
    """Height function used for target placement: bilinear sample of the raster (host)."""
    if not dem.have_raster:
        return lambda lon, lat: dem.ref_height

    from isce3_b200.projections import make_projection
    proj = make_projection(dem.epsg_code)

    def f(lon, lat):
        x, y = proj.forward(lon, lat)
        col = (x - dem.x_start) / dem.delta_x
        row = (y - dem.y_start) / dem.delta_y
        c0, r0 = int(math.floor(col)), int(math.floor(row))
        if r0 < 2 or r0 >= dem.length - 1 or c0 < 2 or c0 >= dem.width - 1:
            return dem.ref_height  # same margin rule as DEMInterpolator.cpp:649-653
        fc_, fr = col - c0, row - r0
        z = dem.data
        return float(z[r0, c0] * (1 - fc_) * (1 - fr) + z[r0, c0 + 1] * fc_ * (1 - fr) +
                     z[r0 + 1, c0] * (1 - fc_) * fr + z[r0 + 1, c0 + 1] * fc_ * fr)
    return f


def make_scene(name="c1", *, pulses=None, bins=None, out_lines=None, out_samples=None,
               n_targets=None, noise_db=None, taps=None, seed=1234, dry_tropo_model=None,
               with_dem=None, ds=None, table_size=2048, doppler_lut=False,
               look_side=LookSide.Left, out_range_spacing_ratio=1.0, out_prf_ratio=1.0,
               dem_epsg=None, prf_dither=0.0):
    """Build one of the named synthetic configurations (see module docstring).
    ``dem_epsg``: CRS of the raster DEM (configurations with relief): None / 4326, "utm" (the
    zone of the scene centre) or an EPSG code createProj knows (3031, 3413, 6933, 326xx...).
    ``prf_dither``: > 0 makes the pulse train NON-uniform -- pulse k is transmitted at
    ``t0 + (k + d_k) / prf`` with d_k a fixed slow pattern of that amplitude (in pulse
    intervals) plus a small pulse-to-pulse jitter, like a dithered-PRF acquisition;
    the times are returned in ``Scene.pulse_times`` and the echoes are simulated there."""
    base = name.lower()
    airborne = base.startswith("c5")
    fc = 1.2575e9
    wvl = C0 / fc
    if airborne:
        cfg = dict(pulses=65536, bins=8192, out_lines=2048, out_samples=2048, prf=500.0,
                   fs=100e6, bw=80e6, r0=14.0e3, ds=1.0, n_targets=3, tropo="nodelay",
                   dem=False, taps=16, noise_db=-40.0)
        radius, speed = A_WGS84 + 12.5e3, 220.0
        sv_dt = 1.0
    elif base.startswith("c4"):
        cfg = dict(pulses=16384, bins=32768, out_lines=2048, out_samples=8192, prf=1520.0,
                   fs=96e6, bw=80e6, r0=900.0e3, ds=6.0, n_targets=9, tropo="tsx", dem=True,
                   taps=9, noise_db=-40.0)
        radius = A_WGS84 + 747.0e3
        speed = math.sqrt(GM / radius)
        sv_dt = 10.0
    elif base.startswith("c2") or base.startswith("c3"):
        cfg = dict(pulses=16384, bins=12288, out_lines=8192, out_samples=8192, prf=1520.0,
                   fs=24e6, bw=20e6, r0=955.0e3, ds=6.0, n_targets=3, tropo="tsx", dem=False,
                   taps=9, noise_db=-40.0)
        radius = A_WGS84 + 747.0e3
        speed = math.sqrt(GM / radius)
        sv_dt = 10.0
    else:  # c1
        cfg = dict(pulses=2048, bins=4096, out_lines=512, out_samples=512, prf=1520.0,
                   fs=24e6, bw=20e6, r0=955.0e3, ds=6.0, n_targets=1, tropo="nodelay",
                   dem=False, taps=9, noise_db=None)
        radius = A_WGS84 + 747.0e3
        speed = math.sqrt(GM / radius)
        sv_dt = 10.0
    for key, val in (("pulses", pulses), ("bins", bins), ("out_lines", out_lines),
                     ("out_samples", out_samples), ("n_targets", n_targets), ("taps", taps),
                     ("tropo", dry_tropo_model), ("dem", with_dem), ("ds", ds)):
        if val is not None:
            cfg[key] = val
    if noise_db is not None:
        cfg["noise_db"] = None if noise_db is False else noise_db
    prf, fs, bw = cfg["prf"], cfg["fs"], cfg["bw"]
    dr = C0 / (2.0 * fs)
    npulse, nbins = int(cfg["pulses"]), int(cfg["bins"])
    nl, ns = int(cfg["out_lines"]), int(cfg["out_samples"])
    t_first = 128.0  # binary-friendly epoch offset (exact in DateTime round trips)
    duration = npulse / prf
    period = 97.0 if airborne else 1.0
    orbit = circular_orbit(radius, speed, t_first - 2 * sv_dt, t_first + duration + 2 * sv_dt,
                           sv_dt, cross_track_amp=50.0 if airborne else 0.0,
                           cross_track_period=period,
                           inclination_deg=60.0 if airborne else 98.4)
    epoch = orbit.reference_epoch
    in_grid = RadarGridParameters(t_first, wvl, prf, cfg["r0"], dr, look_side, npulse, nbins, epoch)
    out_prf = prf * out_prf_ratio
    out_dr = dr * out_range_spacing_ratio
    t_mid = t_first + 0.5 * (npulse - 1) / prf
    r_mid = cfg["r0"] + 0.5 * (nbins - 1) * dr
    out_t0 = t_mid - 0.5 * (nl - 1) / out_prf
    out_r0 = r_mid - 0.5 * (ns - 1) * out_dr
    out_grid = RadarGridParameters(out_t0, wvl, out_prf, out_r0, out_dr, look_side, nl, ns, epoch)

    if doppler_lut:
        # small linear-in-range Doppler on the INPUT grid
        ya = np.array([orbit.start_time, orbit.end_time])
        xa = np.array([cfg["r0"] - 1e4, cfg["r0"] + nbins * dr + 1e4])
        data = np.array([[-30.0, 30.0], [-30.0, 30.0]])
        in_dop = LUT2d(xa[0], ya[0], xa[1] - xa[0], ya[1] - ya[0], data, "bilinear", False)
    else:
        in_dop = LUT2d()
    in_geom = RadarGeometry(in_grid, orbit, in_dop)
    out_geom = RadarGeometry(out_grid, orbit, LUT2d())  # zero-Doppler output grid (focus.py:1998)

    if cfg["dem"]:
        ctr = ecef_to_llh(zero_doppler_target(orbit, t_mid, r_mid, look_side, lambda lo, la: 1000.0))
        swath_m = max(nbins * dr / math.sin(math.radians(30.0)), duration * speed) * 0.5 + 3.0e4
        half_deg = math.degrees(swath_m / A_WGS84) / max(math.cos(ctr[1]), 0.2)
        if dem_epsg in (None, 4326):
            dem = synthetic_dem(math.degrees(ctr[0]), math.degrees(ctr[1]), min(half_deg, 3.0))
        else:
            from isce3_b200.projections import utm_epsg_for
            code = utm_epsg_for(ctr[0], ctr[1]) if dem_epsg == "utm" else int(dem_epsg)
            dem = synthetic_dem_projected(code, ctr[0], ctr[1], min(swath_m, 3.0e5))
    else:
        dem = DEMInterpolator(0.0)
    hfn = _dem_sample_fn(dem)

    # targets on a square array centred in the output grid
    nt = int(cfg["n_targets"])
    side_n = int(round(math.sqrt(nt))) if nt > 1 else 1
    if side_n * side_n != nt:
        side_n = nt  # 1-D row of targets along range
    fr = [0.5] if side_n == 1 else list(np.linspace(0.2, 0.8, side_n))
    targets, tau_atm = [], []
    for fa in (fr if side_n * side_n == nt else [0.5]):
        for frg in fr:
            ai, ri = round(fa * (nl - 1)), round(frg * (ns - 1))
            tt, rr = out_t0 + ai / out_prf, out_r0 + ri * out_dr
            xyz = zero_doppler_target(orbit, tt, rr, look_side, hfn)
            targets.append(Target(float(ai), float(ri), xyz))
            if cfg["tropo"] == "tsx":
                p, _ = orbit.interpolate(tt)
                tau_atm.append(dry_tropo_delay_tsx(p, ecef_to_llh(xyz)))
    pulse_times = None
    if prf_dither > 0:
        k = np.arange(npulse)
        # slow PRF stepping (offsets of up to `prf_dither` pulse intervals accumulate and
        # unwind over 97 pulses) plus a small pulse-to-pulse jitter; stays strictly increasing
        jit = np.random.default_rng(seed + 7).uniform(-0.5, 0.5, npulse)
        pulse_times = t_first + (k + prf_dither * np.sin(2 * np.pi * k / 97.0) +
                                 0.1 * min(prf_dither, 1.0) * jit) / prf
        assert np.all(np.diff(pulse_times) > 0)
    rc = np.zeros((npulse, nbins), np.complex64)
    simulate_echoes(rc, in_grid, orbit, targets, fc, bw, fs,
                    tau_atm=np.array(tau_atm) if tau_atm else None, pulse_times=pulse_times)
    if cfg["noise_db"] is not None:
        add_noise(rc, 10.0 ** (cfg["noise_db"] / 20.0), seed)
    kernel = knab_table_kernel(cfg["taps"], bw / fs if not airborne else 0.8, table_size)
    return Scene(name, in_geom, out_geom, rc, dem, fc, float(cfg["ds"]), kernel, cfg["tropo"],
                 targets, bw, fs, pulse_times=pulse_times)
