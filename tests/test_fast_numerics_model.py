"""CPU model of the fast accumulation kernel's arithmetic (csrc/accumulate_fast.cu), pixel by
pixel in numpy, against the CPU oracle -- no GPU needed.  It follows the kernel's numerical
design, not its code: exact FP64 carrier phase only at 64-pulse segment boundaries, a cubic
through the four surrounding boundaries evaluated in FP32 inside the segment, the sample
coordinate as an FP32 affine function of that phase, magic-number integer/fraction split,
per-tap polynomial weights from the library's own host-side fit (i3b_fit_tap_polynomials),
FP32 tap sums and 16-pulse tile sums, FP64 across tiles.  Shows on the CPU that this design
meets the parity gate (<= 1e-4 relative) with two orders of magnitude to spare, and catches
regressions of the fit."""
import ctypes
import math

import numpy as np
import pytest

from isce3_b200 import _capi, core
from testkit import synth

C0 = 299792458.0
SEG, TK = 64, 16
MAGIC32 = np.float32(12582912.0)


def _fit(kernel):
    L = _capi.load_library()
    fl = _capi.Flattened()
    k = _capi.flatten_kernel(kernel, fl)
    fit = _capi.TapPolyFit()
    assert L.i3b_fit_tap_polynomials(ctypes.byref(k), ctypes.byref(fit)) == 0 and fit.supported
    K = fit.taps
    ev = np.array([[fit.even[m][i] for i in range(4)] for m in range((K + 1) // 2)], np.float32)
    od = np.array([[fit.odd[m][i] for i in range(4)] for m in range((K + 1) // 2)], np.float32)
    return K, ev, od


def _weights(f, K, ev, od):
    """w[m](f) for float32 f (vector over pulses): E_m(f^2) +/- f O_m(f^2), Horner in FP32."""
    h = f * f
    w = np.zeros((K, f.size), np.float32)
    for m in range((K + 1) // 2):
        e = np.full_like(f, ev[m, 3])
        o = np.full_like(f, od[m, 3])
        for i in (2, 1, 0):
            e = e * h + ev[m, i]
            o = o * h + od[m, i]
        w[m] = e + f * o
        w[K - 1 - m] = e - f * o
    return w


def model_pixel(sc, x, tau_atm, kstart, kstop, pos, vel, K, ev, od):
    """Fast-kernel arithmetic for one pixel; pos/vel include 64 pulses of padding in front and
    128 behind the input grid (index k + 64)."""
    fc = sc.fc
    g_in = sc.in_geometry.radar_grid
    dtau = 2.0 * g_in.range_pixel_spacing / C0
    swst = 2.0 * g_in.starting_range / C0
    G, U0 = 1.0 / (fc * dtau), swst / dtau
    A = 2.0 / (np.sum(vel * vel, axis=1) - C0 * C0)
    xx = float(x @ x)
    r2 = xx + np.sum(pos * pos, axis=1) - 2.0 * (pos @ x)
    cyc = (-fc * A * C0) * np.sqrt(r2) + (vel @ x) * fc * A + fc * tau_atm - fc * A * np.sum(pos * vel, axis=1)

    def Y(k):  # exact carrier phase in cycles at pulse k (FP64)
        return cyc[k + 64]

    lowoff = -(K // 2) if K & 1 else 1 - K // 2
    shift = 0.5 if K & 1 else 0.0
    Gr = np.float32(G / (2 * math.pi))
    acc = 0.0 + 0.0j
    nr = sc.rc.shape[1]
    for b in range(kstart, kstop, SEG):
        y0, y1, y2, y3 = Y(b - SEG), Y(b), Y(b + SEG), Y(b + 2 * SEG)
        d1, d2 = y2 - y1, (y2 - y1) - (y1 - y0)
        d3 = ((y3 - y2) - (y2 - y1)) - d2
        two_pi = 2 * math.pi
        c1 = np.float32(two_pi * (d1 - 0.5 * d2 - d3 / 6.0) / SEG)
        c2 = np.float32(two_pi * (0.5 * d2) / SEG ** 2)
        c3 = np.float32(two_pi * (d3 / 6.0) / SEG ** 3)
        ang0 = np.float32(two_pi * (y1 - np.rint(y1)))
        uh = y1 * G + (shift - U0)
        ufl = math.floor(uh)
        f0m = np.float32(np.float32(uh - ufl) - np.float32(0.5)) - Gr * ang0
        i0 = int(ufl) + lowoff
        n = min(SEG, kstop - b)
        j = np.arange(n, dtype=np.float32)
        ang = ((c3 * j + c2) * j + c1) * j + ang0
        g = ang * Gr + f0m
        m = (g + MAGIC32).astype(np.float32)
        r = (m - MAGIC32).astype(np.float32)
        f = (g - r).astype(np.float32)
        low = i0 + r.astype(np.int64)
        w = _weights(f, K, ev, od)
        idx = low[None, :] + np.arange(K)[:, None]
        ok = (idx >= 0) & (idx < nr)
        d = np.where(ok, sc.rc[np.arange(b, b + n)[None, :], np.clip(idx, 0, nr - 1)], 0).astype(np.complex64)
        a = np.zeros(n, np.complex64)
        for t in range(K):  # FP32 tap sum in index order
            a = (a + w[t].astype(np.complex64) * d[t]).astype(np.complex64)
        z = (a * (np.cos(ang) + 1j * np.sin(ang)).astype(np.complex64)).astype(np.complex64)
        for t0 in range(0, n, TK):  # FP32 within a pulse tile, FP64 across tiles
            s = np.complex64(0)
            for v in z[t0:t0 + TK]:
                s = np.complex64(s + v)
            acc += complex(s)
    return acc


@pytest.mark.parametrize("name,kw", [("c2", dict(pulses=6144, bins=1024, out_lines=9, out_samples=33, n_targets=1)),
                                      ("c5", dict(pulses=6144, bins=1024, out_lines=9, out_samples=33, n_targets=1,
                                                  taps=8))])
def test_model_of_the_fast_kernel_meets_the_gate(oracle, name, kw):
    sc = synth.make_scene(name, **kw)
    K, ev, od = _fit(sc.kernel)
    og, ig = sc.out_geometry, sc.in_geometry
    orbit = ig.orbit
    N = ig.grid_length
    dt = 1.0 / ig.radar_grid.prf
    t0 = ig.radar_grid.sensing_start
    # pulses -64 .. N+127 (orbit extrapolation is not needed: the scenes keep apertures interior)
    tk = t0 + (np.arange(-64, N + 128)) * dt
    tk = np.clip(tk, orbit.start_time, orbit.end_time)
    pos, vel = synth.interpolate_orbit_many(orbit, tk)
    wvl = C0 / sc.fc
    ref = np.zeros((og.grid_length, og.grid_width), np.complex64)
    oracle.backproject(ref, *sc.backproject_args())
    model = np.zeros_like(ref)
    rows, cols = [0, 4, 8], [0, 7, 16, 25, 32]
    for j in rows:
        for i in cols:
            t = float(og.sensing_time[j])
            r = float(og.slant_range[i])
            ok, x = oracle.rdr2geo_bracket(t, r, 0.0, og.orbit, sc.dem, wvl, int(og.look_side))
            assert ok == 1  # converged
            p, v = orbit.interpolate(t)
            llh = oracle.xyz_to_llh(x)
            tau_atm = oracle.dry_tropo_tsx(p, llh) if sc.dry_tropo_model == "tsx" else 0.0
            L = wvl * r * (np.linalg.norm(p) / np.linalg.norm(x)) / (2.0 * sc.ds)
            T = L / np.linalg.norm(v)
            kstart = max(int(math.floor((t - 0.5 * T - t0) / dt)), 0)
            kstop = min(int(math.ceil((t + 0.5 * T - t0) / dt)), N)
            assert kstart >= 64 and kstop <= N - 128, "scene must keep the aperture interior"
            model[j, i] = model_pixel(sc, np.asarray(x), tau_atm, kstart, kstop, pos, vel, K, ev, od)
    sel = np.ix_(rows, cols)
    rel = np.linalg.norm(model[sel] - ref[sel]) / np.linalg.norm(ref[sel])
    assert rel <= 2e-5, rel          # gate is 1e-4; the GPU kernel measures 2e-6 .. 9e-6
    peak = (int(sc.targets[0].az_index), int(sc.targets[0].rg_index))
    if peak[0] in rows and peak[1] in cols:
        dphi = abs(np.angle(model[peak] * np.conj(ref[peak])))
        assert dphi <= 1e-3
