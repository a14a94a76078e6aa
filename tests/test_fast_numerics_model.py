"""CPU model of the fast accumulation kernel's arithmetic (csrc/accumulate_fast.cu), pixel by
pixel in numpy, against the CPU oracle -- no GPU needed.  It follows the kernel's numerical
design, not its code: exact FP64 carrier phase only at 64-pulse segment boundaries, a cubic
through the four surrounding boundaries evaluated in FP32 inside the segment, the sample
coordinate as an FP32 affine function of that phase, magic-number integer/fraction split,
per-tap polynomial weights from the library's own host-side fit (i3b_fit_tap_polynomials),
FP32 tap sums and 16-pulse tile sums, FP64 across tiles.  Shows on the CPU that this design
meets the parity gate (<= 1e-4 relative) with two orders of magnitude to spare, and catches
regressions of the fit.

Two models: the round-1 arithmetic (segments from the aperture start, unreduced phase) and the
round-2 arithmetic the kernel runs now (absolute anchors, runs of 8 pulses re-centred with whole
turns removed, steady runs) -- the latter reproduces what the GPU measures (C2 point targets:
5.2e-7 here, 4.8e-7 ... 5.4e-7 on the B200; steady runs 96.9 % here, 96.7 % in the ncu profile)."""
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


# ---- round-2 arithmetic: absolute anchors, runs of 8 pulses, phase reduced per run -------------

SUB = 8
INV_TWO_PI_F = np.float32(0.15915494309189535)
TWO_PI_HI_F = np.float32(6.28125)
TWO_PI_LO_F = np.float32(0.0019353071795864769)


def _f32(x):
    return np.float32(x)


def _fma(a, b, c):
    """float32 fused multiply-add (the product of two float32 is exact in float64)."""
    return np.float32(np.float64(a) * np.float64(b) + np.float64(c))


def fast_segment(wavelength, prf, v_max, r_min):
    """accumulate_fast.cu fast_segment(): 128-pulse segments where the cubic through four exact
    phase values stays within 1.5e-6 rad, else 64."""
    d4 = 1.5 * (4.0 * math.pi / wavelength) * 3.0 * v_max ** 4 / r_min ** 3
    return 128 if 0.0234 * (128.0 / prf) ** 4 * d4 <= 1.5e-6 else 64


def model_pixel_r2(sc, x, tau_atm, kstart, kstop, pos, vel, K, ev, od, seg, pad, reduce_per_run=True):
    """Round-2 fast-kernel arithmetic for one pixel (csrc/accumulate_fast.cu): geometry segments
    and 16-pulse tiles anchored at ABSOLUTE pulse indices, runs of 8 pulses whose phase
    polynomial is re-centred on the run with whole turns removed (RunPoly: two-product of c1 j,
    Cody-Waite 2 pi), steady runs (window position proven fixed: quadratic phase, fraction from
    one FMA) or per-pulse rounding, FP32 sums within a tile and FP64 across tiles.
    ``reduce_per_run=False`` evaluates the segment cubic unreduced instead: here that costs a
    factor of two; on the GPU, where the SFU's own (truncating) range reduction of a phase of
    hundreds of radians adds to it, it measured 6e-5 ... 8e-5 on noise-like scenes.  sin/cos
    are exact in this model."""
    fc = sc.fc
    g_in = sc.in_geometry.radar_grid
    dtau = 2.0 * g_in.range_pixel_spacing / C0
    swst = 2.0 * g_in.starting_range / C0
    G, U0 = 1.0 / (fc * dtau), swst / dtau
    A = 2.0 / (np.sum(vel * vel, axis=1) - C0 * C0)
    xx = float(x @ x)
    r2 = xx + np.sum(pos * pos, axis=1) - 2.0 * (pos @ x)
    cyc = (-fc * A * C0) * np.sqrt(r2) + (vel @ x) * fc * A + fc * tau_atm - fc * A * np.sum(pos * vel, axis=1)
    lowoff = -(K // 2) if K & 1 else 1 - K // 2
    shift = 0.5 if K & 1 else 0.0
    two_pi = 2 * math.pi
    Gr, Gf = _f32(G / two_pi), _f32(G)
    nr = sc.rc.shape[1]
    acc = 0.0 + 0.0j
    n_runs = n_steady = 0
    k_tile0 = (kstart // TK) * TK
    tile_sum = np.complex64(0)
    for kr in range((kstart // SUB) * SUB, kstop, SUB):
        b = (kr // seg) * seg
        if kr == (kstart // SUB) * SUB or kr == b:
            # ---- segment set-up (exact FP64 at the four boundaries around the segment) ----
            y0, y1, y2, y3 = (cyc[k + pad] for k in (b - seg, b, b + seg, b + 2 * seg))
            d1, d2 = y2 - y1, (y2 - y1) - (y1 - y0)
            d3 = ((y3 - y2) - (y2 - y1)) - d2
            c1d = two_pi * (d1 - 0.5 * d2 - d3 / 6.0) / seg
            c1 = _f32(c1d)
            c1lo = _f32(c1d - float(c1))
            c2 = _f32(two_pi * (0.5 * d2) / seg ** 2)
            c3 = _f32(two_pi * (d3 / 6.0) / seg ** 3)
            a0 = _f32(two_pi * (y1 - np.rint(y1)))
            uh = y1 * G + (shift - U0)
            ufl = math.floor(uh)
            f0 = _f32(_f32(uh - ufl) - _f32(0.5))
            i0rel = int(ufl) + lowoff
            f0m = _fma(-Gr, a0, f0)
            curv = _f32(2.0) * abs(c2) + _f32(6.0 * seg) * abs(c3)
            flim = _f32(0.5) - _f32(1e-5) - abs(Gr) * curv * _f32((SUB - 1) ** 2 / 8.0)
            if abs(c3) * _f32(0.0481125 * (SUB - 1) ** 3) > _f32(2e-6):
                flim = _f32(-1.0)
        js = _f32(kr - b)
        if reduce_per_run:
            nq = _f32(c1 * -js)
            e = _fma(c1, js, nq)
            e = _fma(c1lo, js, e)
            t = _fma(nq, -INV_TWO_PI_F, MAGIC32)
            n = _f32(t - MAGIC32)
            rneg = _fma(n, TWO_PI_HI_F, nq)
            rest = _fma(_fma(c3, js, c2), _f32(js * js), _f32(a0 + e))
            rest = _fma(n, -TWO_PI_LO_F, rest)
            A0 = _f32(rest - rneg)
            c3x3, c2x2 = _f32(c3 * _f32(3.0)), _f32(c2 + c2)
            A2 = _fma(c3x3, js, c2)
            A1 = _fma(_fma(c3x3, js, c2x2), js, c1)
            A3 = c3
            f0m_run = _fma(n, Gf, f0m)
        xs = np.arange(SUB, dtype=np.float32)
        ks = kr + np.arange(SUB)
        inside = (ks >= kstart) & (ks < kstop)
        n_runs += 1
        if reduce_per_run:
            def cubic(xv):
                return _fma(_fma(_fma(A3, xv, A2), xv, A1), xv, A0)
            g0 = _fma(A0, Gr, f0m_run)
            ge = _fma(cubic(_f32(SUB - 1)), Gr, f0m_run)
            tt = _f32(_f32(g0 + MAGIC32) - MAGIC32)
            steady = bool(inside.all()) and max(abs(_f32(g0 - tt)), abs(_f32(ge - tt))) <= flim
            if steady:
                n_steady += 1
                A1q = _fma(A3, _f32(-0.5 * (SUB - 1) ** 2), A1)
                A2q = _fma(A3, _f32(1.5 * (SUB - 1)), A2)
                ang = np.array([A0 if xv == 0 else _fma(_fma(A2q, xv, A1q), xv, A0) for xv in xs], np.float32)
                fbase = _f32(f0m_run - tt)
                f = np.array([_fma(a, Gr, fbase) for a in ang], np.float32)
                low = np.full(SUB, i0rel + int(tt), np.int64)
            else:
                ang = np.array([cubic(xv) for xv in xs], np.float32)
                g = np.array([_fma(a, Gr, f0m_run) for a in ang], np.float32)
                r = ((g + MAGIC32).astype(np.float32) - MAGIC32).astype(np.float32)
                f = (g - r).astype(np.float32)
                low = i0rel + r.astype(np.int64)
        else:
            j = (js + xs).astype(np.float32)
            ang = np.array([_fma(_fma(_fma(c3, jv, c2), jv, c1), jv, a0) for jv in j], np.float32)
            g = np.array([_fma(a, Gr, f0m) for a in ang], np.float32)
            r = ((g + MAGIC32).astype(np.float32) - MAGIC32).astype(np.float32)
            f = (g - r).astype(np.float32)
            low = i0rel + r.astype(np.int64)
        w = _weights(f, K, ev, od)
        idx = low[None, :] + np.arange(K)[:, None]
        ok = (idx >= 0) & (idx < nr)
        rows_ = np.clip(ks, 0, sc.rc.shape[0] - 1)
        d = np.where(ok, sc.rc[rows_[None, :], np.clip(idx, 0, nr - 1)], 0).astype(np.complex64)
        a = np.zeros(SUB, np.complex64)
        for tp in range(K):
            a = (a + w[tp].astype(np.complex64) * d[tp]).astype(np.complex64)
        a64 = ang.astype(np.float64)
        z = (a * (np.cos(a64) + 1j * np.sin(a64)).astype(np.complex64)).astype(np.complex64)
        for i_ in range(SUB):
            if inside[i_]:
                tile_sum = np.complex64(tile_sum + z[i_])
        if (kr + SUB) % TK == 0 or kr + SUB >= kstop:  # tile boundary: FP32 -> FP64
            acc += complex(tile_sum)
            tile_sum = np.complex64(0)
    return acc, n_steady / max(n_runs, 1)


def _pixels_of(sc, oracle, rows, cols, pad_lo, pad_hi):
    og, ig = sc.out_geometry, sc.in_geometry
    orbit = ig.orbit
    N = ig.grid_length
    dt = 1.0 / ig.radar_grid.prf
    t0 = ig.radar_grid.sensing_start
    tk = np.clip(t0 + np.arange(-pad_lo, N + pad_hi) * dt, orbit.start_time, orbit.end_time)
    pos, vel = synth.interpolate_orbit_many(orbit, tk)
    wvl = C0 / sc.fc
    out = []
    for j in rows:
        for i in cols:
            t = float(og.sensing_time[j])
            r = float(og.slant_range[i])
            ok, x = oracle.rdr2geo_bracket(t, r, 0.0, og.orbit, sc.dem, wvl, int(og.look_side))
            assert ok == 1
            p, v = orbit.interpolate(t)
            llh = oracle.xyz_to_llh(x)
            tau_atm = oracle.dry_tropo_tsx(p, llh) if sc.dry_tropo_model == "tsx" else 0.0
            L = wvl * r * (np.linalg.norm(p) / np.linalg.norm(x)) / (2.0 * sc.ds)
            T = L / np.linalg.norm(v)
            kstart = max(int(math.floor((t - 0.5 * T - t0) / dt)), 0)
            kstop = min(int(math.ceil((t + 0.5 * T - t0) / dt)), N)
            assert kstart >= pad_lo and kstop <= N - pad_hi, "scene must keep the aperture interior"
            out.append((j, i, np.asarray(x), tau_atm, kstart, kstop))
    return out, pos, vel


@pytest.mark.parametrize("name,kw,bound", [
    ("c2", dict(pulses=6144, bins=1024, out_lines=9, out_samples=33, n_targets=1), 5e-6),
    ("c2", dict(pulses=6144, bins=1024, out_lines=9, out_samples=33, n_targets=1, noise_db=20.0), 1.5e-5),
    ("c5", dict(pulses=6144, bins=1024, out_lines=9, out_samples=33, n_targets=1, taps=8), 2e-5),
    ("c5", dict(pulses=6144, bins=1024, out_lines=9, out_samples=33, n_targets=1, taps=8, noise_db=20.0), 4e-5),
])
def test_round2_model_runs_of_eight_with_per_run_reduction(oracle, name, kw, bound):
    """The round-2 arithmetic in numpy against the reference CPU code, on point-target and on
    noise-like scenes (where per-pulse phase errors do not average out: the case that exposed
    the unreduced FP32 phase).  Gate 1e-4; the GPU measures 5e-7 / 3e-6 (C2) and 5e-6 / 1.6e-5
    (airborne)."""
    sc = synth.make_scene(name, **kw)
    K, ev, od = _fit(sc.kernel)
    g = sc.in_geometry.radar_grid
    og = sc.out_geometry
    orbit = sc.in_geometry.orbit
    v = float(np.linalg.norm(orbit.interpolate(orbit.start_time + 1.0)[1]))
    seg = fast_segment(C0 / sc.fc, g.prf, v, g.starting_range)
    assert seg == (64 if name == "c5" else 128)
    ref = np.zeros((og.grid_length, og.grid_width), np.complex64)
    oracle.backproject(ref, *sc.backproject_args())
    rows, cols = [0, 4, 8], [0, 7, 16, 25, 32]
    pix, pos, vel = _pixels_of(sc, oracle, rows, cols, 2 * seg, 3 * seg)
    model = np.zeros_like(ref)
    plain = np.zeros_like(ref)
    shares = []
    for j, i, x, tau_atm, kstart, kstop in pix:
        model[j, i], share = model_pixel_r2(sc, x, tau_atm, kstart, kstop, pos, vel, K, ev, od, seg, 2 * seg)
        shares.append(share)
        plain[j, i], _ = model_pixel_r2(sc, x, tau_atm, kstart, kstop, pos, vel, K, ev, od, seg, 2 * seg,
                                        reduce_per_run=False)
    sel = np.ix_(rows, cols)
    nrm = np.linalg.norm(ref[sel])
    rel = np.linalg.norm(model[sel] - ref[sel]) / nrm
    rel_plain = np.linalg.norm(plain[sel] - ref[sel]) / nrm
    print(f"{name} noise={kw.get('noise_db')}: seg {seg}, steady runs {np.mean(shares):.3f}, "
          f"rel {rel:.2e} (unreduced phase {rel_plain:.2e})")
    assert rel <= bound, rel
    assert np.mean(shares) > 0.6
