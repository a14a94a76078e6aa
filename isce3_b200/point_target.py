"""Point-target impulse-response metrics (location, -3 dB width, PSLR, ISLR) for the
IRF parity gate.

A numpy restatement of the radar-coordinate branch of the reference's
``isce3.cal.point_target_info.analyze_point_target``
(python/packages/isce3/cal/point_target_info.py:31-66 get_chip, :68-80 frequency
estimate/shift, :108-144 oversample, :147-167 estimate_resolution, :246-306 null
search, :308-455 compute_islr_pslr, :506-612 and :679-916 analyze_point_target).
The reference module cannot be imported here (it pulls in the compiled isce3
extension); tests/golden/point_target_golden.npz pins this restatement against
outputs of the reference's own functions executed from /root/reference with the
isce3 imports stubbed (tests/golden/make_point_target_golden.py).
"""
from __future__ import annotations

import numpy as np


class MissingNull(Exception):
    """Raised when mainlobe null(s) cannot be determined"""


def get_chip(x, i, j, nchip=64):
    i, j = int(i), int(j)
    n2 = nchip // 2
    i0, j0 = i - n2 + 1, j - n2 + 1
    chip = np.array(x[i0:i0 + nchip, j0:j0 + nchip], dtype=np.complex64)
    return i0, j0, chip


def estimate_frequency(z):
    cx = np.sum(z[:, 1:] * z[:, :-1].conj())
    cy = np.sum(z[1:, :] * z[:-1, :].conj())
    return np.angle([cx, cy])


def shift_frequency(z, fx, fy):
    x = np.arange(z.shape[1])
    y = np.arange(z.shape[0])
    z *= np.exp(1j * fx * x)[None, :]
    z *= np.exp(1j * fy * y)[:, None]
    return z


def oversample(x, nov, baseband=False, return_slopes=False):
    m, n = x.shape
    assert m == n and n % 2 == 0
    fx = fy = 0.0
    if not baseband:
        fx, fy = estimate_frequency(x)
        x = shift_frequency(x, -fx, -fy)
    X = np.fft.fft2(x)
    Y = np.zeros((n * nov, n * nov), dtype=X.dtype)
    n2 = n // 2
    Y[:n2, :n2] = X[:n2, :n2]
    Y[-n2:, -n2:] = X[-n2:, -n2:]
    Y[:n2, -n2:] = X[:n2, -n2:]
    Y[-n2:, :n2] = X[-n2:, :n2]
    Y[:n2, n2] = Y[:n2, -n2] = 0.5 * X[:n2, n2]
    Y[-n2:, n2] = Y[-n2:, -n2] = 0.5 * X[-n2:, n2]
    Y[n2, :n2] = Y[-n2, :n2] = 0.5 * X[n2, :n2]
    Y[n2, -n2:] = Y[-n2, -n2:] = 0.5 * X[n2, -n2:]
    Y[n2, n2] = Y[n2, -n2] = Y[-n2, n2] = Y[-n2, -n2] = 0.25 * X[n2, n2]
    y = np.fft.ifft2(Y)
    y *= nov ** 2
    if not baseband:
        y = shift_frequency(y, fx / nov, fy / nov)
    y = np.asarray(y, dtype=x.dtype)
    if return_slopes:
        return y, fx, fy
    return y


def estimate_resolution(x, dt=1.0):
    y = abs(x) ** 2
    i = np.nanargmax(y)
    u = y - 0.5 * y[i]
    if (u[0] >= 0.0) or (u[-1] >= 0.0):
        return dt * len(x)
    z = abs(u)
    ileft = np.nanargmin(z[:i])
    iright = i + np.nanargmin(z[i:])
    return dt * (iright - ileft)


def locate_null(t, half, n=0):
    assert len(t) == len(half)
    if np.any(half > half[0]):
        raise ValueError("IRF not sorted correctly")
    dx = np.diff(half)
    unequal = np.where(dx != 0.0)[0]
    t = t[unequal]
    dx = np.diff(half[unequal])
    nulls = np.where(np.diff(np.sign(dx)) == 2)[0] + 1
    if len(nulls) <= n:
        raise MissingNull("Insufficient nulls found in impulse response.")
    return t[nulls[n]]


def search_first_null_pair(matched_output, mainlobe_peak_idx):
    t = np.arange(len(matched_output))
    left = slice(mainlobe_peak_idx, 0, -1)
    right = slice(mainlobe_peak_idx, None)
    return (locate_null(t[left], matched_output[left]),
            locate_null(t[right], matched_output[right]))


def compute_islr_pslr(data_in_linear, fs_bw_ratio=1.2, num_sidelobes=10, predict_null=False):
    """Rectangular-window case of point_target_info.py:308-455."""
    pwr = np.abs(data_in_linear) ** 2
    pwr_db = 10 * np.log10(pwr)
    peak = np.nanargmax(pwr)
    first_l, first_r = search_first_null_pair(pwr_db, peak)
    if predict_null:
        samples_null_to_peak = int(np.round(2 * fs_bw_ratio))
        main_l, main_r = peak - samples_null_to_peak, peak + samples_null_to_peak
    else:
        main_l, main_r = first_l, first_r
        samples_null_to_peak = peak - first_l
    nside = int(np.round(num_sidelobes * samples_null_to_peak))
    side_l = max(main_l - nside, 0)
    side_r = min(main_r + nside, len(pwr) - 1)
    main = pwr[main_l:main_r + 1]
    side = pwr[np.r_[side_l:main_l, main_r + 1:side_r + 1]]
    islr_db = 10 * np.log10(np.nansum(side) / np.nansum(main))
    pslr_side = pwr[np.r_[side_l:first_l, first_r + 1:side_r + 1]]
    pslr_main = pwr[first_l:first_r + 1]
    pslr_db = 10 * np.log10(np.nanmax(pslr_side) / np.nanmax(pslr_main))
    return islr_db, pslr_db


def analyze_point_target(slc, i, j, nov=32, chipsize=64, fs_bw_ratio=1.2, num_sidelobes=10,
                         predict_null=False, cuts=False):
    """Measure point-target attributes around (row i, column j) of a complex image.

    Returns (dict, None) like the reference: ``magnitude``, ``phase`` and, for
    ``azimuth`` / ``range``: ``index``, ``offset`` (samples), ``resolution`` (-3 dB
    width, samples), ``PSLR``, ``ISLR`` (dB), ``phase ramp``.
    """
    if i > slc.shape[0] or i < 0 or j > slc.shape[1] or j < 0:
        raise ValueError("target location is outside of the image array")
    h = chipsize // 2
    if i < h or i > slc.shape[0] - h or j < h or j > slc.shape[1] - h:
        raise RuntimeError("target is too close to image border -- consider reducing chipsize")
    i0, j0, chip = get_chip(slc, i, j, chipsize)
    up, fx, fy = oversample(chip, nov, return_slopes=True)
    up = np.ascontiguousarray(up)
    k = np.nanargmax(np.abs(up))
    ic, jc = np.unravel_index(k, up.shape)
    cmax = up[ic, jc]
    imax, jmax = i0 + ic / nov, j0 + jc / nov
    az, rg = up[:, jc], up[ic, :]
    out = {"magnitude": float(np.abs(cmax)), "phase": float(np.angle(cmax)),
           "azimuth": {}, "range": {}}
    for name, cut, idx, pos, ramp in (("azimuth", az, imax, i, fy), ("range", rg, jmax, j, fx)):
        islr, pslr = compute_islr_pslr(cut, nov * fs_bw_ratio, num_sidelobes, predict_null)
        out[name] = {"ISLR": float(islr), "PSLR": float(pslr),
                     "resolution": float(estimate_resolution(cut, 1.0 / nov)),
                     "index": float(idx), "offset": float(idx - pos), "phase ramp": float(ramp)}
        if cuts:
            out[name]["magnitude cut"] = np.abs(cut)
            out[name]["phase cut"] = np.angle(cut)
    return out, None
