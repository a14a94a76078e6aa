"""Point-target impulse-response metrics (peak location, -3 dB width, PSLR, ISLR) for the IRF
parity gate of the tests.  Test infrastructure, not product code.

The numbers are defined by the reference's ``isce3.cal.point_target_info.analyze_point_target``
(python/packages/isce3/cal/point_target_info.py: chip extraction :31-66, carrier estimate
:68-80, Fourier oversampling :108-144, -3 dB width :147-167, null search :246-306, PSLR / ISLR
:308-455, driver :506-612 / :679-916); the parity gate compares those metrics between two
images, so this module has to reproduce the reference's definitions -- its radar-coordinate,
rectangular-window branch -- not merely compute "an" IRF.  The reference module needs the
compiled isce3 extension and cannot be imported where the tests run, hence this independent
implementation; tests/golden/point_target_golden.npz holds outputs of the reference's own
functions (executed from /root/reference by tests/golden/make_point_target_golden.py) and
tests/test_host_api.py checks this module against them to 1e-9.
"""
from __future__ import annotations

import numpy as np

C64 = np.complex64


class MissingNull(Exception):
    """The main lobe has no null on one side inside the analysed cut."""


# ---- Fourier-domain oversampling of a square chip ------------------------------------------

def lag1_carrier(z):
    """Mean phase advance per sample along x (columns) and y (rows), from lag-1 products."""
    along_x = np.sum(z[:, 1:] * np.conj(z[:, :-1]))
    along_y = np.sum(z[1:, :] * np.conj(z[:-1, :]))
    return float(np.angle(along_x)), float(np.angle(along_y))


def _demodulate(z, wx, wy):
    """z[r, c] * exp(j (wx c + wy r)), rounded to complex64 after each axis like an in-place
    product on a complex64 array."""
    z = (z * np.exp(1j * wx * np.arange(z.shape[1]))[None, :]).astype(C64)
    return (z * np.exp(1j * wy * np.arange(z.shape[0]))[:, None]).astype(C64)


def _pad_axis(spec, axis, factor):
    """Zero-pad a DFT along one axis to ``factor`` times its length; the Nyquist bin of the
    even-length input is shared equally between +Nyquist and -Nyquist of the output."""
    n = spec.shape[axis]
    half = n // 2
    spec = np.moveaxis(spec, axis, 0)
    out = np.zeros((n * factor,) + spec.shape[1:], dtype=spec.dtype)
    out[:half] = spec[:half]              # DC and positive frequencies
    out[-(half - 1):] = spec[half + 1:]   # negative frequencies above -Nyquist
    nyq = spec.dtype.type(0.5) * spec[half]
    out[half] = nyq
    out[-half] = nyq
    return np.moveaxis(out, 0, axis)


def oversample(chip, nov, baseband=False, return_slopes=False):
    """Band-limited interpolation of an even-sized square complex chip by the integer factor
    ``nov``: remove the carrier, zero-pad the 2-D spectrum, put the (scaled) carrier back."""
    rows, cols = chip.shape
    if rows != cols or rows % 2:
        raise ValueError("chip must be square with an even size")
    wx = wy = 0.0
    work = np.asarray(chip)
    if not baseband:
        wx, wy = lag1_carrier(work)
        work = _demodulate(work, -wx, -wy)
    spec = np.fft.fft2(work)
    big = _pad_axis(_pad_axis(spec, 0, nov), 1, nov)
    fine = np.fft.ifft2(big)
    fine *= nov ** 2
    if not baseband:
        fine = (fine * np.exp(1j * (wx / nov) * np.arange(fine.shape[1]))[None, :]).astype(fine.dtype)
        fine = (fine * np.exp(1j * (wy / nov) * np.arange(fine.shape[0]))[:, None]).astype(fine.dtype)
    fine = np.asarray(fine, dtype=work.dtype)
    return (fine, wx, wy) if return_slopes else fine


# ---- metrics of a 1-D cut through the peak -----------------------------------------------

def half_power_width(cut, spacing=1.0):
    """Distance between the samples closest to half the peak power on either side of the peak."""
    power = np.abs(cut) ** 2
    peak = int(np.nanargmax(power))
    excess = power - 0.5 * power[peak]
    if excess[0] >= 0.0 or excess[-1] >= 0.0:
        return spacing * len(cut)  # the cut never drops below half power
    dist = np.abs(excess)
    left = int(np.nanargmin(dist[:peak]))
    right = peak + int(np.nanargmin(dist[peak:]))
    return spacing * (right - left)


def _first_minimum(positions, values, skip=0):
    """Walking away from the peak (values[0] is the peak): position of the (skip+1)-th local
    minimum, ignoring plateaus (runs of equal values count as one sample)."""
    if np.any(values > values[0]):
        raise ValueError("cut does not start at its maximum")
    keep = np.flatnonzero(np.diff(values) != 0.0)  # first sample of every run
    pos, val = positions[keep], values[keep]
    falling = np.sign(np.diff(val))
    found = 0
    for k in range(1, len(falling)):
        if falling[k - 1] < 0 and falling[k] > 0:  # slope turns from down to up at sample k
            if found == skip:
                return pos[k]
            found += 1
    raise MissingNull("no null found beside the main lobe")


def main_lobe_nulls(power_db, peak):
    idx = np.arange(len(power_db))
    towards_start = slice(peak, 0, -1)
    towards_end = slice(peak, None)
    return (_first_minimum(idx[towards_start], power_db[towards_start]),
            _first_minimum(idx[towards_end], power_db[towards_end]))


def sidelobe_ratios(cut, fs_bw_ratio=1.2, num_sidelobes=10, predict_null=False):
    """(ISLR, PSLR) in dB of a cut through the peak, rectangular-window definitions: the main
    lobe runs between the first nulls (or, with ``predict_null``, +-2 fs/B samples around the
    peak), side lobes over ``num_sidelobes`` main-lobe half-widths on either side."""
    power = np.abs(cut) ** 2
    peak = int(np.nanargmax(power))
    null_lo, null_hi = main_lobe_nulls(10 * np.log10(power), peak)
    if predict_null:
        half = int(np.round(2 * fs_bw_ratio))
        lobe_lo, lobe_hi = peak - half, peak + half
    else:
        lobe_lo, lobe_hi = null_lo, null_hi
        half = peak - null_lo
    reach = int(np.round(num_sidelobes * half))
    side_lo = max(lobe_lo - reach, 0)
    side_hi = min(lobe_hi + reach, len(power) - 1)
    integrated_side = np.nansum(power[side_lo:lobe_lo]) + np.nansum(power[lobe_hi + 1:side_hi + 1])
    islr = 10 * np.log10(integrated_side / np.nansum(power[lobe_lo:lobe_hi + 1]))
    peak_side = max(np.nanmax(power[side_lo:null_lo], initial=-np.inf),
                    np.nanmax(power[null_hi + 1:side_hi + 1], initial=-np.inf))
    pslr = 10 * np.log10(peak_side / np.nanmax(power[null_lo:null_hi + 1]))
    return islr, pslr


# ---- driver -------------------------------------------------------------------------------

def analyze_point_target(slc, i, j, nov=32, chipsize=64, fs_bw_ratio=1.2, num_sidelobes=10,
                         predict_null=False, cuts=False):
    """Point-target attributes around (row i, column j) of a complex image.  Returns
    ``(info, None)`` shaped like the reference's result: ``magnitude``, ``phase`` and, for
    ``azimuth`` / ``range``, ``index``, ``offset`` (samples), ``resolution`` (-3 dB width,
    samples), ``PSLR``, ``ISLR`` (dB), ``phase ramp`` (rad / sample)."""
    nrow, ncol = slc.shape
    if not (0 <= i <= nrow and 0 <= j <= ncol):
        raise ValueError("target location is outside of the image array")
    h = chipsize // 2
    if i < h or i > nrow - h or j < h or j > ncol - h:
        raise RuntimeError("target is too close to image border -- consider reducing chipsize")
    row0, col0 = int(i) - h + 1, int(j) - h + 1
    chip = np.array(slc[row0:row0 + chipsize, col0:col0 + chipsize], dtype=C64)
    fine, wx, wy = oversample(chip, nov, return_slopes=True)
    fine = np.ascontiguousarray(fine)
    pr, pc = np.unravel_index(np.nanargmax(np.abs(fine)), fine.shape)
    top = fine[pr, pc]
    info = {"magnitude": float(np.abs(top)), "phase": float(np.angle(top))}
    axes = (("azimuth", fine[:, pc], row0 + pr / nov, i, wy),
            ("range", fine[pr, :], col0 + pc / nov, j, wx))
    for name, cut, where, nominal, ramp in axes:
        islr, pslr = sidelobe_ratios(cut, nov * fs_bw_ratio, num_sidelobes, predict_null)
        info[name] = {"ISLR": float(islr), "PSLR": float(pslr),
                      "resolution": float(half_power_width(cut, 1.0 / nov)),
                      "index": float(where), "offset": float(where - nominal),
                      "phase ramp": float(ramp)}
        if cuts:
            info[name]["magnitude cut"] = np.abs(cut)
            info[name]["phase cut"] = np.angle(cut)
    return info, None
