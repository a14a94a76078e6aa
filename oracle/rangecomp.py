"""TEST INFRASTRUCTURE -- numpy restatement of the reference's range compression, used only
by tests/ and scripts/ as the checker of the GPU RangeComp.

Follows cxx/isce3/focus/RangeComp.cpp: getOutputSize :10-20, formRangeReference :22-39 (matched
filter = time-reversed complex conjugate of the chirp, zero-padded to the FFT length),
rangecompress :90-131 (zero-pad, FFT, multiply by the reference spectrum and 1/N, inverse FFT,
crop by mode), firstValidSample :78-88; FFT length cxx/isce3/fft/FFTUtil.icc:39-75
(nextFastPower: smallest 2^a 3^b 5^c >= n); formLinearChirp cxx/isce3/focus/Chirp.cpp:10-58.
Pinned by the reference's own known answers (tests/cxx/isce3/focus/rangecomp.cpp:86-183) in
tests/test_rangecomp.py.
"""
import math

import numpy as np

FULL, VALID, SAME = 0, 1, 2


def next_fast_power(n: int) -> int:
    if n <= 1:
        return 1
    best = None
    x5 = 1
    while x5 < 5 * n:
        x3 = x5
        while x3 < 3 * n:
            m = x3
            while m < n:
                m *= 2
            best = m if best is None else min(best, m)
            x3 *= 3
        x5 *= 5
    return best


def output_size(m: int, n: int, mode: int) -> int:
    if mode == FULL:
        return m + n - 1
    if mode == VALID:
        return max(m, n) - min(m, n) + 1
    return n


def first_valid_sample(chirp_size: int, mode: int) -> int:
    return chirp_size - 1 if mode == FULL else 0 if mode == VALID else chirp_size // 2


def rangecompress(chirp, x, mode=FULL):
    chirp = np.asarray(chirp, np.complex64)
    x = np.atleast_2d(np.asarray(x, np.complex64))
    m, n = chirp.size, x.shape[1]
    nfft = next_fast_power(m + n - 1)
    ref = np.zeros(nfft, np.complex64)
    ref[:m] = np.conj(chirp[::-1])
    REF = np.fft.fft(ref.astype(np.complex128))
    X = np.fft.fft(x.astype(np.complex128), nfft, axis=1)
    y = np.fft.ifft(X * REF[None, :], axis=1)
    off = 0 if mode == FULL else m - 1 if mode == VALID else m // 2
    return y[:, off:off + output_size(m, n, mode)].astype(np.complex64)


def form_linear_chirp(chirprate, duration, samplerate, centerfreq=0.0, amplitude=1.0, phi=0.0):
    size = int(math.floor(samplerate * duration + 1))
    if size % 2 == 0:
        size += 1
    spacing = 1.0 / samplerate
    tau = -0.5 * (size - 1) * spacing + spacing * np.arange(size)
    phase = phi + 2.0 * np.pi * (centerfreq + 0.5 * chirprate * tau) * tau
    return (amplitude * np.exp(1j * phase)).astype(np.complex64)
