"""Range compression: the numpy oracle against the reference's own
known answers (tests/cxx/isce3/focus/rangecomp.cpp:86-183, tests/python/extensions/pybind/
focus/rangecomp.py), and -- on a GPU -- isce3_b200's RangeComp against the same known answers
and against the oracle on random data."""
import numpy as np
import pytest

from oracle import rangecomp as orc


def _sinc(x):
    return np.sinc(x)  # sin(pi x)/(pi x), as isce3::math::sinc


def test_oracle_convolve_modes_known_answers():
    chirp = np.ones(5, np.complex64)
    x = np.ones(9, np.complex64)
    want = {orc.FULL: [1, 2, 3, 4, 5, 5, 5, 5, 5, 4, 3, 2, 1], orc.VALID: [5, 5, 5, 5, 5],
            orc.SAME: [3, 4, 5, 5, 5, 5, 5, 4, 3]}
    first = {orc.FULL: 4, orc.VALID: 0, orc.SAME: 2}
    for mode, exp in want.items():
        y = orc.rangecompress(chirp, x, mode)[0]
        assert y.size == orc.output_size(5, 9, mode)
        assert np.max(np.abs(y - np.array(exp))) < 1e-6
        assert orc.first_valid_sample(5, mode) == first[mode]


def test_oracle_chirp_autocorrelation_matches_the_analytic_result():
    chirprate, duration, fs = 100.0, 2.0, 2400.0
    chirp = orc.form_linear_chirp(chirprate, duration, fs)
    assert chirp.size % 2 == 1
    y = orc.rangecompress(chirp, chirp)[0] / fs
    n = y.size
    T = chirp.size / fs
    t = np.arange(n) / fs - 0.5 * (n - 1) / fs
    expected = (T - np.abs(t)) * _sinc(chirprate * t * (T - np.abs(t)))
    k = 200
    assert np.max(np.abs(y[n // 2 - k:n // 2 + k + 1] - expected[n // 2 - k:n // 2 + k + 1])) < 1e-6


def test_next_fast_power():
    assert [orc.next_fast_power(n) for n in (0, 1, 2, 7, 13, 17, 1000, 12288 + 2047)] == \
        [1, 1, 2, 8, 15, 18, 1000, 14400]


# ---- GPU ------------------------------------------------------------------------------------

gpu = pytest.mark.gpu


@gpu
def test_gpu_reference_python_test():
    """tests/python/extensions/pybind/focus/rangecomp.py, same assertions."""
    import isce3_b200.ext.isce3 as isce
    focus = isce.focus
    nchirp = ndata = 1
    batch = 10
    h = np.ones(nchirp, dtype="c8")
    rc = focus.RangeComp(h, ndata, maxbatch=batch)
    assert rc.chirp_size == nchirp and rc.input_size == ndata
    assert rc.mode == focus.RangeComp.Mode.Full
    assert rc.fft_size >= nchirp + ndata - 1 and rc.maxbatch == batch
    assert rc.output_size == nchirp + ndata - 1
    x = np.ones(ndata, dtype="c8")
    y = np.zeros_like(x)
    rc.rangecompress(y, x)
    assert np.allclose(y, x)
    x = np.arange(batch, dtype="c8").reshape((batch, 1))
    y = np.zeros_like(x)
    rc.rangecompress(y, x)
    assert np.allclose(y, x)


@gpu
def test_gpu_convolve_modes_and_chirp_known_answers():
    from isce3_b200.focus import RangeComp, form_linear_chirp
    chirp = np.ones(5, np.complex64)
    x = np.ones(9, np.complex64)
    want = {RangeComp.Mode.Full: ([1, 2, 3, 4, 5, 5, 5, 5, 5, 4, 3, 2, 1], 4),
            RangeComp.Mode.Valid: ([5, 5, 5, 5, 5], 0), RangeComp.Mode.Same: ([3, 4, 5, 5, 5, 5, 5, 4, 3], 2)}
    for mode, (exp, first) in want.items():
        rc = RangeComp(chirp, 9, 1, mode)
        y = np.zeros(rc.output_size, np.complex64)
        rc.rangecompress(y, x)
        assert np.max(np.abs(y - np.array(exp))) < 1e-6
        assert rc.first_valid_sample == first
    chirprate, duration, fs = 100.0, 2.0, 2400.0
    c = form_linear_chirp(chirprate, duration, fs)
    np.testing.assert_array_equal(c, orc.form_linear_chirp(chirprate, duration, fs))
    rc = RangeComp(c, c.size)
    y = np.zeros(rc.output_size, np.complex64)
    rc.rangecompress(y, c)
    y = y / fs
    n = y.size
    T = c.size / fs
    t = np.arange(n) / fs - 0.5 * (n - 1) / fs
    expected = (T - np.abs(t)) * _sinc(chirprate * t * (T - np.abs(t)))
    k = 200
    assert np.max(np.abs(y[n // 2 - k:n // 2 + k + 1] - expected[n // 2 - k:n // 2 + k + 1])) < 1e-6


@gpu
@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("nchirp,ndata,batch", [(481, 3000, 7), (1025, 12288, 16), (33, 20, 3)])
def test_gpu_matches_oracle_on_random_data(mode, nchirp, ndata, batch):
    from isce3_b200.focus import RangeComp
    rng = np.random.default_rng(5)
    chirp = (rng.standard_normal(nchirp) + 1j * rng.standard_normal(nchirp)).astype(np.complex64)
    x = (rng.standard_normal((batch, ndata)) + 1j * rng.standard_normal((batch, ndata))).astype(np.complex64)
    rc = RangeComp(chirp, ndata, maxbatch=batch, mode=RangeComp.Mode(mode))
    assert rc.fft_size == orc.next_fast_power(nchirp + ndata - 1)
    assert rc.output_size == orc.output_size(nchirp, ndata, mode)
    y = np.zeros((batch, rc.output_size), np.complex64)
    rc.rangecompress(y, x)
    want = orc.rangecompress(chirp, x, mode)
    assert np.linalg.norm(y - want) <= 2e-6 * np.linalg.norm(want)
    # a smaller batch through the same object, 1-D call
    y1 = np.zeros(rc.output_size, np.complex64)
    rc.rangecompress(y1, x[2])
    assert np.linalg.norm(y1 - want[2]) <= 2e-6 * np.linalg.norm(want[2])


@gpu
def test_gpu_argument_errors_mirror_the_binding():
    from isce3_b200 import focus
    rc = focus.RangeComp(np.ones(4, np.complex64), 16, maxbatch=2)
    with pytest.raises(ValueError, match="batch size exceeds max batch"):
        rc.rangecompress(np.zeros((3, rc.output_size), np.complex64), np.zeros((3, 16), np.complex64))
    with pytest.raises(ValueError, match="unexpected input length"):
        rc.rangecompress(np.zeros((2, rc.output_size), np.complex64), np.zeros((2, 15), np.complex64))
    with pytest.raises(ValueError, match="unexpected output length"):
        rc.rangecompress(np.zeros((2, 5), np.complex64), np.zeros((2, 16), np.complex64))
    with pytest.raises(ValueError, match="same ndim"):
        rc.rangecompress(np.zeros(rc.output_size, np.complex64), np.zeros((1, 16), np.complex64))
    with pytest.raises(focus.DomainError):
        focus.RangeComp(np.ones(4, np.complex64), 0)
    with pytest.raises(focus.DomainError):
        focus.RangeComp(np.ones(4, np.complex64), 8, maxbatch=0)


@gpu
def test_gpu_range_compressed_swath_stays_in_hbm_for_backprojection(oracle):
    """Raw -> range compression -> backprojection without the swath going back to the host:
    point-target raw echoes (the range-compressed synthetic scene convolved with the chirp,
    so that compressing them gives the chirp's autocorrelation around each target) are
    compressed on the GPU into HBM and focused from there; the result equals focusing the
    host copy of the same compressed data, and matches the CPU reference on it."""
    from testkit import synth
    from isce3_b200.focus import RangeComp, backproject, last_stats
    sc = synth.make_scene("c2", pulses=2048, bins=1024, out_lines=16, out_samples=200, n_targets=1, noise_db=False)
    chirp = orc.form_linear_chirp(20e6 / 20e-6, 20e-6, 24e6)           # 481 samples
    chirp = (chirp / np.sqrt(np.sum(np.abs(chirp) ** 2))).astype(np.complex64)
    # "raw" data: each RC line spread by the chirp (full convolution), so that matched
    # filtering in mode Valid returns lines of the original length
    raw = np.stack([np.convolve(line, chirp) for line in sc.rc]).astype(np.complex64)
    rc = RangeComp(chirp, raw.shape[1], maxbatch=300, mode=RangeComp.Mode.Valid)
    assert rc.output_size == sc.rc.shape[1]
    dev = rc.rangecompress_to_device(raw)
    host = dev.to_host()
    want = orc.rangecompress(chirp, raw, orc.VALID)
    assert np.linalg.norm(host - want) <= 2e-6 * np.linalg.norm(want)
    shape = (sc.out_geometry.grid_length, sc.out_geometry.grid_width)
    args = list(sc.backproject_args())
    out_dev = np.zeros(shape, np.complex64)
    args[1] = dev
    assert backproject(out_dev, *args) is False
    assert last_stats()["h2d_bytes"] == 0
    out_host = np.zeros(shape, np.complex64)
    args[1] = host
    backproject(out_host, *args)
    assert np.linalg.norm(out_dev - out_host) <= 1e-6 * np.linalg.norm(out_host)
    ref = np.zeros(shape, np.complex64)
    oracle.backproject(ref, *args)
    assert np.linalg.norm(out_dev - ref) <= 1e-4 * np.linalg.norm(ref)
    # and the image still shows the target where it was placed
    tg = sc.targets[0]
    assert np.unravel_index(np.argmax(np.abs(out_dev)), shape) == (int(tg.az_index), int(tg.rg_index))
    dev.free()


@gpu
def test_gpu_fused_radiometric_corrections_match_the_workflow_host_pass():
    """set_scaling / patterns: the workflow's host pass over every range-compressed block
    (nisar/workflows/focus.py:1956-1975) -- ``*= deramp_rc[None, :]``, per line
    ``/= np.interp(slant_ranges, pat_ranges, patterns[pulse])``, ``*= slant_ranges / ref_range``
    -- fused into the kernel that writes the output."""
    from isce3_b200.focus import RangeComp
    rng = np.random.default_rng(21)
    nchirp, ndata, batch = 257, 4000, 9
    chirp = np.exp(1j * np.pi * 0.3 * (np.arange(nchirp) - nchirp / 2) ** 2 / nchirp).astype(np.complex64)
    x = (rng.standard_normal((batch, ndata)) + 1j * rng.standard_normal((batch, ndata))).astype(np.complex64)
    rc = RangeComp(chirp, ndata, maxbatch=4, mode=RangeComp.Mode.Valid)
    n = rc.output_size
    plain = np.zeros((4, n), np.complex64)
    rc.rangecompress(plain, x[:4])
    slant = 9.0e5 + 6.25 * np.arange(n)
    deramp = np.exp(1j * 2 * np.pi * 0.013 * np.arange(n))
    pat_ranges = np.linspace(slant[10], slant[-300], 40)  # does not cover the swath: clamped ends
    patterns = ((1.0 + 0.3 * rng.uniform(size=(batch, 40))) *
                np.exp(1j * 0.2 * rng.standard_normal((batch, 40)))).astype(np.complex64)
    ref_range = slant[0]
    want = plain.astype(np.complex128) * deramp[None, :]
    for b in range(4):
        want[b] /= np.interp(slant, pat_ranges, patterns[b])
    want *= (slant / ref_range)[None, :]
    rc.set_scaling(column_scale=deramp * slant / ref_range, slant_ranges=slant, pattern_ranges=pat_ranges)
    got = np.zeros((4, n), np.complex64)
    rc.rangecompress(got, x[:4], patterns=patterns[:4])
    assert np.max(np.abs(got - want)) <= 2e-6 * np.max(np.abs(want))
    # column factors only; and the resident route (chunks of maxbatch) with per-line patterns
    rc.set_scaling(column_scale=deramp)
    rc.rangecompress(got, x[:4])
    assert np.max(np.abs(got - plain * deramp[None, :])) <= 2e-6 * np.max(np.abs(plain))
    rc.set_scaling(column_scale=deramp * slant / ref_range, slant_ranges=slant, pattern_ranges=pat_ranges)
    dev = rc.rangecompress_to_device(x, patterns=patterns)
    full = dev.to_host()
    dev.free()
    rc.rangecompress(got, x[:4], patterns=patterns[:4])
    np.testing.assert_array_equal(full[:4], got)
    assert np.isfinite(full).all()
    # cleared again: plain output
    rc.set_scaling()
    rc.rangecompress(got, x[:4])
    np.testing.assert_array_equal(got, plain)
    with pytest.raises(ValueError):
        rc.rangecompress(got, x[:4], patterns=patterns[:4])
