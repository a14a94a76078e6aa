"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol,
the Python operator mirrors the reference binding's argument checks
(python/extensions/pybind_isce3/cuda/focus/Backproject.cpp:42-89), the point-target
analysis restatement reproduces golden values produced by the reference's own module."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from isce3_b200 import _capi, core, focus, point_target, synth
import isce3_b200.ext.isce3 as isce

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    if not _capi.library_path().exists():
        import __graft_entry__
        __graft_entry__.build()
    return ctypes.CDLL(str(_capi.library_path()))


def test_library_exports_every_declared_symbol(lib):
    header = (ROOT / "include" / "isce3_b200_backproject.h").read_text()
    declared = set(re.findall(r"\b(i3b_[a-z_]+)\s*\(", header))
    assert declared == set(_capi.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layout_matches_header(lib):
    # spot-check the ctypes mirror against the C layout rules of the header
    assert ctypes.sizeof(_capi.RadarGrid) == 5 * 8 + 2 * 8 + 2 * 4
    assert ctypes.sizeof(_capi.Orbit) == 40
    assert ctypes.sizeof(_capi.LUT2d) == 16 + 16 + 5 * 8 + 8
    assert ctypes.sizeof(_capi.Kernel) == 32
    assert ctypes.sizeof(_capi.Geo2RdrBracketParams) == 32
    assert _capi.BackprojectArgs.out.offset == 8
    lib.i3b_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.i3b_version()


def test_no_device_is_a_loud_error_not_a_fallback():
    """Without a GPU the product path must fail (there is no CPU route)."""
    L = _capi.load_library()
    if L.i3b_device_count() > 0:
        pytest.skip("GPU present")
    sc = synth.make_scene("c1", pulses=64, bins=128, out_lines=4, out_samples=8)
    out = np.zeros((4, 8), np.complex64)
    with pytest.raises(focus.CudaError):
        focus.backproject(out, *sc.backproject_args())


def _tiny():
    return synth.make_scene("c1", pulses=64, bins=128, out_lines=4, out_samples=8)


def test_argument_checks_mirror_the_binding():
    sc = _tiny()
    args = sc.backproject_args()
    good = np.zeros((4, 8), np.complex64)
    with pytest.raises(focus.InvalidArgument, match="output array shape"):
        focus.build_args(np.zeros((4, 9), np.complex64), *args)
    with pytest.raises(focus.InvalidArgument, match="2-D"):
        focus.build_args(np.zeros(32, np.complex64), *args)
    bad_in = list(args)
    bad_in[1] = np.zeros((64, 127), np.complex64)
    with pytest.raises(focus.InvalidArgument, match="input signal data shape"):
        focus.build_args(good, *bad_in)
    with pytest.raises(focus.InvalidArgument, match="height array shape"):
        focus.build_args(good, *args, height=np.zeros((4, 7), np.float32))
    with pytest.raises(focus.InvalidArgument, match="dry troposphere"):
        focus.build_args(good, *args[:7], "bogus")
    with pytest.raises(focus.InvalidArgument, match="rdr2geo_bracket keyword"):
        focus.build_args(good, *args[:7], "tsx", {"threshold": 1e-3})
    with pytest.raises(focus.InvalidArgument, match="geo2rdr_bracket keyword"):
        focus.build_args(good, *args[:7], "tsx", {}, {"maxiter": 3})
    with pytest.raises(focus.DomainError, match="batch size"):
        focus.build_args(good, *args, batch=0)
    with pytest.raises(TypeError):
        focus.build_args(good.astype(np.complex128), *args)
    with pytest.raises(TypeError, match="Kernel<float>"):
        focus.build_args(good, *args[:6], core.KnabKernel(9, 0.8))
    fl = focus.build_args(good, *args[:7], "nodelay", {"tol_height": 1e-4, "look_min": 0.1},
                          {"tol_aztime": 1e-6, "time_start": None, "time_end": 130.0}, batch=7)
    a = fl.args
    assert a.dry_tropo_model == 0 and a.batch == 7
    assert a.rdr2geo.tol_height == 1e-4 and a.rdr2geo.look_min == 0.1
    assert a.rdr2geo.look_max == pytest.approx(np.pi / 2)
    assert a.geo2rdr.has_time_start == 0 and a.geo2rdr.has_time_end == 1 and a.geo2rdr.time_end == 130.0
    assert a.kernel.kind == _capi.KERNEL_TABULATED and a.kernel.n == 2048 and a.kernel.width == 9.0
    assert a.in_geometry.grid.length == 64 and a.out_geometry.grid.width == 8
    assert a.in_geometry.ref_epoch_sec == a.out_geometry.ref_epoch_sec


def test_reference_style_construction_reads_like_the_reference_test():
    """Same construction sequence as tests/python/extensions/pybind/focus/backproject.py:97-124."""
    c = isce.core.speed_of_light
    epoch = isce.core.DateTime(2020, 1, 1)
    svs = [isce.core.StateVector(epoch + 10.0 * i, [7e6, 7e3 * 10 * i, 0.0], [0.0, 7e3, 0.0])
           for i in range(6)]
    orbit = isce.core.Orbit(svs, epoch)
    assert orbit.time.first == 0.0 and orbit.time.spacing == 10.0 and orbit.size == 6
    grid = isce.product.RadarGridParameters(12.0, 0.24, 1000.0, 800e3, 5.0, "left", 100, 200, epoch)
    assert grid.az_time_interval == 1e-3 and grid.lookside == isce.core.LookSide.Left
    kernel = isce.core.TabulatedKernelF32(isce.core.KnabKernel(9.0, 20e6 / 24e6), 2048)
    geom = isce.container.RadarGeometry(grid, orbit, isce.core.LUT2d())
    assert geom.sensing_time.size == 100 and geom.slant_range[3] == 800e3 + 15.0
    later = isce.core.DateTime(2020, 1, 1, 0, 0, 5)
    grid2 = isce.product.RadarGridParameters(7.0, 0.24, 1000.0, 800e3, 5.0, "left", 100, 200, later)
    geom2 = isce.container.RadarGeometry(grid2, orbit, isce.core.LUT2d())
    assert geom2.radar_grid.sensing_start == 12.0  # re-based to the orbit epoch
    dem = isce.geometry.DEMInterpolator(50.0)
    assert dem.ref_height == 50.0 and not dem.have_raster and dem.epsg_code == 4326
    assert callable(isce.focus.backproject) and callable(isce.cuda.focus.backproject)
    assert kernel.width == 9.0 and c == 299792458.0


def test_point_target_analysis_matches_reference_golden():
    g = np.load(ROOT / "tests" / "golden" / "point_target_golden.npz")
    z = g["image"]
    for tag, kw in (("a", dict(nov=32, chipsize=64)),
                    ("b", dict(nov=16, chipsize=32, predict_null=True, fs_bw_ratio=1.2))):
        info, _ = point_target.analyze_point_target(z, 80, 80, **kw)
        vals = []
        for axis in ("azimuth", "range"):
            for key in ("index", "offset", "resolution", "PSLR", "ISLR", "phase ramp"):
                vals.append(info[axis][key])
        vals += [info["magnitude"], info["phase"]]
        np.testing.assert_allclose(vals, g[f"info_{tag}"], rtol=1e-9, atol=1e-9)
    ov = point_target.oversample(z[48:112, 48:112].copy(), 4)[::8, ::8]
    np.testing.assert_allclose(ov, g["oversampled_4_decimated"], rtol=1e-5, atol=1e-6)


def test_synthetic_scene_is_deterministic_and_focusable(oracle):
    a = synth.make_scene("c2", pulses=256, bins=512, out_lines=8, out_samples=16, n_targets=1)
    b = synth.make_scene("c2", pulses=256, bins=512, out_lines=8, out_samples=16, n_targets=1)
    np.testing.assert_array_equal(a.rc, b.rc)
    tg = a.targets[0]
    out = np.zeros((8, 16), np.complex64)
    oracle.backproject(out, *a.backproject_args())
    peak = np.unravel_index(np.argmax(np.abs(out)), out.shape)
    assert peak == (int(tg.az_index), int(tg.rg_index))
