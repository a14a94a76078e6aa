"""ctypes mirror of ``include/isce3_b200_backproject.h`` and the loader for the
product library ``libisce3_b200_backproject.so``.

This is the Python side of the drop-in boundary: the reference reaches its CUDA
backprojector through a pybind11 lambda that unpacks isce3 value types
(python/extensions/pybind_isce3/cuda/focus/Backproject.cpp:25-117); here the
same unpacking targets the flat C descriptors.  There is NO CPU fallback: if the
CUDA library is missing or no sm_100 device is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

ABI_VERSION = 3

# status codes (include/isce3_b200_backproject.h)
SUCCESS = 0
FAILED_TO_CONVERGE = 6
WRONG_LOOK_SIDE = 7
INVALID_INTERVAL = 11
EXC_INVALID_ARGUMENT = -1
EXC_RUNTIME_ERROR = -2
EXC_DOMAIN_ERROR = -3
EXC_OVERFLOW_ERROR = -4
EXC_OUT_OF_RANGE = -5
EXC_CUDA_ERROR = -6
EXC_NO_DEVICE = -7
EXC_LENGTH_ERROR = -8

ERROR_STRINGS = {
    0: "Success",
    1: "OrbitInterpSizeError",
    2: "OrbitInterpDomainError",
    3: "OrbitInterpUnknownMethod",
    4: "OutOfBoundsDem",
    5: "InvalidDem",
    6: "FailedToConverge",
    7: "WrongLookSide",
    8: "OutOfBoundsLookup",
    9: "NullDereference",
    10: "InvalidTolerance",
    11: "InvalidInterval",
}

FLAG_FORCE_GENERIC = 1
FLAG_DEVICE_INPUT = 4
FLAG_DEVICE_POINTERS = 2

KERNEL_BARTLETT, KERNEL_LINEAR, KERNEL_KNAB, KERNEL_TABULATED, KERNEL_CHEBY = range(5)
INTERP_SINC, INTERP_BILINEAR, INTERP_BICUBIC, INTERP_NEAREST, INTERP_BIQUINTIC = range(5)
TROPO_NODELAY, TROPO_TSX = 0, 1


class RadarGrid(C.Structure):
    _fields_ = [
        ("sensing_start", C.c_double),
        ("prf", C.c_double),
        ("starting_range", C.c_double),
        ("range_pixel_spacing", C.c_double),
        ("wavelength", C.c_double),
        ("length", C.c_int64),
        ("width", C.c_int64),
        ("look_side", C.c_int32),
        ("_pad", C.c_int32),
    ]


class Orbit(C.Structure):
    _fields_ = [
        ("t0", C.c_double),
        ("dt", C.c_double),
        ("n", C.c_int32),
        ("method", C.c_int32),
        ("pos", C.POINTER(C.c_double)),
        ("vel", C.POINTER(C.c_double)),
    ]


class LUT2d(C.Structure):
    _fields_ = [
        ("have_data", C.c_int32),
        ("bounds_error", C.c_int32),
        ("method", C.c_int32),
        ("_pad", C.c_int32),
        ("length", C.c_int64),
        ("width", C.c_int64),
        ("ref_value", C.c_double),
        ("xstart", C.c_double),
        ("ystart", C.c_double),
        ("dx", C.c_double),
        ("dy", C.c_double),
        ("data", C.POINTER(C.c_double)),
    ]


class RadarGeometry(C.Structure):
    _fields_ = [
        ("grid", RadarGrid),
        ("orbit", Orbit),
        ("doppler", LUT2d),
        ("ref_epoch_sec", C.c_int64),
        ("ref_epoch_frac", C.c_double),
    ]


class DEM(C.Structure):
    _fields_ = [
        ("have_raster", C.c_int32),
        ("epsg", C.c_int32),
        ("method", C.c_int32),
        ("_pad", C.c_int32),
        ("length", C.c_int64),
        ("width", C.c_int64),
        ("ref_height", C.c_double),
        ("xstart", C.c_double),
        ("ystart", C.c_double),
        ("dx", C.c_double),
        ("dy", C.c_double),
        ("data", C.POINTER(C.c_float)),
    ]


class Kernel(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("n", C.c_int32),
        ("width", C.c_double),
        ("bandwidth", C.c_double),
        ("data", C.POINTER(C.c_float)),
    ]


class Rdr2GeoBracketParams(C.Structure):
    _fields_ = [("tol_height", C.c_double), ("look_min", C.c_double), ("look_max", C.c_double)]


class Geo2RdrBracketParams(C.Structure):
    _fields_ = [
        ("tol_aztime", C.c_double),
        ("has_time_start", C.c_int32),
        ("has_time_end", C.c_int32),
        ("time_start", C.c_double),
        ("time_end", C.c_double),
    ]


class BackprojectArgs(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("flags", C.c_uint32),
        ("out", C.c_void_p),
        ("in_", C.c_void_p),
        ("height", C.c_void_p),
        ("out_geometry", RadarGeometry),
        ("in_geometry", RadarGeometry),
        ("dem", DEM),
        ("fc", C.c_double),
        ("ds", C.c_double),
        ("kernel", Kernel),
        ("dry_tropo_model", C.c_int32),
        ("batch", C.c_int32),
        ("rdr2geo", Rdr2GeoBracketParams),
        ("geo2rdr", Geo2RdrBracketParams),
        ("n_devices", C.c_int32),
        ("devices", C.POINTER(C.c_int32)),
        ("range_cor", C.c_void_p),
        ("mantissa_nbits", C.c_int32),
        ("_pad2", C.c_int32),
        ("pulse_times", C.c_void_p),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("pixel_pulses", C.c_double),
        ("ms_total", C.c_double),
        ("ms_h2d", C.c_double),
        ("ms_target_solve", C.c_double),
        ("ms_accumulate", C.c_double),
        ("ms_d2h", C.c_double),
        ("accumulate_launches", C.c_int32),
        ("total_launches", C.c_int32),
        ("used_fast_kernel", C.c_int32),
        ("taps", C.c_int32),
        ("h2d_bytes", C.c_int64),
        ("d2h_bytes", C.c_int64),
        ("pulse_first", C.c_int32),
        ("pulse_last", C.c_int32),
        ("n_devices", C.c_int32),
        ("fast_variant", C.c_int32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Peaks(C.Structure):
    _fields_ = [
        ("fp32_tflops", C.c_double),
        ("fp64_tflops", C.c_double),
        ("sfu_gops", C.c_double),
        ("sm_mhz", C.c_double),
        ("sm_count", C.c_int32),
        ("_pad", C.c_int32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("_")}


# entry points include/isce3_b200_backproject.h declares; tests check that the
# built library exports every one of them.
FIT_MAX_TAP_PAIRS = 17
FIT_MAX_COEF = 4


class RangeCompScaling(C.Structure):
    _fields_ = [
        ("column_scale", C.c_void_p),
        ("slant_ranges", C.c_void_p),
        ("pattern_ranges", C.c_void_p),
        ("n_pattern", C.c_int32),
        ("_pad", C.c_int32),
    ]


class TapPolyFit(C.Structure):
    _fields_ = [
        ("taps", C.c_int32),
        ("degree", C.c_int32),
        ("supported", C.c_int32),
        ("imm_variant", C.c_int32),
        ("max_err", C.c_double),
        ("even", (C.c_float * FIT_MAX_COEF) * FIT_MAX_TAP_PAIRS),
        ("odd", (C.c_float * FIT_MAX_COEF) * FIT_MAX_TAP_PAIRS),
        ("pair_degree", C.c_int32 * FIT_MAX_TAP_PAIRS),
        ("_pad", C.c_int32),
    ]


EXPORTED_SYMBOLS = (
    "i3b_backproject",
    "i3b_plan_create",
    "i3b_plan_execute",
    "i3b_plan_download",
    "i3b_plan_destroy",
    "i3b_last_stats",
    "i3b_last_error",
    "i3b_version",
    "i3b_device_count",
    "i3b_current_device",
    "i3b_measure_peaks",
    "i3b_fit_tap_polynomials",
    "i3b_release_device_memory",
    "i3b_set_device_memory_pool",
    "i3b_blocks_create",
    "i3b_blocks_run",
    "i3b_blocks_destroy",
    "i3b_rdr2geo_bracket_batch",
    "i3b_geo2rdr_bracket_batch",
    "i3b_rangecomp_create",
    "i3b_rangecomp_query",
    "i3b_rangecomp_execute",
    "i3b_rangecomp_execute_to_device",
    "i3b_rangecomp_set_scaling",
    "i3b_rangecomp_execute_scaled",
    "i3b_rangecomp_execute_to_device_scaled",
    "i3b_device_free",
    "i3b_device_to_host",
    "i3b_rangecomp_last_device_ms",
    "i3b_rangecomp_last_error",
    "i3b_rangecomp_destroy",
)

LIB_NAME = "libisce3_b200_backproject.so"
_lib = None


def library_path() -> Path:
    return Path(__file__).resolve().parent / LIB_NAME


def load_library() -> C.CDLL:
    """Load the CUDA backend.  Raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("ISCE3_B200_LIB", library_path()))
    if not path.exists():
        raise RuntimeError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  isce3_b200 has no CPU fallback.")
    lib = C.CDLL(str(path))
    lib.i3b_backproject.argtypes = [C.POINTER(BackprojectArgs)]
    lib.i3b_backproject.restype = C.c_int
    lib.i3b_plan_create.argtypes = [C.POINTER(BackprojectArgs), C.POINTER(C.c_void_p)]
    lib.i3b_plan_create.restype = C.c_int
    lib.i3b_plan_execute.argtypes = [C.c_void_p]
    lib.i3b_plan_execute.restype = C.c_int
    lib.i3b_plan_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.i3b_plan_download.restype = C.c_int
    lib.i3b_plan_destroy.argtypes = [C.c_void_p]
    lib.i3b_plan_destroy.restype = C.c_int
    lib.i3b_last_stats.argtypes = [C.POINTER(Stats)]
    lib.i3b_last_stats.restype = C.c_int
    lib.i3b_last_error.restype = C.c_char_p
    lib.i3b_version.restype = C.c_char_p
    lib.i3b_device_count.restype = C.c_int
    lib.i3b_measure_peaks.argtypes = [C.c_int, C.POINTER(Peaks)]
    lib.i3b_measure_peaks.restype = C.c_int
    lib.i3b_fit_tap_polynomials.argtypes = [C.POINTER(Kernel), C.POINTER(TapPolyFit)]
    lib.i3b_fit_tap_polynomials.restype = C.c_int
    lib.i3b_release_device_memory.restype = C.c_int
    lib.i3b_set_device_memory_pool.argtypes = [C.c_int64]
    lib.i3b_set_device_memory_pool.restype = C.c_int
    lib.i3b_current_device.restype = C.c_int
    lib.i3b_blocks_create.argtypes = [C.POINTER(BackprojectArgs), C.POINTER(C.c_void_p)]
    lib.i3b_blocks_create.restype = C.c_int
    lib.i3b_blocks_run.argtypes = [C.c_void_p, C.c_int32, C.POINTER(RadarGrid), C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_void_p)]
    lib.i3b_blocks_run.restype = C.c_int
    lib.i3b_blocks_destroy.argtypes = [C.c_void_p]
    lib.i3b_blocks_destroy.restype = C.c_int
    lib.i3b_rdr2geo_bracket_batch.argtypes = [
        C.POINTER(Orbit), C.POINTER(DEM), C.c_double, C.c_int32, C.POINTER(Rdr2GeoBracketParams), C.c_int64,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    lib.i3b_rdr2geo_bracket_batch.restype = C.c_int
    lib.i3b_geo2rdr_bracket_batch.argtypes = [
        C.POINTER(Orbit), C.POINTER(LUT2d), C.c_double, C.c_int32, C.POINTER(Geo2RdrBracketParams), C.c_int64,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    lib.i3b_geo2rdr_bracket_batch.restype = C.c_int
    lib.i3b_rangecomp_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.i3b_rangecomp_create.restype = C.c_int
    lib.i3b_rangecomp_query.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.i3b_rangecomp_query.restype = C.c_int
    lib.i3b_rangecomp_execute.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32]
    lib.i3b_rangecomp_execute.restype = C.c_int
    lib.i3b_rangecomp_execute_to_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]
    lib.i3b_rangecomp_execute_to_device.restype = C.c_int
    lib.i3b_rangecomp_set_scaling.argtypes = [C.c_void_p, C.c_void_p]
    lib.i3b_rangecomp_set_scaling.restype = C.c_int
    lib.i3b_rangecomp_execute_scaled.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_void_p]
    lib.i3b_rangecomp_execute_scaled.restype = C.c_int
    lib.i3b_rangecomp_execute_to_device_scaled.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                                           C.POINTER(C.c_void_p)]
    lib.i3b_rangecomp_execute_to_device_scaled.restype = C.c_int
    lib.i3b_device_free.argtypes = [C.c_void_p]
    lib.i3b_device_free.restype = C.c_int
    lib.i3b_device_to_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.i3b_device_to_host.restype = C.c_int
    lib.i3b_rangecomp_last_device_ms.argtypes = [C.c_void_p]
    lib.i3b_rangecomp_last_device_ms.restype = C.c_double
    lib.i3b_rangecomp_last_error.restype = C.c_char_p
    lib.i3b_rangecomp_destroy.argtypes = [C.c_void_p]
    lib.i3b_rangecomp_destroy.restype = C.c_int
    _lib = lib
    return lib


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Flattened:
    """A BackprojectArgs plus the numpy buffers that keep its pointers alive."""

    def __init__(self):
        self.args = BackprojectArgs()
        self.keep = []

    def hold(self, a, dtype):
        a = np.ascontiguousarray(a, dtype=dtype)
        self.keep.append(a)
        return a


def flatten_orbit(orbit, fl: Flattened) -> Orbit:
    pos = fl.hold(orbit.position, np.float64)
    vel = fl.hold(orbit.velocity, np.float64)
    return Orbit(float(orbit.time.first), float(orbit.time.spacing), int(orbit.size),
                 int(orbit.interp_method), _dptr(pos), _dptr(vel))


def flatten_lut2d(lut, fl: Flattened) -> LUT2d:
    d = LUT2d()
    d.have_data = int(lut.have_data)
    d.bounds_error = int(lut.bounds_error)
    d.method = int(lut.interp_method)
    d.ref_value = float(lut.ref_value)
    if lut.have_data:
        data = fl.hold(lut.data, np.float64)
        d.length, d.width = data.shape
        d.xstart, d.ystart = float(lut.x_start), float(lut.y_start)
        d.dx, d.dy = float(lut.x_spacing), float(lut.y_spacing)
        d.data = _dptr(data)
    return d


def flatten_grid(g) -> RadarGrid:
    return RadarGrid(float(g.sensing_start), float(g.prf), float(g.starting_range),
                     float(g.range_pixel_spacing), float(g.wavelength), int(g.length),
                     int(g.width), int(g.lookside), 0)


def flatten_geometry(geom, fl: Flattened) -> RadarGeometry:
    d = RadarGeometry()
    d.grid = flatten_grid(geom.radar_grid)
    d.orbit = flatten_orbit(geom.orbit, fl)
    d.doppler = flatten_lut2d(geom.doppler, fl)
    sec, frac = geom.reference_epoch.epoch_pair()
    d.ref_epoch_sec, d.ref_epoch_frac = sec, frac
    return d


def flatten_dem(dem, fl: Flattened) -> DEM:
    d = DEM()
    d.have_raster = int(dem.have_raster)
    d.epsg = int(dem.epsg_code)
    d.method = int(dem.interp_method)
    d.ref_height = float(dem.ref_height)
    if dem.have_raster:
        data = fl.hold(dem.data, np.float32)
        d.length, d.width = data.shape
        d.xstart, d.ystart = float(dem.x_start), float(dem.y_start)
        d.dx, d.dy = float(dem.delta_x), float(dem.delta_y)
        d.data = _fptr(data)
    return d


def flatten_kernel(kernel, fl: Flattened) -> Kernel:
    d = Kernel()
    d.kind, d.width, d.bandwidth, data = kernel._flatten()
    if data is not None:
        data = fl.hold(data, np.float32)
        d.n = data.size
        d.data = _fptr(data)
    return d
