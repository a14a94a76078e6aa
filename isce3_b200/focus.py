"""``backproject`` with the reference's Python signature, routed to the B200 CUDA
backend through the C-ABI of ``include/isce3_b200_backproject.h``.

Mirrors python/extensions/pybind_isce3/cuda/focus/Backproject.cpp:25-117 (argument
checks, defaults, bool return) and pybind_isce3/focus/Backproject.cpp:29-78 (the
``rdr2geo_params`` / ``geo2rdr_params`` dict parsers).  The exceptions the
reference throws map to Python ones the way pybind11 translates them:
InvalidArgument -> ValueError, DomainError -> ValueError, RuntimeError ->
RuntimeError, OverflowError -> OverflowError, OutOfRange -> IndexError.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _capi
from ._capi import Flattened


class InvalidArgument(ValueError):
    """isce3::except::InvalidArgument (std::invalid_argument)"""


class DomainError(ValueError):
    """isce3::except::DomainError (std::domain_error)"""


class CudaError(RuntimeError):
    """isce3::cuda::except::CudaError"""


def parse_dry_tropo_model(s: str) -> int:
    """isce3::focus::parseDryTropoModel (cxx/isce3/focus/DryTroposphereModel.cpp:10-21)"""
    if s == "nodelay":
        return _capi.TROPO_NODELAY
    if s == "tsx":
        return _capi.TROPO_TSX
    raise InvalidArgument(f"unexpected dry troposphere model '{s}'")


def parse_rdr2geo_params(params: dict) -> _capi.Rdr2GeoBracketParams:
    out = _capi.Rdr2GeoBracketParams(1e-5, 0.0, math.pi / 2)  # geometry/detail/Rdr2Geo.h:85-97
    for key, val in dict(params or {}).items():
        if key in ("tol_height", "look_min", "look_max"):
            setattr(out, key, float(val))
        else:
            raise InvalidArgument(f"unexpected rdr2geo_bracket keyword: {key}")
    return out


def parse_geo2rdr_params(params: dict) -> _capi.Geo2RdrBracketParams:
    out = _capi.Geo2RdrBracketParams(1e-7, 0, 0, 0.0, 0.0)  # geometry/detail/Geo2Rdr.h:54-68
    for key, val in dict(params or {}).items():
        if key == "tol_aztime":
            out.tol_aztime = float(val)
        elif key == "time_start":
            if val is not None:
                out.has_time_start, out.time_start = 1, float(val)
        elif key == "time_end":
            if val is not None:
                out.has_time_end, out.time_end = 1, float(val)
        else:
            raise InvalidArgument(f"unexpected geo2rdr_bracket keyword: {key}")
    return out


class DeviceLines:
    """complex64 [lines][samples] living in HBM (output of ``RangeComp.rangecompress_to_device``);
    pass it as ``in_`` of ``backproject`` and the swath never crosses the host link."""

    dtype = np.dtype(np.complex64)
    ndim = 2

    def __init__(self, pointer: int, shape):
        self.pointer, self.shape = int(pointer), tuple(int(x) for x in shape)

    def to_host(self) -> np.ndarray:
        out = np.empty(self.shape, np.complex64)
        status = _capi.load_library().i3b_device_to_host(out.ctypes.data, self.pointer, out.nbytes)
        if status < 0:
            raise CudaError("device to host copy failed")
        return out

    def free(self):
        if self.pointer:
            _capi.load_library().i3b_device_free(self.pointer)
            self.pointer = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _check_array(a, name, dtype, shape, what):
    if isinstance(a, DeviceLines) and name == "input":
        if tuple(a.shape) != tuple(shape):
            raise InvalidArgument(f"{what} shape must match {name} radar grid shape")
        return
    if not isinstance(a, np.ndarray) or a.dtype != dtype:
        raise TypeError(f"{name} must be a numpy array of {np.dtype(dtype).name}")
    if a.ndim != 2:
        raise InvalidArgument(f"{what} must be 2-D")
    if tuple(a.shape) != tuple(shape):
        raise InvalidArgument(f"{what} shape must match {name} radar grid shape")
    if not a.flags.c_contiguous:
        raise TypeError(f"{name} must be C-contiguous")


def build_args(out, out_geometry, in_, in_geometry, dem, fc, ds, kernel,
               dry_tropo_model="tsx", rdr2geo_params=None, geo2rdr_params=None, batch=1024,
               height=None, devices=None, force_generic=False, range_cor=None,
               mantissa_nbits=None, pulse_times=None) -> Flattened:
    """Validate like the reference binding and flatten into an I3B_BackprojectArgs."""
    oshape = (out_geometry.grid_length, out_geometry.grid_width)
    ishape = (in_geometry.grid_length, in_geometry.grid_width)
    if out is not None:
        _check_array(out, "output", np.complex64, oshape, "output array")
        if not out.flags.writeable:
            raise TypeError("output array must be writeable")
    _check_array(in_, "input", np.complex64, ishape, "input signal data")
    if height is not None:
        if not isinstance(height, np.ndarray) or height.dtype != np.float32:
            raise TypeError("height must be a numpy array of float32")
        if tuple(height.shape) != oshape:
            raise InvalidArgument("height array shape must match output radar grid shape")
        if not height.flags.c_contiguous:
            raise TypeError("height must be C-contiguous")
    atm = parse_dry_tropo_model(dry_tropo_model)
    r2g = parse_rdr2geo_params(rdr2geo_params)
    g2r = parse_geo2rdr_params(geo2rdr_params)
    if int(batch) < 1:
        raise DomainError("batch size must be > 0")

    fl = Flattened()
    a = fl.args
    a.abi_version = _capi.ABI_VERSION
    a.flags = _capi.FLAG_FORCE_GENERIC if force_generic else 0
    a.out = out.ctypes.data if out is not None else None
    if isinstance(in_, DeviceLines):
        a.flags |= _capi.FLAG_DEVICE_INPUT
        a.in_ = in_.pointer
    else:
        a.in_ = in_.ctypes.data
    a.height = height.ctypes.data if height is not None else None
    fl.keep += [out, in_, height]
    a.out_geometry = _capi.flatten_geometry(out_geometry, fl)
    a.in_geometry = _capi.flatten_geometry(in_geometry, fl)
    a.dem = _capi.flatten_dem(dem, fl)
    a.fc, a.ds = float(fc), float(ds)
    a.kernel = _capi.flatten_kernel(kernel, fl)
    a.dry_tropo_model = atm
    a.batch = int(batch)
    a.rdr2geo, a.geo2rdr = r2g, g2r
    if range_cor is not None:
        rcor = fl.hold(range_cor, np.complex64)
        if rcor.shape != (oshape[1],):
            raise InvalidArgument("range_cor length must match the output radar grid width")
        a.range_cor = rcor.ctypes.data
    if mantissa_nbits is not None:
        if not 0 < int(mantissa_nbits) <= 23:  # isce3/core/types.py:147-151
            raise InvalidArgument(f"Require 0 < significant_bits={mantissa_nbits} <= 23")
        a.mantissa_nbits = int(mantissa_nbits) % 23  # 23 keeps every bit
    if pulse_times is not None:
        pt = fl.hold(pulse_times, np.float64)
        if pt.shape != (ishape[0],):
            raise InvalidArgument("pulse_times length must match the number of input lines")
        a.pulse_times = pt.ctypes.data
    if devices:
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        fl.keep.append(dev)
        a.n_devices = dev.size
        a.devices = dev.ctypes.data_as(C.POINTER(C.c_int32))
    return fl


def raise_for_status(status: int, message: str):
    """Translate an I3B_EXC_* status the way the adapter rethrows isce3 exceptions."""
    if status >= 0:
        return
    table = {
        _capi.EXC_INVALID_ARGUMENT: InvalidArgument,
        _capi.EXC_RUNTIME_ERROR: RuntimeError,
        _capi.EXC_DOMAIN_ERROR: DomainError,
        _capi.EXC_OVERFLOW_ERROR: OverflowError,
        _capi.EXC_OUT_OF_RANGE: IndexError,
        _capi.EXC_CUDA_ERROR: CudaError,
        _capi.EXC_NO_DEVICE: CudaError,
    }
    raise table.get(status, RuntimeError)(message or f"isce3_b200 error {status}")


def last_stats() -> dict:
    st = _capi.Stats()
    _capi.load_library().i3b_last_stats(C.byref(st))
    return st.as_dict()


def backproject(out, out_geometry, in_, in_geometry, dem, fc, ds, kernel,
                dry_tropo_model="tsx", rdr2geo_params=None, geo2rdr_params=None, batch=1024,
                height=None, *, devices=None, force_generic=False, range_cor=None,
                mantissa_nbits=None, pulse_times=None) -> bool:
    """Focus in azimuth via time-domain backprojection on B200.

    Same positional arguments, defaults and return value as
    ``isce3.cuda.focus.backproject`` (returns True when any pixel's geometry failed
    to converge; those pixels are NaN).  ``devices`` (keyword-only extension) lists
    CUDA device ordinals to shard the output grid over by azimuth block.  ``range_cor``
    (complex64 per output range column) and ``mantissa_nbits`` (keyword-only extensions) fuse
    what the workflow's writer does to each block on the host -- ``z *= range_cor`` and
    ``truncate_mantissa(z, n)`` (nisar/workflows/focus.py:899-925) -- into the device pass
    that produces ``out``.  ``pulse_times`` (keyword-only extension): azimuth time of every
    input line for non-uniformly spaced pulses (dithered PRF) -- focuses them where they were
    recorded instead of resampling the raw data to a uniform grid first
    (nisar/workflows/focus.py:973-1061).
    """
    fl = build_args(out, out_geometry, in_, in_geometry, dem, fc, ds, kernel, dry_tropo_model,
                    rdr2geo_params, geo2rdr_params, batch, height, devices, force_generic,
                    range_cor, mantissa_nbits, pulse_times)
    if out is None:
        raise TypeError("output array is required")
    lib = _capi.load_library()
    status = lib.i3b_backproject(C.byref(fl.args))
    if status < 0:
        raise_for_status(status, (lib.i3b_last_error() or b"").decode())
    return status != _capi.SUCCESS


class BackprojectPlan:
    """Resident variant (i3b_plan_*): the range-compressed swath and geometry are
    uploaded once; ``execute`` runs target solve + accumulation on the device and
    ``download`` copies the image back.  Used by bench.py for the HBM-resident
    throughput figure."""

    def __init__(self, out_geometry, in_, in_geometry, dem, fc, ds, kernel,
                 dry_tropo_model="tsx", rdr2geo_params=None, geo2rdr_params=None, batch=1024,
                 force_generic=False, devices=None, pulse_times=None):
        self._lib = _capi.load_library()
        self._shape = (out_geometry.grid_length, out_geometry.grid_width)
        fl = build_args(None, out_geometry, in_, in_geometry, dem, fc, ds, kernel,
                        dry_tropo_model, rdr2geo_params, geo2rdr_params, batch, None, devices,
                        force_generic, pulse_times=pulse_times)
        self._handle = C.c_void_p()
        status = self._lib.i3b_plan_create(C.byref(fl.args), C.byref(self._handle))
        if status < 0:
            raise_for_status(status, (self._lib.i3b_last_error() or b"").decode())

    def execute(self) -> bool:
        status = self._lib.i3b_plan_execute(self._handle)
        if status < 0:
            raise_for_status(status, (self._lib.i3b_last_error() or b"").decode())
        return status != _capi.SUCCESS

    def download(self, out=None, height=None):
        if out is None:
            out = np.empty(self._shape, np.complex64)
        status = self._lib.i3b_plan_download(
            self._handle, out.ctypes.data, height.ctypes.data if height is not None else None)
        if status < 0:
            raise_for_status(status, (self._lib.i3b_last_error() or b"").decode())
        return out

    def stats(self) -> dict:
        return last_stats()

    def close(self):
        if self._handle:
            self._lib.i3b_plan_destroy(self._handle)
            self._handle = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BlockFocuser:
    """One swath, many output blocks (i3b_blocks_*): what the workflow does with one
    ``backproject`` call per block and the whole swath pointer each time
    (nisar/workflows/focus.py:726-783, :1988-2007), with the swath, DEM, LUTs and per-pulse
    tables uploaded ONCE per device.  ``out_geometry`` is the geometry of the whole image
    (orbit and Doppler LUT are shared by every block); ``run`` takes the blocks as
    ``(radar_grid_of_the_block, out_array[, height_array])`` and hands them to the listed devices
    dynamically.  Returns True if any pixel failed (like ``backproject``)."""

    def __init__(self, out_geometry, in_, in_geometry, dem, fc, ds, kernel, dry_tropo_model="tsx",
                 rdr2geo_params=None, geo2rdr_params=None, devices=None, force_generic=False,
                 range_cor=None, mantissa_nbits=None, pulse_times=None):
        self._lib = _capi.load_library()
        self._fl = build_args(None, out_geometry, in_, in_geometry, dem, fc, ds, kernel, dry_tropo_model,
                              rdr2geo_params, geo2rdr_params, 1024, None, devices, force_generic,
                              range_cor, mantissa_nbits, pulse_times)
        self._handle = C.c_void_p()
        status = self._lib.i3b_blocks_create(C.byref(self._fl.args), C.byref(self._handle))
        if status < 0:
            raise_for_status(status, (self._lib.i3b_last_error() or b"").decode())

    def run(self, blocks) -> bool:
        blocks = list(blocks)
        n = len(blocks)
        grids = (_capi.RadarGrid * max(n, 1))()
        outs = (C.c_void_p * max(n, 1))()
        heights = (C.c_void_p * max(n, 1))()
        any_height = False
        for i, blk in enumerate(blocks):
            grid, out = blk[0], blk[1]
            height = blk[2] if len(blk) > 2 else None
            grid = getattr(grid, "radar_grid", grid)
            shape = (grid.length, grid.width)
            _check_array(out, "output", np.complex64, shape, "output array")
            grids[i] = _capi.flatten_grid(grid)
            outs[i] = out.ctypes.data
            if height is not None:
                _check_array(height, "height", np.float32, shape, "height array")
                heights[i] = height.ctypes.data
                any_height = True
        status = self._lib.i3b_blocks_run(self._handle, n, grids, outs, heights if any_height else None)
        if status < 0:
            raise_for_status(status, (self._lib.i3b_last_error() or b"").decode())
        return status != _capi.SUCCESS

    def stats(self) -> dict:
        return last_stats()

    def close(self):
        if self._handle:
            self._lib.i3b_blocks_destroy(self._handle)
            self._handle = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def keep_device_memory(limit_mb=-1) -> None:
    """Opt in to the per-process cache of device buffers (default: every call returns its device
    memory to the driver, like the reference).  ``limit_mb`` < 0: unlimited, 0: off."""
    _capi.load_library().i3b_set_device_memory_pool(int(limit_mb))


def release_device_memory() -> None:
    """Hand the device memory cached by earlier calls back to the driver."""
    _capi.load_library().i3b_release_device_memory()


def measure_peaks(device=0) -> dict:
    pk = _capi.Peaks()
    lib = _capi.load_library()
    status = lib.i3b_measure_peaks(int(device), C.byref(pk))
    if status < 0:
        raise_for_status(status, (lib.i3b_last_error() or b"").decode())
    return pk.as_dict()


# ---- range compression (isce3.focus.RangeComp, form_linear_chirp) ------------------------------

import enum  # noqa: E402


def form_linear_chirp(chirprate, duration, samplerate, centerfreq=0.0, amplitude=1.0, phi=0.0):
    """isce3.focus.form_linear_chirp (cxx/isce3/focus/Chirp.cpp:10-58): complex64 LFM replica,
    odd number of samples centred on t = 0."""
    if duration <= 0.0:
        raise DomainError("chirp duration must be > 0")
    if samplerate <= 0.0:
        raise DomainError("sampling rate must be > 0")
    if amplitude <= 0.0:
        raise DomainError("amplitude must be > 0")
    size = int(math.floor(samplerate * duration + 1))
    if size % 2 == 0:
        size += 1
    spacing = 1.0 / samplerate
    tau = -0.5 * (size - 1) * spacing + spacing * np.arange(size)
    phase = phi + 2.0 * np.pi * (centerfreq + 0.5 * chirprate * tau) * tau
    return (amplitude * np.exp(1j * phase)).astype(np.complex64)


class RangeComp:
    """``isce3.focus.RangeComp`` on the GPU (python/extensions/pybind_isce3/focus/RangeComp.cpp:
    same constructor, ``rangecompress(out, in)``, read-only properties and length checks).  The
    reference has no CUDA twin of this class; see csrc/rangecomp.cu."""

    class Mode(enum.IntEnum):
        Full = 0
        Valid = 1
        Same = 2

    def __init__(self, chirp, inputsize, maxbatch=1, mode=Mode.Full):
        self._lib = _capi.load_library()
        self._chirp = np.ascontiguousarray(chirp, dtype=np.complex64).ravel()
        self._handle = C.c_void_p()
        self._mode = RangeComp.Mode(int(mode))
        status = self._lib.i3b_rangecomp_create(self._chirp.ctypes.data, self._chirp.size, int(inputsize),
                                                int(maxbatch), int(self._mode), C.byref(self._handle))
        self._check(status)
        self._input_size, self._maxbatch = int(inputsize), int(maxbatch)
        nfft, nout, first = C.c_int(), C.c_int(), C.c_int()
        self._lib.i3b_rangecomp_query(self._handle, C.byref(nfft), C.byref(nout), C.byref(first))
        self._fft_size, self._output_size, self._first = nfft.value, nout.value, first.value

    def _check(self, status):
        if status < 0:
            msg = (self._lib.i3b_rangecomp_last_error() or b"").decode()
            if status == _capi.EXC_LENGTH_ERROR:
                raise ValueError(msg)  # std::length_error -> ValueError in pybind11
            raise_for_status(status, msg)

    chirp_size = property(lambda self: self._chirp.size)
    input_size = property(lambda self: self._input_size)
    fft_size = property(lambda self: self._fft_size)
    maxbatch = property(lambda self: self._maxbatch)
    mode = property(lambda self: self._mode)
    output_size = property(lambda self: self._output_size)
    first_valid_sample = property(lambda self: self._first)

    def set_scaling(self, column_scale=None, slant_ranges=None, pattern_ranges=None):
        """Radiometric corrections fused into the pass that writes the output (extension): what
        the workflow applies to every range-compressed block on the host
        (nisar/workflows/focus.py:1956-1975).  ``column_scale`` (complex per output sample): the
        per-column factors multiplied together -- baseband-shift phasors ``deramp_rc`` and range
        loss ``slant_ranges / ref_range``.  ``slant_ranges`` + ``pattern_ranges``: enable the
        per-line antenna-pattern division; ``rangecompress(..., patterns=P)`` then divides line b
        by ``numpy.interp(slant_ranges, pattern_ranges, P[b])``.  No arguments: clear."""
        self._scaling_keep = []
        self._n_pattern = 0
        if column_scale is None and pattern_ranges is None:
            self._check(self._lib.i3b_rangecomp_set_scaling(self._handle, None))
            return
        sc = _capi.RangeCompScaling()
        if column_scale is not None:
            col = np.ascontiguousarray(column_scale, dtype=np.complex64)
            if col.shape != (self._output_size,):
                raise ValueError("column_scale length must equal output_size")
            self._scaling_keep.append(col)
            sc.column_scale = col.ctypes.data
        if pattern_ranges is not None:
            pr = np.ascontiguousarray(pattern_ranges, dtype=np.float64)
            sr = np.ascontiguousarray(slant_ranges, dtype=np.float64)
            if sr.shape != (self._output_size,):
                raise ValueError("slant_ranges length must equal output_size")
            self._scaling_keep += [pr, sr]
            sc.slant_ranges, sc.pattern_ranges, sc.n_pattern = sr.ctypes.data, pr.ctypes.data, pr.size
            self._n_pattern = pr.size
        self._check(self._lib.i3b_rangecomp_set_scaling(self._handle, C.byref(sc)))

    def _patterns(self, patterns, batch):
        if patterns is None:
            return None
        if not getattr(self, "_n_pattern", 0):
            raise ValueError("set_scaling(pattern_ranges=...) first")
        pat = np.ascontiguousarray(patterns, dtype=np.complex64).reshape(batch, -1)
        if pat.shape[1] != self._n_pattern:
            raise ValueError("patterns must have one row of len(pattern_ranges) samples per line")
        return pat

    def rangecompress(self, out, in_, patterns=None):
        for a, name in ((out, "out"), (in_, "in")):
            if not isinstance(a, np.ndarray) or a.dtype != np.complex64 or not a.flags.c_contiguous:
                raise TypeError(f"{name} must be a C-contiguous numpy array of complex64")
        if in_.ndim != out.ndim:
            raise ValueError("require same ndim on input and output")
        if in_.ndim == 2:
            batch = in_.shape[0]
            if in_.shape[0] != out.shape[0]:
                raise ValueError("require equal batch size on input and output")
            nin, nout = in_.shape[1], out.shape[1]
        elif in_.ndim == 1:
            batch, nin, nout = 1, in_.shape[0], out.shape[0]
        else:
            raise ValueError("require 1D or 2D data")
        if nin != self._input_size:
            raise ValueError("unexpected input length")
        if nout != self._output_size:
            raise ValueError("unexpected output length")
        pat = self._patterns(patterns, batch)
        self._check(self._lib.i3b_rangecomp_execute_scaled(self._handle, out.ctypes.data, in_.ctypes.data, batch, 0,
                                                           pat.ctypes.data if pat is not None else None))

    def rangecompress_to_device(self, in_, patterns=None) -> "DeviceLines":
        """Range-compress all lines of ``in_`` (2-D, any number of lines: chunks of ``maxbatch``)
        and leave the result in HBM (extension; see I3B_FLAG_DEVICE_INPUT)."""
        if not isinstance(in_, np.ndarray) or in_.dtype != np.complex64 or not in_.flags.c_contiguous:
            raise TypeError("in must be a C-contiguous numpy array of complex64")
        if in_.ndim != 2:
            raise ValueError("require 2D data")
        if in_.shape[1] != self._input_size:
            raise ValueError("unexpected input length")
        ptr = C.c_void_p()
        pat = self._patterns(patterns, in_.shape[0])
        self._check(self._lib.i3b_rangecomp_execute_to_device_scaled(
            self._handle, in_.ctypes.data, in_.shape[0], pat.ctypes.data if pat is not None else None,
            C.byref(ptr)))
        return DeviceLines(ptr.value, (in_.shape[0], self._output_size))

    def last_device_ms(self) -> float:
        return float(self._lib.i3b_rangecomp_last_device_ms(self._handle))

    def close(self):
        if getattr(self, "_handle", None):
            self._lib.i3b_rangecomp_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
