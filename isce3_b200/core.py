"""Host-side value types that cross the backproject boundary, mirroring the names
the reference registers in ``isce3.ext.isce3.core``
(python/extensions/pybind_isce3/core/core.cpp:34-84).

Only what the TDBP path needs is provided: these are plain containers whose
contents are flattened into the C descriptors of
``include/isce3_b200_backproject.h`` -- the numerics of the path run on the GPU.
The few evaluators kept on the host (``Orbit.interpolate``, ``LUT2d.eval``,
``Kernel.__call__``) exist because the reference exposes them on these types
and the scene generator / tests use them; they are numpy restatements citing
the reference lines they follow.
"""
from __future__ import annotations

import datetime as _dt
import enum
import math

import numpy as np

from . import _capi

speed_of_light = 299792458.0  # cxx/isce3/core/Constants.h:50
earth_semi_major_axis = 6378137.0  # Constants.h:41
earth_eccentricity_squared = 0.006694379990141317  # Constants.h:44


class LookSide(enum.IntEnum):
    """cxx/isce3/core/LookSide.h:13-17"""
    Left = 1
    Right = -1


def parse_look_side(s) -> LookSide:
    if isinstance(s, LookSide):
        return s
    key = str(s).lower()
    if key == "left":
        return LookSide.Left
    if key == "right":
        return LookSide.Right
    raise ValueError(f"invalid look side {s!r}")


class OrbitInterpMethod(enum.IntEnum):
    """cxx/isce3/core/Orbit.h:21-24"""
    HERMITE = 0
    LEGENDRE = 1


class OrbitInterpBorderMode(enum.IntEnum):
    """cxx/isce3/core/Orbit.h:27-31"""
    ERROR = 0
    EXTRAPOLATE = 1
    FILL_NAN = 2


class DataInterpMethod(enum.IntEnum):
    """cxx/isce3/core/Constants.h:23-29"""
    SINC = 0
    BILINEAR = 1
    BICUBIC = 2
    NEAREST = 3
    BIQUINTIC = 4


def parse_interp_method(m) -> DataInterpMethod:
    if isinstance(m, DataInterpMethod):
        return m
    if isinstance(m, (int, np.integer)):
        return DataInterpMethod(int(m))
    table = {"sinc": 0, "bilinear": 1, "bicubic": 2, "nearest": 3, "biquintic": 4}
    key = str(m).lower()
    if key not in table:
        raise ValueError(f"unknown interpolation method {m!r}")
    return DataInterpMethod(table[key])


class DateTime:
    """Reference epoch.  Only equality matters on the TDBP path
    (Backproject.cpp:88-92); stored as whole seconds since 1970 + fraction."""

    def __init__(self, *args):
        if len(args) == 1 and isinstance(args[0], DateTime):
            self._sec, self._frac = args[0]._sec, args[0]._frac
        elif len(args) == 1 and isinstance(args[0], str):
            s = args[0].strip().replace("T", " ")
            frac = 0.0
            if "." in s:
                s, f = s.split(".")
                frac = float("0." + f)
            d = _dt.datetime.strptime(s, "%Y-%m-%d %H:%M:%S").replace(tzinfo=_dt.timezone.utc)
            self._sec, self._frac = int(d.timestamp()), frac
        elif len(args) == 1 and isinstance(args[0], _dt.datetime):
            d = args[0].replace(tzinfo=_dt.timezone.utc)
            self._sec, self._frac = int(d.replace(microsecond=0).timestamp()), d.microsecond * 1e-6
        elif len(args) >= 3:
            y, mo, d = args[:3]
            hh, mn = (list(args[3:5]) + [0, 0])[:2]
            ss = args[5] if len(args) > 5 else 0
            ff = args[6] if len(args) > 6 else 0.0
            isec = int(math.floor(ss))
            d0 = _dt.datetime(y, mo, d, hh, mn, isec, tzinfo=_dt.timezone.utc)
            self._sec, self._frac = int(d0.timestamp()), float(ss - isec) + float(ff)
        elif len(args) == 0:
            self._sec, self._frac = 0, 0.0
        else:
            raise TypeError("unsupported DateTime arguments")

    def epoch_pair(self):
        return int(self._sec), float(self._frac)

    def __add__(self, seconds: float) -> "DateTime":
        out = DateTime(self)
        tot = out._frac + float(seconds)
        whole = math.floor(tot)
        out._sec += int(whole)
        out._frac = tot - whole
        return out

    def __sub__(self, other: "DateTime") -> float:
        return float(self._sec - other._sec) + (self._frac - other._frac)

    def __eq__(self, other):
        return isinstance(other, DateTime) and self.epoch_pair() == other.epoch_pair()

    def __hash__(self):
        return hash(self.epoch_pair())

    def isoformat(self):
        d = _dt.datetime.fromtimestamp(self._sec, tz=_dt.timezone.utc)
        return d.strftime("%Y-%m-%dT%H:%M:%S") + ("%.9f" % self._frac)[1:]

    def __repr__(self):
        return f"DateTime({self.isoformat()})"


class Linspace:
    """cxx/isce3/core/Linspace.h:9-127"""

    def __init__(self, first, spacing, size):
        self.first, self.spacing, self.size = float(first), float(spacing), int(size)

    @property
    def last(self):
        return self.first + (self.size - 1) * self.spacing

    def __len__(self):
        return self.size

    def __getitem__(self, i):
        return self.first + np.asarray(i) * self.spacing

    def __array__(self, dtype=None, copy=None):
        a = self.first + np.arange(self.size) * self.spacing
        return a.astype(dtype) if dtype is not None else a

    def search(self, val):
        """Linspace.icc:73-89"""
        if self.spacing >= 0:
            if val < self.first:
                return 0
            if val > self.last:
                return self.size
        else:
            if val > self.first:
                return 0
            if val < self.last:
                return self.size
        return int((val - self.first) / self.spacing + 1)


class StateVector:
    """pybind_isce3/core/StateVector.cpp:15-19"""

    def __init__(self, datetime: DateTime, position, velocity):
        self.datetime = datetime
        self.position = np.asarray(position, dtype=np.float64).reshape(3)
        self.velocity = np.asarray(velocity, dtype=np.float64).reshape(3)


class Orbit:
    """isce3::core::Orbit (cxx/isce3/core/Orbit.h:36-199): uniformly spaced state
    vectors + interpolation method.  ``Orbit(state_vectors[, reference_epoch],
    interp_method)`` as pybind_isce3/core/Orbit.cpp:33-51."""

    def __init__(self, state_vectors, *args, interp_method=OrbitInterpMethod.HERMITE, type=""):
        ref = None
        for a in args:
            if isinstance(a, DateTime):
                ref = a
            elif isinstance(a, (OrbitInterpMethod, int, np.integer)):
                interp_method = OrbitInterpMethod(int(a))
            elif isinstance(a, str):
                type = a
        svs = list(state_vectors)
        if len(svs) < 2:
            raise ValueError("at least two state vectors are required")
        if ref is None:
            ref = svs[0].datetime
        self.reference_epoch = ref
        t = np.array([sv.datetime - ref for sv in svs])
        spacing = t[1] - t[0]
        if not np.allclose(np.diff(t), spacing, rtol=0, atol=1e-9):
            raise ValueError("non-uniform spacing between state vectors")
        self.time = Linspace(t[0], spacing, len(svs))
        self.position = np.array([sv.position for sv in svs], dtype=np.float64)
        self.velocity = np.array([sv.velocity for sv in svs], dtype=np.float64)
        self.interp_method = OrbitInterpMethod(interp_method)
        self.type = type

    @classmethod
    def from_arrays(cls, t0, dt, position, velocity, reference_epoch=None,
                    interp_method=OrbitInterpMethod.HERMITE):
        """Build directly from a uniform time axis (exact t0/dt, no DateTime round trip)."""
        self = cls.__new__(cls)
        self.reference_epoch = reference_epoch or DateTime(2000, 1, 1)
        self.position = np.ascontiguousarray(position, dtype=np.float64).reshape(-1, 3)
        self.velocity = np.ascontiguousarray(velocity, dtype=np.float64).reshape(-1, 3)
        self.time = Linspace(t0, dt, len(self.position))
        self.interp_method = OrbitInterpMethod(interp_method)
        self.type = ""
        return self

    @property
    def size(self):
        return self.time.size

    @property
    def spacing(self):
        return self.time.spacing

    @property
    def start_time(self):
        return self.time.first

    @property
    def end_time(self):
        return self.time.last

    @property
    def mid_time(self):
        return self.start_time + 0.5 * (self.size - 1) * self.spacing

    def contains(self, t):
        return self.start_time <= t <= self.end_time

    def interpolate(self, t, border_mode=OrbitInterpBorderMode.ERROR):
        """(position, velocity) at time t: core/Orbit.cpp:71-86 ->
        core/detail/InterpolateOrbit.icc:15-109 (Hermite), :116-155 (Legendre)."""
        t = float(t)
        if t < self.start_time or t > self.end_time:
            if border_mode == OrbitInterpBorderMode.ERROR:
                raise IndexError("orbit interpolation outside of orbit domain")
            if border_mode == OrbitInterpBorderMode.FILL_NAN:
                return np.full(3, np.nan), np.full(3, np.nan)
        n = self.size
        if self.interp_method == OrbitInterpMethod.HERMITE:
            if n < 4:
                raise ValueError("need >= 4 state vectors for Hermite interpolation")
            idx = min(max(self.time.search(t) - 2, 0), n - 4)
            tt = self.time[np.arange(idx, idx + 4)]
            P, V = self.position[idx:idx + 4], self.velocity[idx:idx + 4]
            f1 = t - tt
            f0, h, hdot, gsum = np.empty(4), np.ones(4), np.zeros(4), np.empty(4)
            for i in range(4):
                s = sum(1.0 / (tt[i] - tt[j]) for j in range(4) if j != i)
                gsum[i] = s
                f0[i] = 1.0 - 2.0 * s * (t - tt[i])
                for j in range(4):
                    if j != i:
                        h[i] *= (t - tt[j]) / (tt[i] - tt[j])
                for j in range(4):
                    if j == i:
                        continue
                    prod = 1.0 / (tt[i] - tt[j])
                    for k in range(4):
                        if k != i and k != j:
                            prod *= (t - tt[k]) / (tt[i] - tt[k])
                    hdot[i] += prod
            g1 = h + 2.0 * hdot * (t - tt)
            g0 = 2.0 * (f0 * hdot - gsum * h)
            pos = ((h * h)[:, None] * (P * f0[:, None] + V * f1[:, None])).sum(0)
            vel = (h[:, None] * (P * g0[:, None] + V * g1[:, None])).sum(0)
            return pos, vel
        if n < 9:
            raise ValueError("need >= 9 state vectors for Legendre interpolation")
        idx = min(max(self.time.search(t) - 5, 0), n - 9)
        trel = 8.0 * (t - self.time[idx]) / (self.time[idx + 8] - self.time[idx])
        teller = float(np.prod(trel - np.arange(9)))
        if teller == 0.0:
            i = int(trel)
            return self.position[idx + i].copy(), self.velocity[idx + i].copy()
        noemer = np.array([40320.0, -5040.0, 1440.0, -720.0, 576.0, -720.0, 1440.0, -5040.0, 40320.0])
        coeff = (teller / noemer) / (trel - np.arange(9))
        return coeff @ self.position[idx:idx + 9], coeff @ self.velocity[idx:idx + 9]


class LUT2d:
    """isce3::core::LUT2d<double> (cxx/isce3/core/LUT2d.h:54-95).  ``LUT2d()`` has no
    data and evaluates to ``ref_value`` (0); ``LUT2d(xstart, ystart, dx, dy, data,
    method="bilinear", b_error=True)`` as pybind_isce3/core/LUT2d.cpp:39-95."""

    def __init__(self, *args, method="bilinear", b_error=True):
        self.have_data = False
        self.ref_value = 0.0
        self.bounds_error = bool(b_error)
        self.interp_method = parse_interp_method(method)
        self.data = None
        self.x_start = self.y_start = 0.0
        self.x_spacing = self.y_spacing = 1.0
        if len(args) == 0:
            return
        if len(args) >= 6:
            self.interp_method = parse_interp_method(args[5])
        if len(args) >= 7:
            self.bounds_error = bool(args[6])
        if len(args) >= 5:
            xstart, ystart, dx, dy, data = args[:5]
        elif len(args) >= 3:
            xc, yc, data = np.asarray(args[0], float), np.asarray(args[1], float), args[2]
            if len(args) >= 4:
                self.interp_method = parse_interp_method(args[3])
            if len(args) >= 5:
                self.bounds_error = bool(args[4])
            xstart, ystart = xc[0], yc[0]
            dx, dy = xc[1] - xc[0], yc[1] - yc[0]
        else:
            raise TypeError("unsupported LUT2d arguments")
        data = np.ascontiguousarray(data, dtype=np.float64)
        if data.ndim != 2:
            raise ValueError("LUT2d data must be 2-D")
        self.data = data
        self.have_data = True
        self.ref_value = float(data[0, 0])  # LUT2d.cpp:121
        self.x_start, self.y_start = float(xstart), float(ystart)
        self.x_spacing, self.y_spacing = float(dx), float(dy)

    @property
    def length(self):
        return 0 if self.data is None else self.data.shape[0]

    @property
    def width(self):
        return 0 if self.data is None else self.data.shape[1]

    def contains(self, y, x):
        """LUT2d.h:84-95"""
        if not self.have_data:
            return True
        i = (x - self.x_start) / self.x_spacing
        j = (y - self.y_start) / self.y_spacing
        return 0.0 <= i <= self.width - 1.0 and 0.0 <= j <= self.length - 1.0

    def eval(self, y, x):
        """fD(azimuth time y, slant range x): LUT2d.cpp:127-160, bilinear only on host
        (BilinearInterpolator.cpp:13-48); other methods are evaluated on device."""
        if not self.have_data:
            return self.ref_value
        xi = min(max((x - self.x_start) / self.x_spacing, 0.0), self.width - 1.0)
        yi = min(max((y - self.y_start) / self.y_spacing, 0.0), self.length - 1.0)
        if self.interp_method != DataInterpMethod.BILINEAR:
            raise NotImplementedError("host LUT2d.eval supports bilinear only")
        x1, x2 = int(math.floor(xi)), int(math.ceil(xi))
        y1, y2 = int(math.floor(yi)), int(math.ceil(yi))
        z = self.data
        if x1 == x2 and y1 == y2:
            return float(z[y1, x1])
        if y1 == y2:
            return float((x2 - xi) * z[y1, x1] + (xi - x1) * z[y1, x2])
        if x1 == x2:
            return float((y2 - yi) * z[y1, x1] + (yi - y1) * z[y2, x1])
        return float(z[y1, x1] * (x2 - xi) * (y2 - yi) + z[y1, x2] * (xi - x1) * (y2 - yi) +
                     z[y2, x1] * (x2 - xi) * (yi - y1) + z[y2, x2] * (xi - x1) * (yi - y1))


# ---- interpolation kernels -------------------------------------------------

def _sinc(t):
    """cxx/isce3/math/Sinc.icc:69-91 (numpy.sinc is sin(pi t)/(pi t))"""
    return np.sinc(t)


class Kernel:
    """isce3::core::Kernel<double> base (cxx/isce3/core/Kernels.h:18-37)."""

    def __init__(self, width):
        self._halfwidth = abs(width / 2.0)

    @property
    def width(self):
        return self._halfwidth * 2.0

    def __call__(self, t):
        raise NotImplementedError

    def _flatten(self):
        raise TypeError(
            f"{type(self).__name__} is a double-precision kernel; backproject takes a "
            "Kernel<float>: wrap it in TabulatedKernelF32 or ChebyKernelF32 "
            "(pybind_isce3/core/Kernels.cpp:13-141)")


class BartlettKernel(Kernel):
    """Kernels.icc:15-23"""

    def __call__(self, t):
        t2 = np.abs(np.asarray(t, dtype=np.float64) / self._halfwidth)
        return np.where(t2 > 1.0, 0.0, 1.0 - t2)


class LinearKernel(BartlettKernel):
    def __init__(self):
        super().__init__(2.0)


class KnabKernel(Kernel):
    """Knab (1983) windowed sinc: Kernels.icc:29-52."""

    def __init__(self, width, bandwidth):
        super().__init__(width)
        if not (0.0 < bandwidth < 1.0):
            raise ValueError("Require 0 < bandwidth < 1")
        self.bandwidth = float(bandwidth)

    def __call__(self, t):
        t = np.asarray(t, dtype=np.float64)
        c = math.pi * self._halfwidth * (1.0 - self.bandwidth)
        tf = t / self._halfwidth
        y = np.sqrt((1.0 - tf * tf).astype(np.complex128))
        window = np.real(np.cosh(c * y) / np.cosh(c))
        return window * _sinc(t)


class AzimuthKernel(Kernel):
    """Kernels.icc:93-108"""

    def __init__(self, scale):
        super().__init__(2.0 * scale)

    def __call__(self, t):
        x = np.abs(np.asarray(t, dtype=np.float64) * 2 / self._halfwidth)
        a = x * x * (0.75 * x - 1.5) + 1.0
        b = x * (x * (-0.25 * x + 1.5) - 3.0) + 2.0
        return np.where(x > 2.0, 0.0, np.where(x < 1.0, a, b))


class KernelF32(Kernel):
    """isce3::core::Kernel<float>: what backproject accepts."""


class BartlettKernelF32(KernelF32):
    """Reachable from C++ only in the reference."""

    def __call__(self, t):
        t2 = np.abs(np.asarray(t, dtype=np.float64) / self._halfwidth)
        return np.where(t2 > 1.0, 0.0, 1.0 - t2).astype(np.float32)

    def _flatten(self):
        return _capi.KERNEL_BARTLETT, self.width, 0.0, None


class LinearKernelF32(BartlettKernelF32):
    def __init__(self):
        super().__init__(2.0)

    def _flatten(self):
        return _capi.KERNEL_LINEAR, 2.0, 0.0, None


class KnabKernelF32(KernelF32):
    """KnabKernel<float>: evaluated in float on device (Kernels.icc:29-52)."""

    def __init__(self, width, bandwidth):
        super().__init__(width)
        if not (0.0 < bandwidth < 1.0):
            raise ValueError("Require 0 < bandwidth < 1")
        self.bandwidth = float(bandwidth)

    def __call__(self, t):
        return KnabKernel(self.width, self.bandwidth)(t).astype(np.float32)

    def _flatten(self):
        return _capi.KERNEL_KNAB, self.width, self.bandwidth, None


class TabulatedKernelF32(KernelF32):
    """TabulatedKernel<float>(kernel, n): Kernels.icc:114-154.  The table holds the
    even kernel on [0, halfwidth]; evaluation is linear interpolation with
    ``_imax = n-2`` and ``_1_dx`` stored in float."""

    def __init__(self, kernel, n):
        super().__init__(kernel.width)
        n = int(n)
        if n < 2:
            raise ValueError("Require table size >= 2.")
        dx = self._halfwidth / (n - 1.0)
        self._one_dx = np.float32(1.0 / dx)
        self._imax = n - 2
        self.table = np.array([kernel(i * dx) for i in range(n)], dtype=np.float64).astype(np.float32)

    def __call__(self, t):
        ax = np.abs(np.asarray(t, dtype=np.float64))
        axn = ax * np.float64(self._one_dx)
        i = np.minimum(np.floor(axn).astype(np.int64), self._imax)
        i = np.clip(i, 0, self._imax)
        tb = self.table
        val = tb[i].astype(np.float64) + (axn - i) * (tb[i + 1] - tb[i]).astype(np.float64)
        return np.where(ax > self._halfwidth, 0.0, val).astype(np.float32)

    def _flatten(self):
        return _capi.KERNEL_TABULATED, self.width, 0.0, self.table


class ChebyKernelF32(KernelF32):
    """ChebyKernel<float>(kernel, n): Kernels.icc:156-211 (fit in float)."""

    def __init__(self, kernel, n):
        super().__init__(kernel.width)
        n = int(n)
        if n < 1:
            raise ValueError("Need at least one coefficient.")
        f32 = np.float32
        self._scale = f32(4.0 / kernel.width)
        q = (math.pi * (2.0 * np.arange(n) + 1.0) / (2.0 * n)).astype(f32)
        x = ((np.cos(q).astype(np.float64) + 1.0) / np.float64(self._scale)).astype(f32)
        fx = np.array([kernel(float(xi)) for xi in x], dtype=np.float64).astype(f32)
        coeffs = np.zeros(n, dtype=f32)
        for i in range(n):
            acc = f32(0.0)
            for j in range(n):
                w = np.cos(f32(i) * q[j]).astype(f32)
                acc = f32(acc + w * fx[j])
            coeffs[i] = f32(np.float64(acc) * (2.0 / n))
        coeffs[0] = f32(np.float64(coeffs[0]) * 0.5)
        self.coeffs = coeffs

    def __call__(self, t):
        ax = np.abs(np.asarray(t, dtype=np.float64))
        q = ((ax * np.float64(self._scale)) - 1.0).astype(np.float32)
        twoq = np.float32(2) * q
        bk1 = np.zeros_like(q)
        bk2 = np.zeros_like(q)
        for i in range(len(self.coeffs) - 1, 0, -1):
            bk = self.coeffs[i] + twoq * bk1 - bk2
            bk2, bk1 = bk1, bk
        val = self.coeffs[0] + q * bk1 - bk2
        return np.where(ax > self._halfwidth, np.float32(0), val).astype(np.float32)

    def _flatten(self):
        return _capi.KERNEL_CHEBY, self.width, 0.0, self.coeffs
