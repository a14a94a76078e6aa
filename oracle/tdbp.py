"""TEST INFRASTRUCTURE -- Python handle on the two CPU oracles.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  It loads

* ``oracle/libtdbp_oracle.so``      the restatement ("port", oracle/tdbp_oracle.cpp)
* ``oracle/_ref/libtdbp_ref.so``    the reference's own sources compiled here
                                    (oracle/ref_capi.cpp + /root/reference/cxx)

and exposes the same calls on both, taking the host value types of
``isce3_b200`` (they share the flat C descriptors with the product library).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from isce3_b200 import _capi
from isce3_b200.focus import build_args

HERE = Path(__file__).resolve().parent
PORT_LIB = HERE / "libtdbp_oracle.so"
REF_LIB = HERE / "_ref" / "libtdbp_ref.so"
REFCUDA_LIB = HERE / "_ref" / "libtdbp_refcuda.so"
ADAPTER_LIB = HERE / "_ref" / "libtdbp_adapter.so"

BRENT_FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)


def build(verbose=False):
    """Compile the port always, and oracle/_ref when /root/reference is mounted."""
    r = subprocess.run(["make", "-C", str(HERE), "port", "ref"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


class Oracle:
    def __init__(self, path: Path, prefix: str):
        self.path, self.prefix = path, prefix
        self.lib = C.CDLL(str(path))
        self.kind = self._f("kind", C.c_char_p)().decode()
        f = self._f
        P = C.POINTER
        dp = P(C.c_double)
        self._backproject = f("backproject", C.c_int, P(_capi.BackprojectArgs))
        self._bistatic = f("bistatic_delay", C.c_double, dp, dp, dp)
        self._orbit = f("orbit_interpolate", C.c_int, P(_capi.Orbit), C.c_double, C.c_int, dp, dp)
        self._kernel = f("kernel_eval", C.c_int, P(_capi.Kernel), dp, C.c_int, P(C.c_float))
        self._tab = f("tabulate_knab", C.c_int, C.c_double, C.c_double, C.c_int, P(C.c_float))
        self._cheb = f("cheby_knab", C.c_int, C.c_double, C.c_double, C.c_int, P(C.c_float))
        self._interp1d = f("interp1d", C.c_int, P(_capi.Kernel), P(C.c_float), C.c_size_t, dp,
                           C.c_int, P(C.c_float))
        self._r2g = f("rdr2geo_bracket", C.c_int, C.c_double, C.c_double, C.c_double,
                      P(_capi.Orbit), P(_capi.DEM), C.c_double, C.c_int,
                      P(_capi.Rdr2GeoBracketParams), dp)
        self._g2r = f("geo2rdr_bracket", C.c_int, dp, P(_capi.Orbit), P(_capi.LUT2d), C.c_double,
                      C.c_int, P(_capi.Geo2RdrBracketParams), dp, dp)
        self._x2l = f("xyz_to_llh", None, dp, dp)
        self._l2x = f("llh_to_xyz", None, dp, dp)
        self._tropo = f("dry_tropo_tsx", C.c_double, dp, dp)
        self._brent = f("brent", C.c_int, C.c_double, C.c_double, BRENT_FN, C.c_void_p,
                        C.c_double, dp)
        self._lut = f("lut2d_eval", C.c_double, P(_capi.LUT2d), C.c_double, C.c_double)
        self._dem = f("dem_interp", C.c_double, P(_capi.DEM), C.c_double, C.c_double)
        self._proj = f("project_forward", C.c_int, C.c_int, C.c_double, C.c_double, dp)
        self._last_error = f("last_error", C.c_char_p)
        self._backproject_pp = None
        if prefix == "tdbp_oracle":
            self._backproject_pp = f("backproject_pp", C.c_int, P(_capi.BackprojectArgs), dp)

    def _f(self, name, restype, *argtypes):
        fn = getattr(self.lib, f"{self.prefix}_{name}")
        fn.restype = restype
        fn.argtypes = list(argtypes)
        return fn

    # -- whole path -----------------------------------------------------------
    def backproject(self, out, out_geometry, in_, in_geometry, dem, fc, ds, kernel,
                    dry_tropo_model="tsx", rdr2geo_params=None, geo2rdr_params=None,
                    height=None, return_status=False, pulse_times=None):
        """``pulse_times`` (explicit, possibly non-uniform azimuth time of every input line; an
        extension the reference API cannot express) is honoured by the restated port only."""
        if pulse_times is not None and "port" not in self.kind:
            raise RuntimeError("pulse_times needs the restated port oracle (tdbp.port())")
        fl = build_args(out, out_geometry, in_, in_geometry, dem, fc, ds, kernel,
                        dry_tropo_model, rdr2geo_params, geo2rdr_params, 1024, height,
                        pulse_times=pulse_times)
        status = self._backproject(C.byref(fl.args))
        if status < 0:
            raise RuntimeError(f"oracle[{self.kind}] status {status}: "
                               f"{(self._last_error() or b'').decode()}")
        return status if return_status else status != 0

    # -- components -------------------------------------------------------------
    @staticmethod
    def _v3(a):
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(3)
        return a, a.ctypes.data_as(C.POINTER(C.c_double))

    def bistatic_delay(self, p, v, x):
        (_, pp), (_, vp), (_, xp) = self._v3(p), self._v3(v), self._v3(x)
        return self._bistatic(pp, vp, xp)

    def orbit_interpolate(self, orbit, t, border_mode=0):
        fl = _capi.Flattened()
        o = _capi.flatten_orbit(orbit, fl)
        pos, pp = self._v3(np.zeros(3))
        vel, vp = self._v3(np.zeros(3))
        st = self._orbit(C.byref(o), float(t), int(border_mode), pp, vp)
        return st, pos, vel

    def kernel_eval(self, kernel, t):
        fl = _capi.Flattened()
        k = _capi.flatten_kernel(kernel, fl)
        t = np.ascontiguousarray(np.atleast_1d(t), dtype=np.float64)
        out = np.empty(t.size, np.float32)
        self._kernel(C.byref(k), t.ctypes.data_as(C.POINTER(C.c_double)), t.size,
                     out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def tabulate_knab(self, width, bandwidth, n):
        out = np.empty(n, np.float32)
        self._tab(width, bandwidth, n, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def cheby_knab(self, width, bandwidth, n):
        out = np.empty(n, np.float32)
        self._cheb(width, bandwidth, n, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def interp1d(self, kernel, data, t):
        fl = _capi.Flattened()
        k = _capi.flatten_kernel(kernel, fl)
        data = np.ascontiguousarray(data, dtype=np.complex64)
        t = np.ascontiguousarray(np.atleast_1d(t), dtype=np.float64)
        out = np.empty(t.size, np.complex64)
        self._interp1d(C.byref(k), data.ctypes.data_as(C.POINTER(C.c_float)), data.size,
                       t.ctypes.data_as(C.POINTER(C.c_double)), t.size,
                       out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def rdr2geo_bracket(self, t, r, fd, orbit, dem, wvl, side, tol_height=1e-5, look_min=0.0,
                        look_max=np.pi / 2):
        fl = _capi.Flattened()
        o = _capi.flatten_orbit(orbit, fl)
        d = _capi.flatten_dem(dem, fl)
        prm = _capi.Rdr2GeoBracketParams(tol_height, look_min, look_max)
        xyz, xp = self._v3(np.zeros(3))
        ok = self._r2g(t, r, fd, C.byref(o), C.byref(d), wvl, int(side), C.byref(prm), xp)
        return ok, xyz

    def geo2rdr_bracket(self, x, orbit, doppler, wvl, side, tol_aztime=1e-7, time_start=None,
                        time_end=None):
        fl = _capi.Flattened()
        o = _capi.flatten_orbit(orbit, fl)
        l = _capi.flatten_lut2d(doppler, fl)
        prm = _capi.Geo2RdrBracketParams(tol_aztime, int(time_start is not None),
                                         int(time_end is not None), time_start or 0.0,
                                         time_end or 0.0)
        _, xp = self._v3(x)
        t, r = C.c_double(0), C.c_double(0)
        ok = self._g2r(xp, C.byref(o), C.byref(l), wvl, int(side), C.byref(prm), C.byref(t),
                       C.byref(r))
        return ok, t.value, r.value

    def xyz_to_llh(self, x):
        _, xp = self._v3(x)
        out, op = self._v3(np.zeros(3))
        self._x2l(xp, op)
        return out

    def llh_to_xyz(self, llh):
        _, lp = self._v3(llh)
        out, op = self._v3(np.zeros(3))
        self._l2x(lp, op)
        return out

    def dry_tropo_tsx(self, p, llh):
        (_, pp), (_, lp) = self._v3(p), self._v3(llh)
        return self._tropo(pp, lp)

    def brent(self, a, b, fn, tol):
        cb = BRENT_FN(lambda x, _ctx: float(fn(x)))
        root = C.c_double(0)
        st = self._brent(a, b, cb, None, tol, C.byref(root))
        return st, root.value

    def lut2d_eval(self, lut, y, x):
        fl = _capi.Flattened()
        l = _capi.flatten_lut2d(lut, fl)
        return self._lut(C.byref(l), y, x)

    def project_forward(self, epsg, lon, lat):
        """(status, x, y) of createProj(epsg)->forward for lon, lat in radians."""
        xy, xp = self._v3(np.zeros(3))
        st = self._proj(int(epsg), float(lon), float(lat), xp)
        return st, xy[0], xy[1]

    def dem_interp(self, dem, lon, lat):
        fl = _capi.Flattened()
        d = _capi.flatten_dem(dem, fl)
        return self._dem(C.byref(d), lon, lat)


_cache = {}


def port() -> Oracle:
    if "port" not in _cache:
        if not PORT_LIB.exists():
            build()
        _cache["port"] = Oracle(PORT_LIB, "tdbp_oracle")
    return _cache["port"]


def have_ref() -> bool:
    return REF_LIB.exists()


def ref() -> Oracle:
    if "ref" not in _cache:
        if not REF_LIB.exists():
            raise FileNotFoundError(f"{REF_LIB} not built (needs /root/reference; `make -C oracle ref`)")
        _cache["ref"] = Oracle(REF_LIB, "tdbp_ref")
    return _cache["ref"]


class ReferenceCuda:
    """The reference's own CUDA backprojection (cuda/focus/Backproject.cu, unmodified) in the
    reduced harness of oracle/cuda_ref: constant-height DEM and no-data Doppler LUTs only.
    A same-GPU comparator for bench.py and a cross-check in tests/ -- never a product path."""

    kind = "reference CUDA kernels, reduced harness"

    def __init__(self, path: Path = REFCUDA_LIB):
        self.lib = C.CDLL(str(path))
        self._backproject = self.lib.tdbp_refcuda_backproject
        self._backproject.restype = C.c_int
        self._backproject.argtypes = [C.POINTER(_capi.BackprojectArgs)]
        self.lib.tdbp_refcuda_last_error.restype = C.c_char_p

    def backproject(self, out, out_geometry, in_, in_geometry, dem, fc, ds, kernel,
                    dry_tropo_model="tsx", rdr2geo_params=None, geo2rdr_params=None, batch=1024,
                    height=None):
        fl = build_args(out, out_geometry, in_, in_geometry, dem, fc, ds, kernel,
                        dry_tropo_model, rdr2geo_params, geo2rdr_params, batch, height)
        status = self._backproject(C.byref(fl.args))
        if status < 0:
            raise RuntimeError(f"reference CUDA status {status}: "
                               f"{(self.lib.tdbp_refcuda_last_error() or b'').decode()}")
        return status != 0


class Isce3Adapter:
    """integration/isce3/cuda/focus/BackprojectB200.cpp compiled against the reference's
    headers: builds the reference's own objects from the flat arguments and calls
    isce3::cuda::focus::backproject, i.e. the adapter, which calls the product library."""

    def __init__(self, path: Path = ADAPTER_LIB):
        self.lib = C.CDLL(str(path))
        self._backproject = self.lib.tdbp_adapter_backproject
        self._backproject.restype = C.c_int
        self._backproject.argtypes = [C.POINTER(_capi.BackprojectArgs)]
        self.lib.tdbp_adapter_last_error.restype = C.c_char_p

    def backproject(self, out, out_geometry, in_, in_geometry, dem, fc, ds, kernel,
                    dry_tropo_model="tsx", rdr2geo_params=None, geo2rdr_params=None, batch=1024,
                    height=None):
        fl = build_args(out, out_geometry, in_, in_geometry, dem, fc, ds, kernel,
                        dry_tropo_model, rdr2geo_params, geo2rdr_params, batch, height)
        status = self._backproject(C.byref(fl.args))
        if status < 0:
            raise RuntimeError(f"adapter status {status}: "
                               f"{(self.lib.tdbp_adapter_last_error() or b'').decode()}")
        return status != 0


def have_adapter() -> bool:
    return ADAPTER_LIB.exists()


def adapter() -> Isce3Adapter:
    if "adapter" not in _cache:
        _cache["adapter"] = Isce3Adapter()
    return _cache["adapter"]


def have_ref_cuda() -> bool:
    return REFCUDA_LIB.exists()


def ref_cuda() -> ReferenceCuda:
    if "refcuda" not in _cache:
        if not REFCUDA_LIB.exists():
            raise FileNotFoundError(f"{REFCUDA_LIB} not built (`make -C oracle refcuda`)")
        _cache["refcuda"] = ReferenceCuda()
    return _cache["refcuda"]


def best() -> Oracle:
    """The strongest oracle available: the compiled reference if present, else the port."""
    return ref() if have_ref() else port()


def set_threads(n: int):
    """Thread count of the OpenMP oracles.  torchrun exports OMP_NUM_THREADS=1 to its workers,
    which would make the "all host cores" CPU baseline single-threaded: set the environment
    for a runtime that is not initialised yet, and tell an initialised one directly."""
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        gomp = C.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(int(n))
    except OSError:
        pass
