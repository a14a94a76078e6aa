"""isce3.geometry.DEMInterpolator stand-in (cxx/isce3/geometry/DEMInterpolator.h:31-52;
python/extensions/pybind_isce3/geometry/DEMInterpolator.cpp:22-120).  The reference has
no numpy constructor (rasters arrive through GDAL); ``from_array`` fills that gap for
synthetic DEMs."""
from __future__ import annotations

import numpy as np

from .core import DataInterpMethod, parse_interp_method


class DEMInterpolator:
    def __init__(self, height=0.0, method="bilinear", epsg=4326):
        self.ref_height = float(height)
        self.interp_method = parse_interp_method(method)
        self.epsg_code = int(epsg)
        self.have_raster = False
        self.data = None
        self.x_start = self.y_start = 0.0
        self.delta_x = self.delta_y = 1.0

    @classmethod
    def from_array(cls, data, x_start, y_start, delta_x, delta_y, epsg=4326,
                   method="biquintic", ref_height=None):
        """Raster DEM: data[row, col] at (x_start + col*delta_x, y_start + row*delta_y),
        pixel centres; x = longitude / y = latitude in degrees for EPSG:4326, projected metres
        for the other codes createProj knows (UTM 326xx/327xx, 3031, 3413, 6933;
        core/Projections.cpp:373-402)."""
        data = np.ascontiguousarray(data, dtype=np.float32)
        if data.ndim != 2:
            raise ValueError("DEM raster must be 2-D")
        from .projections import make_projection
        make_projection(int(epsg))  # raises for codes createProj does not know
        self = cls(float(np.mean(data)) if ref_height is None else ref_height, method, epsg)
        self.data = data
        self.have_raster = True
        self.x_start, self.y_start = float(x_start), float(y_start)
        self.delta_x, self.delta_y = float(delta_x), float(delta_y)
        return self

    @property
    def width(self):
        return 0 if self.data is None else self.data.shape[1]

    @property
    def length(self):
        return 0 if self.data is None else self.data.shape[0]


# ---- per-point geometry on the GPU (batch API of include/isce3_b200_backproject.h) ---------

def rdr2geo_bracket(aztime, slant_range, doppler, orbit, dem, wavelength, side,
                    tol_height=1e-5, look_min=0.0, look_max=np.pi / 2):
    """Arrays of radar coordinates -> target ECEF positions on the DEM, per point like
    ``isce3.geometry.rdr2geo_bracket`` (geometry/rdr2geo_roots.cpp:14-27,
    python/extensions/pybind_isce3/geometry/rdr2geo_roots.cpp): returns ``(xyz[n,3], status[n])``
    with ``status`` the per-point ErrorCode (0 = converged; failed points are NaN)."""
    import ctypes as C

    from . import _capi
    from .focus import raise_for_status
    t = np.ascontiguousarray(np.atleast_1d(aztime), dtype=np.float64)
    r = np.ascontiguousarray(np.broadcast_to(np.asarray(slant_range, np.float64), t.shape))
    fd = None if doppler is None else np.ascontiguousarray(np.broadcast_to(np.asarray(doppler, np.float64), t.shape))
    n = t.size
    xyz = np.empty((n, 3), np.float64)
    status = np.empty(n, np.int32)
    fl = _capi.Flattened()
    o = _capi.flatten_orbit(orbit, fl)
    d = _capi.flatten_dem(dem, fl)
    prm = _capi.Rdr2GeoBracketParams(tol_height, look_min, look_max)
    lib = _capi.load_library()
    rc = lib.i3b_rdr2geo_bracket_batch(C.byref(o), C.byref(d), C.c_double(wavelength), int(side), C.byref(prm),
                                       C.c_int64(n), t.ctypes.data, r.ctypes.data,
                                       fd.ctypes.data if fd is not None else None, xyz.ctypes.data,
                                       status.ctypes.data, 0)
    if rc < 0:
        raise_for_status(rc, (lib.i3b_last_error() or b"").decode())
    return xyz, status


def geo2rdr_bracket(xyz, orbit, doppler, wavelength, side, tol_aztime=1e-7, time_start=None,
                    time_end=None):
    """Arrays of ECEF targets -> (aztime[n], slant_range[n], status[n]), per point like
    ``isce3.geometry.geo2rdr_bracket`` (geometry/geo2rdr_roots.cpp:16-25)."""
    import ctypes as C

    from . import _capi
    from .focus import raise_for_status
    x = np.ascontiguousarray(np.asarray(xyz, np.float64).reshape(-1, 3))
    n = x.shape[0]
    t = np.empty(n, np.float64)
    r = np.empty(n, np.float64)
    status = np.empty(n, np.int32)
    fl = _capi.Flattened()
    o = _capi.flatten_orbit(orbit, fl)
    lut = _capi.flatten_lut2d(doppler, fl)
    prm = _capi.Geo2RdrBracketParams(tol_aztime, int(time_start is not None), int(time_end is not None),
                                     time_start or 0.0, time_end or 0.0)
    lib = _capi.load_library()
    rc = lib.i3b_geo2rdr_bracket_batch(C.byref(o), C.byref(lut), C.c_double(wavelength), int(side), C.byref(prm),
                                       C.c_int64(n), x.ctypes.data, t.ctypes.data, r.ctypes.data,
                                       status.ctypes.data, 0)
    if rc < 0:
        raise_for_status(rc, (lib.i3b_last_error() or b"").decode())
    return t, r, status
