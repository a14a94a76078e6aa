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
        if self.interp_method == DataInterpMethod.SINC:
            raise ValueError("sinc DEM interpolation is not supported on the TDBP path")
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
