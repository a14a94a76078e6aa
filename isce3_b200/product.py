"""isce3.product.RadarGridParameters stand-in
(cxx/isce3/product/RadarGridParameters.h:43-163;
python/extensions/pybind_isce3/product/RadarGridParameters.cpp:32-65)."""
from __future__ import annotations

import numpy as np

from .core import DateTime, LookSide, parse_look_side


class RadarGridParameters:
    def __init__(self, sensing_start, wavelength, prf, starting_range, range_pixel_spacing,
                 lookside, length, width, ref_epoch: DateTime):
        self.sensing_start = float(sensing_start)
        self.wavelength = float(wavelength)
        self.prf = float(prf)
        self.starting_range = float(starting_range)
        self.range_pixel_spacing = float(range_pixel_spacing)
        self.lookside = parse_look_side(lookside)
        if int(length) < 0 or int(width) < 0:
            raise ValueError("grid dimensions must be non-negative")
        self.length = int(length)
        self.width = int(width)
        self.ref_epoch = ref_epoch

    @property
    def az_time_interval(self):
        return 1.0 / self.prf

    @property
    def shape(self):
        return (self.length, self.width)

    @property
    def sensing_times(self):
        return self.sensing_start + np.arange(self.length) / self.prf

    @property
    def slant_ranges(self):
        return self.starting_range + np.arange(self.width) * self.range_pixel_spacing

    def copy(self):
        return RadarGridParameters(self.sensing_start, self.wavelength, self.prf,
                                   self.starting_range, self.range_pixel_spacing, self.lookside,
                                   self.length, self.width, self.ref_epoch)

    def __getitem__(self, key):
        """grid[a0:a1, r0:r1] -> sub-grid (RadarGridParameters.h:165-179 offsetAndResize)."""
        ka, kr = key
        a0, a1, sa = ka.indices(self.length)
        r0, r1, sr = kr.indices(self.width)
        if sa != 1 or sr != 1:
            raise ValueError("strided sub-grids are not supported")
        g = self.copy()
        g.sensing_start = self.sensing_start + a0 / self.prf
        g.starting_range = self.starting_range + r0 * self.range_pixel_spacing
        g.length, g.width = max(a1 - a0, 0), max(r1 - r0, 0)
        return g
