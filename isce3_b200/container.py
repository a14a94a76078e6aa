"""isce3.container.RadarGeometry stand-in (cxx/isce3/container/RadarGeometry.h:16-64,
.icc:7-58; python/extensions/pybind_isce3/container/RadarGeometry.cpp:7-31)."""
from __future__ import annotations

from .core import Linspace
from .product import RadarGridParameters


class RadarGeometry:
    def __init__(self, radar_grid: RadarGridParameters, orbit, doppler):
        grid = radar_grid.copy()
        # re-base the grid to the orbit's reference epoch (RadarGeometry.icc:14-25)
        if grid.ref_epoch != orbit.reference_epoch:
            dt = grid.ref_epoch - orbit.reference_epoch
            grid.sensing_start = grid.sensing_start + dt
            grid.ref_epoch = orbit.reference_epoch
        self.radar_grid = grid
        self.orbit = orbit
        self.doppler = doppler

    @property
    def reference_epoch(self):
        return self.orbit.reference_epoch

    @property
    def grid_length(self):
        return self.radar_grid.length

    @property
    def grid_width(self):
        return self.radar_grid.width

    @property
    def sensing_time(self):
        g = self.radar_grid
        if g.length > 2**31 - 1:
            raise OverflowError("grid length exceeds max int")  # RadarGeometry.icc:30-36
        return Linspace(g.sensing_start, g.az_time_interval, g.length)

    @property
    def slant_range(self):
        g = self.radar_grid
        if g.width > 2**31 - 1:
            raise OverflowError("grid width exceeds max int")
        return Linspace(g.starting_range, g.range_pixel_spacing, g.width)

    @property
    def look_side(self):
        return self.radar_grid.lookside

    @property
    def wavelength(self):
        return self.radar_grid.wavelength
