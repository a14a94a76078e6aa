// TEST INFRASTRUCTURE (oracle/_ref build only).  Shadows
// cxx/isce3/geometry/DEMInterpolator.h (needs GDAL + pyre) with the three
// members the TDBP path calls -- epsgCode(), refHeight(), interpolateLonLat()
// (Backproject.cpp:115; geometry/detail/Rdr2Geo.icc:223-228) -- backed by the
// restated sampler in oracle/tdbp_samplers.h.
#pragma once
#include <isce3/core/forward.h>
#include <isce3/geometry/forward.h>
#include "../../../tdbp_samplers.h"
namespace isce3 { namespace geometry {
class DEMInterpolator {
public:
    explicit DEMInterpolator(const I3B_DEM& d) : _d(d) {}
    int epsgCode() const { return _d.epsg; }
    double refHeight() const { return _d.ref_height; }
    bool haveRaster() const { return _d.have_raster != 0; }
    double interpolateLonLat(double lon, double lat) const
    {
        return tdbp_oracle::dem_interp_lonlat(_d, lon, lat);
    }
private:
    I3B_DEM _d;
};
}}
