// TEST INFRASTRUCTURE (oracle/_ref build only).  Shadows
// cxx/isce3/geometry/DEMInterpolator.h (needs GDAL + pyre) with the three
// members the TDBP path calls -- epsgCode(), refHeight(), interpolateLonLat()
// (Backproject.cpp:115; geometry/detail/Rdr2Geo.icc:223-228) -- backed by the
// restated sampler in oracle/tdbp_samplers.h; the map projection in front of it is the
// reference's own (core/Projections.cpp compiled unchanged, createProj :373-402).
#pragma once
#include <memory>
#include <isce3/core/Projections.h>
#include <isce3/core/Constants.h>
#include <isce3/core/forward.h>
#include <isce3/geometry/forward.h>
#include "../../../tdbp_samplers.h"
namespace isce3 { namespace geometry {
class DEMInterpolator {
public:
    explicit DEMInterpolator(const I3B_DEM& d) : _d(d)
    {
        if (d.have_raster) _proj.reset(isce3::core::createProj(d.epsg));
    }
    int epsgCode() const { return _d.epsg; }
    double refHeight() const { return _d.ref_height; }
    bool haveRaster() const { return _d.have_raster != 0; }
    // accessors the isce3-side adapter reads (geometry/DEMInterpolator.h:96-188)
    isce3::core::dataInterpMethod interpMethod() const
    {
        return static_cast<isce3::core::dataInterpMethod>(_d.method);
    }
    double xStart() const { return _d.xstart; }
    double yStart() const { return _d.ystart; }
    double deltaX() const { return _d.dx; }
    double deltaY() const { return _d.dy; }
    size_t width() const { return size_t(_d.width); }
    size_t length() const { return size_t(_d.length); }
    const float* data() const { return _d.data; }
    double interpolateLonLat(double lon, double lat) const
    {
        // DEMInterpolator.cpp:592-611
        if (!_d.have_raster) return _d.ref_height;
        isce3::core::Vec3 xyz;
        const isce3::core::Vec3 llh {lon, lat, 0.0};
        if (_proj->forward(llh, xyz) != 0) return _d.ref_height; // (reference: unset point)
        return tdbp_oracle::dem_interp_xy(_d, xyz[0], xyz[1]);
    }
private:
    I3B_DEM _d;
    std::shared_ptr<isce3::core::ProjectionBase> _proj;
};
}}
