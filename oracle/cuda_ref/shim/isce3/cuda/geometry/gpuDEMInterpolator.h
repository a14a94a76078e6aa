// TEST INFRASTRUCTURE (oracle/_ref/libtdbp_refcuda.so only).  Shadows
// cxx/isce3/cuda/geometry/gpuDEMInterpolator.h (device-side `new` of projection and
// interpolator objects; host twin needs GDAL + pyre).  REDUCED: constant-height DEMs only
// (every bench configuration; raster DEMs throw at construction).
#pragma once
#include <isce3/core/Common.h>
#include <isce3/except/Error.h>
#include <isce3/geometry/DEMInterpolator.h>
namespace isce3 { namespace cuda { namespace geometry {
class gpuDEMInterpolator {
public:
    gpuDEMInterpolator(const isce3::geometry::DEMInterpolator& dem)
        : _ref_height(dem.refHeight()), _epsg(dem.epsgCode())
    {
        if (dem.haveRaster())
            throw isce3::except::RuntimeError(ISCE_SRCINFO(),
                    "reference-CUDA reduced harness: raster DEMs are not supported");
    }
    CUDA_HOSTDEV int epsgCode() const { return _epsg; }
    CUDA_HOSTDEV double refHeight() const { return _ref_height; }
    CUDA_HOSTDEV double interpolateLonLat(double, double) const { return _ref_height; }
private:
    double _ref_height;
    int _epsg;
};
}}}
