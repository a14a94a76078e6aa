// TEST INFRASTRUCTURE (oracle/_ref/libtdbp_refcuda.so only).  Shadows
// cxx/isce3/cuda/geometry/gpuGeometry.h: the two device wrappers the TDBP path calls,
// restated from cxx/isce3/cuda/geometry/gpuGeometry.cu:57-68 and :166-181 (they forward to the
// reference's own detail::rdr2geo_bracket / detail::geo2rdr_bracket templates, which are
// included unchanged); the rest of that file is the Newton solvers, unused here.
#pragma once
#include <optional>
#include <isce3/core/Common.h>
#include <isce3/core/Ellipsoid.h>
#include <isce3/core/LookSide.h>
#include <isce3/core/Vector.h>
#include <isce3/cuda/core/OrbitView.h>
#include <isce3/cuda/core/gpuLUT2d.h>
#include <isce3/cuda/geometry/gpuDEMInterpolator.h>
#include <isce3/error/ErrorCode.h>
#include <isce3/geometry/detail/Geo2Rdr.h>
#include <isce3/geometry/detail/Rdr2Geo.h>
namespace isce3 { namespace cuda { namespace geometry {
CUDA_DEV inline int rdr2geo_bracket(double aztime, double slantRange, double doppler,
        const isce3::cuda::core::OrbitView& orbit, const isce3::core::Ellipsoid& ellipsoid,
        const gpuDEMInterpolator& dem, isce3::core::Vec3& targetXYZ, double wvl,
        isce3::core::LookSide side, double tolHeight, double lookMin, double lookMax)
{
    auto status = isce3::geometry::detail::rdr2geo_bracket(&targetXYZ, aztime, slantRange,
            doppler, orbit, dem, ellipsoid, wvl, side, {tolHeight, lookMin, lookMax});
    return status == isce3::error::ErrorCode::Success;
}
CUDA_DEV inline int geo2rdr_bracket(const isce3::core::Vec3& x,
        const isce3::cuda::core::OrbitView& orbit,
        const isce3::cuda::core::gpuLUT2d<double>& doppler, double* aztime, double* range,
        const double wavelength, const isce3::core::LookSide side, const double dt,
        std::optional<double> timeStart, std::optional<double> timeEnd)
{
    auto err = isce3::geometry::detail::geo2rdr_bracket(aztime, range, x, orbit, doppler,
            wavelength, side, {dt, timeStart, timeEnd});
    return err == isce3::error::ErrorCode::Success;
}
}}}
