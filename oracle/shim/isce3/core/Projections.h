// TEST INFRASTRUCTURE (oracle/_ref build only).  Shadows
// cxx/isce3/core/Projections.h: the TDBP path only asks a projection for its
// ellipsoid (Backproject.cpp:115-116, rdr2geo_roots.cpp:19-21), which is WGS84
// for every EPSG code createProj accepts (Projections.h:36-39).
#pragma once
#include <memory>
#include <isce3/core/Ellipsoid.h>
namespace isce3 { namespace core {
class ProjectionBase {
public:
    explicit ProjectionBase(int code)
        : _epsgcode(code), _ellipse(EarthSemiMajorAxis, EarthEccentricitySquared) {}
    int code() const { return _epsgcode; }
    const Ellipsoid& ellipsoid() const { return _ellipse; }
private:
    int _epsgcode;
    Ellipsoid _ellipse;
};
inline std::unique_ptr<ProjectionBase> makeProjection(int epsg)
{
    return std::make_unique<ProjectionBase>(epsg);
}
}}
