// TEST INFRASTRUCTURE (oracle/_ref build only).  Shadows cxx/isce3/core/Basis.h.
// geometry/detail/Rdr2Geo.icc:4 includes it for the Newton rdr2geo template,
// which backproject never instantiates (it calls rdr2geo_bracket).  Only the
// declarations that template's signature names are provided.
#pragma once
#include <isce3/core/forward.h>
#include <isce3/core/Vector.h>
namespace isce3 { namespace core {
class Basis {
public:
    Basis() {}
    Basis(const Vec3&, const Vec3&) {}
    const Vec3& x0() const { return _x[0]; }
    const Vec3& x1() const { return _x[1]; }
    const Vec3& x2() const { return _x[2]; }
private:
    Vec3 _x[3];
};
}}
