// TEST INFRASTRUCTURE (oracle/_ref build only).  Shadows cxx/isce3/core/LUT2d.h
// (whose implementation needs Eigen::Map and pyre) with a class exposing the
// members the TDBP path touches -- eval, contains, haveData, yStart, ySpacing,
// length, boundsError (Backproject.cpp:134; geometry/detail/Geo2Rdr.icc:193-221)
// -- backed by the restated sampler in oracle/tdbp_samplers.h.
#pragma once
#include <isce3/core/Constants.h>
#include <isce3/core/forward.h>
#include <vector>
#include "../../../tdbp_samplers.h"
namespace isce3 { namespace core {
template<typename T>
class LUT2d {
public:
    LUT2d() : _d {} { _d.have_data = 0; _d.ref_value = 0.0; _d.bounds_error = 1; }
    explicit LUT2d(const I3B_LUT2d& d) : _d(d)
    {
        if (d.have_data) {
            _buf.assign(d.data, d.data + d.length * d.width);
            _d.data = _buf.data();
        }
    }
    LUT2d(const LUT2d& o) : _d(o._d), _buf(o._buf) { if (_d.have_data) _d.data = _buf.data(); }
    LUT2d& operator=(const LUT2d& o)
    {
        _d = o._d; _buf = o._buf; if (_d.have_data) _d.data = _buf.data(); return *this;
    }
    bool haveData() const { return _d.have_data != 0; }
    bool boundsError() const { return _d.bounds_error != 0; }
    T refValue() const { return T(_d.ref_value); }
    double xStart() const { return _d.xstart; }
    double yStart() const { return _d.ystart; }
    double xSpacing() const { return _d.dx; }
    double ySpacing() const { return _d.dy; }
    size_t length() const { return size_t(_d.length); }
    size_t width() const { return size_t(_d.width); }
    // accessors the isce3-side adapter (integration/.../BackprojectB200.cpp) reads
    isce3::core::dataInterpMethod interpMethod() const
    {
        return static_cast<isce3::core::dataInterpMethod>(_d.method);
    }
    struct DataView { // stands for Matrix<T>: only .data() is used
        const T* p;
        const T* data() const { return p; }
    };
    DataView data() const { return DataView {reinterpret_cast<const T*>(_d.data)}; }
    bool contains(double y, double x) const { return tdbp_oracle::lut2d_contains(_d, y, x); }
    T eval(double y, double x) const { return T(tdbp_oracle::lut2d_eval(_d, y, x)); }
private:
    I3B_LUT2d _d;
    std::vector<double> _buf;
};
}}
