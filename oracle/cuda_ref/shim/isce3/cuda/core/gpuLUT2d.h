// TEST INFRASTRUCTURE (oracle/_ref/libtdbp_refcuda.so only).  Shadows
// cxx/isce3/cuda/core/gpuLUT2d.h, whose implementation allocates a virtual interpolator with
// device-side `new` and needs the real (Eigen-backed) host LUT2d.  REDUCED: only the
// "no data" LUT (eval == refValue, the zero-Doppler grids of focus.py:1579,1998 and of every
// bench configuration) is supported; a LUT with data throws at construction.
#pragma once
#include <isce3/core/Common.h>
#include <isce3/core/LUT2d.h>
#include <isce3/except/Error.h>
namespace isce3 { namespace cuda { namespace core {
template<typename T>
class gpuLUT2d {
public:
    gpuLUT2d(const isce3::core::LUT2d<T>& lut) : _ref(lut.refValue()), _bounds_error(lut.boundsError())
    {
        if (lut.haveData())
            throw isce3::except::RuntimeError(ISCE_SRCINFO(),
                    "reference-CUDA reduced harness: Doppler LUTs with data are not supported");
    }
    CUDA_HOSTDEV bool haveData() const { return false; }
    CUDA_HOSTDEV bool boundsError() const { return _bounds_error; }
    CUDA_HOSTDEV T refValue() const { return _ref; }
    CUDA_HOSTDEV double xStart() const { return 0.; }
    CUDA_HOSTDEV double yStart() const { return 0.; }
    CUDA_HOSTDEV double xSpacing() const { return 1.; }
    CUDA_HOSTDEV double ySpacing() const { return 1.; }
    CUDA_HOSTDEV size_t length() const { return 0; }
    CUDA_HOSTDEV size_t width() const { return 0; }
    CUDA_HOSTDEV bool contains(double, double) const { return true; }
    CUDA_HOSTDEV T eval(double, double) const { return _ref; }
private:
    T _ref;
    bool _bounds_error;
};
}}}
