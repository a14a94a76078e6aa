// Adapter a maintainer adds to isce3 as cxx/isce3/cuda/focus/BackprojectB200.cpp (in place of
// cuda/focus/Backproject.cu in the isce3-cuda target): keeps the public signature of
// isce3::cuda::focus::backproject (cxx/isce3/cuda/focus/Backproject.h:78-88), flattens the
// isce3 value types into the C-ABI of include/isce3_b200_backproject.h and maps status codes
// back to the reference's ErrorCode / exceptions.  See INTEGRATION.md.
//
// isce3 itself cannot be built in this repo's container; `make -C oracle adapter` compiles
// this file against the reference's own headers (plus the oracle/shim stand-ins for the
// Eigen/GDAL-backed LUT2d and DEMInterpolator) and tests/test_gpu_parity.py calls it.
#include <isce3/cuda/focus/Backproject.h>      // keeps the public declaration
#include <isce3_b200_backproject.h>
#include <isce3/container/RadarGeometry.h>
#include <isce3/core/Kernels.h>
#include <isce3/core/TimeDelta.h>
#include <isce3/except/Error.h>
#include <isce3/geometry/DEMInterpolator.h>
#include <cmath>
#include <complex>
#include <string>
#include <typeinfo>

namespace isce3::cuda::focus {
namespace {

I3B_RadarGeometry flatten(const isce3::container::RadarGeometry& g)
{
    I3B_RadarGeometry f {};
    const auto& rg = g.radarGrid();             // already re-based to the orbit epoch
    f.grid = {rg.sensingStart(), rg.prf(), rg.startingRange(), rg.rangePixelSpacing(),
              rg.wavelength(), (int64_t) rg.length(), (int64_t) rg.width(),
              rg.lookSide() == isce3::core::LookSide::Right ? I3B_LOOK_RIGHT : I3B_LOOK_LEFT, 0};
    const auto& o = g.orbit();                  // std::vector<Vec3> is 3 contiguous doubles each
    f.orbit = {o.time().first(), o.time().spacing(), o.size(),
               o.interpMethod() == isce3::core::OrbitInterpMethod::Legendre ? I3B_ORBIT_LEGENDRE
                                                                           : I3B_ORBIT_HERMITE,
               o.position()[0].data(), o.velocity()[0].data()};
    const auto& d = g.doppler();
    f.doppler.have_data = d.haveData();
    f.doppler.bounds_error = d.boundsError();
    f.doppler.method = (int) d.interpMethod();  // dataInterpMethod values are shared
    f.doppler.ref_value = d.refValue();
    if (d.haveData()) {
        f.doppler.length = d.length();  f.doppler.width = d.width();
        f.doppler.xstart = d.xStart();  f.doppler.ystart = d.yStart();
        f.doppler.dx = d.xSpacing();    f.doppler.dy = d.ySpacing();
        f.doppler.data = d.data().data();       // Matrix<double>, row-major
    }
    const auto dt = g.referenceEpoch() - isce3::core::DateTime(1970, 1, 1);
    const double sec = std::floor(dt.getTotalSeconds());
    f.ref_epoch_sec = (int64_t) sec;
    f.ref_epoch_frac = dt.getTotalSeconds() - sec;
    return f;
}

I3B_Kernel flatten(const isce3::core::Kernel<float>& k)
{
    using namespace isce3::core;
    I3B_Kernel f {};
    f.width = k.width();
    if (typeid(k) == typeid(LinearKernel<float>))        f.kind = I3B_KERNEL_LINEAR;
    else if (typeid(k) == typeid(BartlettKernel<float>)) f.kind = I3B_KERNEL_BARTLETT;
    else if (auto p = dynamic_cast<const KnabKernel<float>*>(&k)) {
        f.kind = I3B_KERNEL_KNAB;  f.bandwidth = p->bandwidth();
    } else if (auto p = dynamic_cast<const TabulatedKernel<float>*>(&k)) {
        f.kind = I3B_KERNEL_TABULATED;  f.n = (int) p->table().size();  f.data = p->table().data();
    } else if (auto p = dynamic_cast<const ChebyKernel<float>*>(&k)) {
        f.kind = I3B_KERNEL_CHEBY;  f.n = (int) p->coeffs().size();  f.data = p->coeffs().data();
    } else {
        throw isce3::except::RuntimeError(ISCE_SRCINFO(), "not implemented");   // Backproject.cu:750-752
    }
    return f;
}

} // namespace

isce3::error::ErrorCode
backproject(std::complex<float>* out, const isce3::container::RadarGeometry& out_geometry,
            const std::complex<float>* in, const isce3::container::RadarGeometry& in_geometry,
            const isce3::geometry::DEMInterpolator& dem, double fc, double ds,
            const isce3::core::Kernel<float>& kernel, DryTroposphereModel dry_tropo_model,
            const isce3::geometry::detail::Rdr2GeoBracketParams& r2g,
            const isce3::geometry::detail::Geo2RdrBracketParams& g2r, int batch, float* height)
{
    // Backproject.cpp:88-92 / Backproject.cu:480-484, compared on the DateTime values themselves
    // (the descriptor carries the epoch as seconds + fraction, a double near 1.7e9 only
    // resolves ~2.4e-7 s)
    if (out_geometry.referenceEpoch() != in_geometry.referenceEpoch()) {
        throw isce3::except::RuntimeError(ISCE_SRCINFO(),
                "input reference epoch must match output reference epoch");
    }
    I3B_BackprojectArgs a {};
    a.abi_version = I3B_ABI_VERSION;
    a.out = reinterpret_cast<float*>(out);
    a.in = reinterpret_cast<const float*>(in);
    a.height = height;
    a.out_geometry = flatten(out_geometry);
    a.in_geometry = flatten(in_geometry);
    a.dem.have_raster = dem.haveRaster();
    a.dem.epsg = dem.epsgCode();
    a.dem.method = (int) dem.interpMethod();
    a.dem.ref_height = dem.refHeight();
    if (dem.haveRaster()) {
        a.dem.length = dem.length();   a.dem.width = dem.width();
        a.dem.xstart = dem.xStart();   a.dem.ystart = dem.yStart();
        a.dem.dx = dem.deltaX();       a.dem.dy = dem.deltaY();
        a.dem.data = dem.data();
    }
    a.fc = fc;  a.ds = ds;
    a.kernel = flatten(kernel);
    a.dry_tropo_model = dry_tropo_model == DryTroposphereModel::TSX ? I3B_TROPO_TSX : I3B_TROPO_NODELAY;
    a.batch = batch;
    a.rdr2geo = {r2g.tol_height, r2g.look_min, r2g.look_max};
    a.geo2rdr = {g2r.tol_aztime, g2r.time_start.has_value(), g2r.time_end.has_value(),
                 g2r.time_start.value_or(0.0), g2r.time_end.value_or(0.0)};
    // a.n_devices = 0: the current CUDA device, as the reference does (focus.py:1589-1595);
    // list devices here to shard the block over several GPUs.
    const int st = i3b_backproject(&a);
    if (st >= 0) return static_cast<isce3::error::ErrorCode>(st);
    const std::string msg = i3b_last_error();
    switch (st) {
    case I3B_EXC_INVALID_ARGUMENT: throw isce3::except::InvalidArgument(ISCE_SRCINFO(), msg);
    case I3B_EXC_DOMAIN_ERROR:     throw isce3::except::DomainError(ISCE_SRCINFO(), msg);
    case I3B_EXC_OVERFLOW_ERROR:   throw isce3::except::OverflowError(ISCE_SRCINFO(), msg);
    case I3B_EXC_OUT_OF_RANGE:     throw isce3::except::OutOfRange(ISCE_SRCINFO(), msg);
    default:                       throw isce3::except::RuntimeError(ISCE_SRCINFO(), msg);
    }
}

} // namespace isce3::cuda::focus
