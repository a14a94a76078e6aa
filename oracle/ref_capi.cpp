// TEST INFRASTRUCTURE -- builds into oracle/_ref/libtdbp_ref.so.  Not product
// code: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load it.
//
// C wrapper around the UNMODIFIED reference sources, compiled where they lie
// under /root/reference/cxx (see oracle/Makefile):
//
//   isce3/focus/Backproject.cpp          isce3::focus::backproject (:65-212)
//   isce3/geometry/rdr2geo_roots.cpp     rdr2geo_bracket
//   isce3/geometry/geo2rdr_roots.cpp     geo2rdr_bracket
//   isce3/core/Orbit.cpp, detail/BuildOrbit.cpp, DateTime.cpp, TimeDelta.cpp
//   isce3/except/Error.cpp, isce3/error/ErrorCode.cpp
//   isce3/focus/DryTroposphereModel.cpp
//   + every header-only template those pull in (Interp1d, Kernels, Brent,
//     Rdr2Geo/Geo2Rdr, InterpolateOrbit, Ellipsoid, BistaticDelay, ...)
//
// What is NOT the reference's own code in this library: the headers under
// oracle/shim/ (fixed-size vector algebra standing in for Eigen 3.3.7, and
// LUT2d / DEMInterpolator classes backed by the restated
// samplers of oracle/tdbp_samplers.h), because Eigen, GDAL and pyre are absent
// from this image.  This file only flattens/unflattens arguments.
#include <cmath>
#include <complex>
#include <cstring>
#include <memory>
#include <optional>
#include <string>
#include <vector>

#include <isce3/container/RadarGeometry.h>
#include <isce3/core/DateTime.h>
#include <isce3/core/Ellipsoid.h>
#include <isce3/core/Interp1d.h>
#include <isce3/core/Kernels.h>
#include <isce3/core/LUT2d.h>
#include <isce3/core/Orbit.h>
#include <isce3/core/Projections.h>
#include <isce3/core/TimeDelta.h>
#include <isce3/error/ErrorCode.h>
#include <isce3/except/Error.h>
#include <isce3/focus/Backproject.h>
#include <isce3/focus/BistaticDelay.h>
#include <isce3/focus/DryTroposphereModel.h>
#include <isce3/geometry/DEMInterpolator.h>
#include <isce3/geometry/geo2rdr_roots.h>
#include <isce3/geometry/rdr2geo_roots.h>
#include <isce3/math/RootFind1dBracket.h>
#include <isce3/product/RadarGridParameters.h>

#include "../include/isce3_b200_backproject.h"

using namespace isce3::core;
using isce3::error::ErrorCode;

namespace {

thread_local std::string g_err;

DateTime make_epoch(int64_t sec, double frac)
{
    return DateTime(1970, 1, 1) + TimeDelta(double(sec)) + TimeDelta(frac);
}

Orbit make_orbit(const I3B_Orbit& o, const DateTime& epoch)
{
    std::vector<StateVector> sv(o.n);
    for (int i = 0; i < o.n; ++i) {
        sv[i].datetime = epoch + TimeDelta(o.t0 + i * o.dt);
        sv[i].position = Vec3(o.pos[3 * i], o.pos[3 * i + 1], o.pos[3 * i + 2]);
        sv[i].velocity = Vec3(o.vel[3 * i], o.vel[3 * i + 1], o.vel[3 * i + 2]);
    }
    const auto method = o.method == I3B_ORBIT_LEGENDRE ? OrbitInterpMethod::Legendre
                                                       : OrbitInterpMethod::Hermite;
    Orbit orbit(sv, epoch, method);
    // The orbit's time axis is rebuilt from DateTime differences
    // (core/detail/BuildOrbit.cpp:11-48); insist that it reproduces the
    // caller's axis exactly so both oracles and the GPU see identical inputs.
    if (orbit.time().first() != o.t0 || orbit.time().spacing() != o.dt) {
        throw isce3::except::InvalidArgument(ISCE_SRCINFO(),
                "oracle/_ref: state vector times must be exactly representable "
                "through isce3::core::DateTime (use binary-friendly t0/dt)");
    }
    return orbit;
}

isce3::product::RadarGridParameters make_grid(const I3B_RadarGrid& g, const DateTime& epoch)
{
    const auto side = g.look_side == I3B_LOOK_RIGHT ? LookSide::Right : LookSide::Left;
    return isce3::product::RadarGridParameters(g.sensing_start, g.wavelength, g.prf,
            g.starting_range, g.range_pixel_spacing, side, size_t(g.length),
            size_t(g.width), epoch);
}

isce3::container::RadarGeometry make_geometry(const I3B_RadarGeometry& g)
{
    const DateTime epoch = make_epoch(g.ref_epoch_sec, g.ref_epoch_frac);
    return isce3::container::RadarGeometry(
            make_grid(g.grid, epoch), make_orbit(g.orbit, epoch), LUT2d<double>(g.doppler));
}

// A Kernel<float> whose samples were tabulated by the caller: same state as
// TabulatedKernel<float> (core/Kernels.h:128-147) but filled from a table
// instead of from another kernel.  operator() is the reference's
// (core/Kernels.icc:139-154) via a real TabulatedKernel built from a kernel
// that replays the table.
class ReplayKernel : public Kernel<float> {
public:
    ReplayKernel(const float* t, int n, double width) : Kernel<float>(width), _t(t), _n(n) {}
    float operator()(double x) const override
    {
        // called by TabulatedKernel's ctor at x = i*dx, i = 0..n-1
        const double dx = this->_halfwidth / (_n - 1.0);
        long i = std::lround(x / dx);
        i = std::min<long>(std::max<long>(i, 0), _n - 1);
        return _t[i];
    }
private:
    const float* _t;
    int _n;
};

class ReplayCheby : public Kernel<float> {
public:
    // ChebyKernel has no way to be filled from coefficients; restate its
    // operator() (core/Kernels.icc:191-211) around caller-provided coeffs.
    ReplayCheby(const float* c, int n, double width)
        : Kernel<float>(width), _c(c, c + n), _scale(float(4.0 / width)) {}
    float operator()(double x) const override
    {
        const auto ax = std::abs(x);
        if (ax > this->_halfwidth) return 0.f;
        const float q = (ax * _scale) - 1.f;
        const float twoq = 2.f * q;
        const int n = int(_c.size());
        float bk = 0, bk1 = 0, bk2 = 0;
        for (int i = n - 1; i > 0; --i) {
            bk = _c[i] + twoq * bk1 - bk2;
            bk2 = bk1;
            bk1 = bk;
        }
        return _c[0] + q * bk1 - bk2;
    }
private:
    std::vector<float> _c;
    float _scale;
};

std::unique_ptr<Kernel<float>> make_kernel(const I3B_Kernel& k)
{
    switch (k.kind) {
    case I3B_KERNEL_BARTLETT: return std::make_unique<BartlettKernel<float>>(k.width);
    case I3B_KERNEL_LINEAR: return std::make_unique<LinearKernel<float>>();
    case I3B_KERNEL_KNAB: return std::make_unique<KnabKernel<float>>(k.width, k.bandwidth);
    case I3B_KERNEL_TABULATED: {
        ReplayKernel rk(k.data, k.n, k.width);
        return std::make_unique<TabulatedKernel<float>>(rk, k.n);
    }
    case I3B_KERNEL_CHEBY: return std::make_unique<ReplayCheby>(k.data, k.n, k.width);
    default: throw isce3::except::RuntimeError(ISCE_SRCINFO(), "not implemented");
    }
}

template<class F>
int guarded(F&& f)
{
    try {
        return f();
    } catch (const isce3::except::InvalidArgument& e) {
        g_err = e.what();
        return I3B_EXC_INVALID_ARGUMENT;
    } catch (const isce3::except::DomainError& e) {
        g_err = e.what();
        return I3B_EXC_DOMAIN_ERROR;
    } catch (const isce3::except::OutOfRange& e) {
        g_err = e.what();
        return I3B_EXC_OUT_OF_RANGE;
    } catch (const isce3::except::OverflowError& e) {
        g_err = e.what();
        return I3B_EXC_OVERFLOW_ERROR;
    } catch (const std::exception& e) {
        g_err = e.what();
        return I3B_EXC_RUNTIME_ERROR;
    }
}

} // namespace

extern "C" {

const char* tdbp_ref_last_error() { return g_err.c_str(); }

const char* tdbp_ref_kind() { return "reference"; }

// isce3::focus::backproject, cxx/isce3/focus/Backproject.cpp:65-212
int tdbp_ref_backproject(const I3B_BackprojectArgs* a)
{
    return guarded([&]() {
        const auto out_geom = make_geometry(a->out_geometry);
        const auto in_geom = make_geometry(a->in_geometry);
        const isce3::geometry::DEMInterpolator dem(a->dem);
        const auto kernel = make_kernel(a->kernel);
        isce3::geometry::detail::Rdr2GeoBracketParams r2g {
                a->rdr2geo.tol_height, a->rdr2geo.look_min, a->rdr2geo.look_max};
        isce3::geometry::detail::Geo2RdrBracketParams g2r;
        g2r.tol_aztime = a->geo2rdr.tol_aztime;
        if (a->geo2rdr.has_time_start) g2r.time_start = a->geo2rdr.time_start;
        if (a->geo2rdr.has_time_end) g2r.time_end = a->geo2rdr.time_end;
        const auto model = static_cast<isce3::focus::DryTroposphereModel>(a->dry_tropo_model);
        const ErrorCode ec = isce3::focus::backproject(
                reinterpret_cast<std::complex<float>*>(a->out), out_geom,
                reinterpret_cast<const std::complex<float>*>(a->in), in_geom, dem, a->fc,
                a->ds, *kernel, model, r2g, g2r, a->height);
        return static_cast<int>(ec);
    });
}

// focus/BistaticDelay.icc:10-17
double tdbp_ref_bistatic_delay(const double* p, const double* v, const double* x)
{
    return isce3::focus::bistaticDelay(Vec3(p[0], p[1], p[2]), Vec3(v[0], v[1], v[2]),
                                       Vec3(x[0], x[1], x[2]));
}

// core/Orbit.cpp:71-86 -> core/detail/InterpolateOrbit.icc:161-193
// border_mode: 0 Error (throws -> negative status), 1 Extrapolate, 2 FillNaN
int tdbp_ref_orbit_interpolate(const I3B_Orbit* o, double t, int border_mode, double* pos,
                               double* vel)
{
    return guarded([&]() {
        const Orbit orbit = make_orbit(*o, DateTime(2000, 1, 1));
        Vec3 p, v;
        const auto ec = orbit.interpolate(&p, &v, t, static_cast<OrbitInterpBorderMode>(border_mode));
        for (int i = 0; i < 3; ++i) {
            pos[i] = p[i];
            vel[i] = v[i];
        }
        return static_cast<int>(ec);
    });
}

// core/Kernels.icc (Kernel<float>::operator())
int tdbp_ref_kernel_eval(const I3B_Kernel* k, const double* t, int n, float* out)
{
    return guarded([&]() {
        const auto kernel = make_kernel(*k);
        for (int i = 0; i < n; ++i) out[i] = (*kernel)(t[i]);
        return 0;
    });
}

// TabulatedKernel<float>(KnabKernel<double>(width, bandwidth), n).table()
// core/Kernels.icc:114-137 ; what focus.py:796-808 and the reference test build
int tdbp_ref_tabulate_knab(double width, double bandwidth, int n, float* table)
{
    return guarded([&]() {
        const KnabKernel<double> knab(width, bandwidth);
        const TabulatedKernel<float> tab(knab, n);
        std::memcpy(table, tab.table().data(), sizeof(float) * n);
        return 0;
    });
}

// ChebyKernel<float>(KnabKernel<double>(width, bandwidth), n).coeffs()
// core/Kernels.icc:156-189
int tdbp_ref_cheby_knab(double width, double bandwidth, int n, float* coeffs)
{
    return guarded([&]() {
        const KnabKernel<double> knab(width, bandwidth);
        const ChebyKernel<float> ch(knab, n);
        std::memcpy(coeffs, ch.coeffs().data(), sizeof(float) * n);
        return 0;
    });
}

// core/Interp1d.icc:7-22 on complex<float> data, stride 1, non-periodic
int tdbp_ref_interp1d(const I3B_Kernel* k, const float* data, size_t n, const double* t,
                      int nt, float* out)
{
    return guarded([&]() {
        const auto kernel = make_kernel(*k);
        const auto* z = reinterpret_cast<const std::complex<float>*>(data);
        for (int i = 0; i < nt; ++i) {
            const std::complex<float> s = interp1d(*kernel, z, n, 1, t[i]);
            out[2 * i] = s.real();
            out[2 * i + 1] = s.imag();
        }
        return 0;
    });
}

// geometry/rdr2geo_roots.cpp:14-27 ; returns 1 when converged like the reference
int tdbp_ref_rdr2geo_bracket(double t, double r, double fd, const I3B_Orbit* o,
                             const I3B_DEM* d, double wvl, int side,
                             const I3B_Rdr2GeoBracketParams* p, double* xyz)
{
    return guarded([&]() {
        const Orbit orbit = make_orbit(*o, DateTime(2000, 1, 1));
        const isce3::geometry::DEMInterpolator dem(*d);
        Vec3 x;
        const int ok = isce3::geometry::rdr2geo_bracket(t, r, fd, orbit, dem, x, wvl,
                side == I3B_LOOK_RIGHT ? LookSide::Right : LookSide::Left, p->tol_height,
                p->look_min, p->look_max);
        for (int i = 0; i < 3; ++i) xyz[i] = x[i];
        return ok;
    });
}

// geometry/geo2rdr_roots.cpp:16-25
int tdbp_ref_geo2rdr_bracket(const double* x, const I3B_Orbit* o, const I3B_LUT2d* l,
                             double wvl, int side, const I3B_Geo2RdrBracketParams* p,
                             double* t, double* r)
{
    return guarded([&]() {
        const Orbit orbit = make_orbit(*o, DateTime(2000, 1, 1));
        const LUT2d<double> dop(*l);
        std::optional<double> ts, te;
        if (p->has_time_start) ts = p->time_start;
        if (p->has_time_end) te = p->time_end;
        return isce3::geometry::geo2rdr_bracket(Vec3(x[0], x[1], x[2]), orbit, dop, *t, *r,
                wvl, side == I3B_LOOK_RIGHT ? LookSide::Right : LookSide::Left,
                p->tol_aztime, ts, te);
    });
}

// core/Ellipsoid.h:177-224 (WGS84)
void tdbp_ref_xyz_to_llh(const double* x, double* llh)
{
    const Ellipsoid e(EarthSemiMajorAxis, EarthEccentricitySquared);
    const Vec3 r = e.xyzToLonLat(Vec3(x[0], x[1], x[2]));
    for (int i = 0; i < 3; ++i) llh[i] = r[i];
}

void tdbp_ref_llh_to_xyz(const double* llh, double* x)
{
    const Ellipsoid e(EarthSemiMajorAxis, EarthEccentricitySquared);
    const Vec3 r = e.lonLatToXyz(Vec3(llh[0], llh[1], llh[2]));
    for (int i = 0; i < 3; ++i) x[i] = r[i];
}

// focus/DryTroposphereModel.icc:10-29
double tdbp_ref_dry_tropo_tsx(const double* p, const double* llh)
{
    const Ellipsoid e(EarthSemiMajorAxis, EarthEccentricitySquared);
    return isce3::focus::dryTropoDelayTSX(Vec3(p[0], p[1], p[2]),
                                          Vec3(llh[0], llh[1], llh[2]), e);
}

// math/RootFind1dBracket.icc:57-216 on a caller-supplied function
int tdbp_ref_brent(double a, double b, double (*f)(double, void*), void* ctx, double tol,
                   double* root)
{
    const auto ec = isce3::math::find_zero_brent(
            a, b, [&](double x) { return f(x, ctx); }, tol, root);
    return static_cast<int>(ec);
}

double tdbp_ref_lut2d_eval(const I3B_LUT2d* l, double y, double x)
{
    return LUT2d<double>(*l).eval(y, x);
}

// core/Projections.cpp: createProj(epsg)->forward  (lon, lat in radians)
int tdbp_ref_project_forward(int epsg, double lon, double lat, double* xy)
{
    return guarded([&]() {
        std::unique_ptr<ProjectionBase> pj(createProj(epsg));
        Vec3 out;
        const int st = pj->forward(Vec3(lon, lat, 0.0), out);
        xy[0] = out[0];
        xy[1] = out[1];
        return st;
    });
}

double tdbp_ref_dem_interp(const I3B_DEM* d, double lon, double lat)
{
    return isce3::geometry::DEMInterpolator(*d).interpolateLonLat(lon, lat);
}

} // extern "C"
