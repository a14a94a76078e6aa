// TEST INFRASTRUCTURE -- CPU oracle ("port") for the TDBP parity tests; builds
// into oracle/libtdbp_oracle.so.  Not product code: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may build, load or call it.  The product (isce3_b200/csrc) never links it.
//
// A self-contained restatement, in plain C++ on flat arrays, of what
// isce3::focus::backproject computes (cxx/isce3/focus/Backproject.cpp:65-212).
// Every function cites the reference lines it follows.  PARITY PIN: this file
// is checked in tests/test_oracle.py against (1) the reference's own data-free
// known-answer tests and (2) oracle/_ref/libtdbp_ref.so, which
// is the reference's Backproject.cpp + geometry/orbit sources compiled
// unchanged; the reference holds no golden *output array* for backproject
// (its only fixture, tests/data/point-target-sim-rc.h5, is stripped from the
// mount and its test asserts IRF metrics only), so array parity is pinned on
// "reference code run here on identical synthetic inputs".
//
// Floating-point evaluation order follows the reference expression by
// expression (left-to-right dot products as Eigen's 3-vector redux, divisions
// where the reference divides) so that the two agree to the last few ulps.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../include/isce3_b200_backproject.h"
#include "tdbp_samplers.h"

namespace tdbp_oracle {

constexpr double kC = 299792458.0;           // core/Constants.h:50
constexpr double kA = 6378137.0;             // core/Constants.h:41
constexpr double kE2 = 0.006694379990141317; // core/Constants.h:44

struct V3 {
    double x, y, z;
};
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
static inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
static inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
static inline V3 unit(V3 a) { return a / norm(a); }
static inline V3 cross(V3 a, V3 b)
{
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline V3 ld3(const double* p, int i) { return {p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }
static const V3 kNaN3 = {std::numeric_limits<double>::quiet_NaN(),
                         std::numeric_limits<double>::quiet_NaN(),
                         std::numeric_limits<double>::quiet_NaN()};

// ---- orbit ---------------------------------------------------------------

// core/Linspace.icc:73-89 (search: index of first sample > val)
static inline int linspace_search(double first, double spacing, int size, double val)
{
    const double last = first + (size - 1) * spacing;
    if (spacing >= 0) {
        if (val < first) return 0;
        if (val > last) return size;
    } else {
        if (val > first) return 0;
        if (val < last) return size;
    }
    return (int) ((val - first) / spacing + 1);
}

static inline double otime(const I3B_Orbit& o, int i) { return o.t0 + i * o.dt; }

// core/detail/InterpolateOrbit.icc:15-109 (third-order Hermite on 4 vectors)
static void orbit_hermite(const I3B_Orbit& o, double t, V3* pos, V3* vel)
{
    int idx = linspace_search(o.t0, o.dt, o.n, t) - 2;
    idx = std::min(std::max(idx, 0), o.n - 4);
    double f1[4], f0[4], h[4], hdot[4], g1[4], g0[4];
    for (int i = 0; i < 4; ++i) f1[i] = t - otime(o, idx + i);
    for (int i = 0; i < 4; ++i) {
        double sum = 0.;
        for (int j = 0; j < 4; ++j) {
            if (j == i) continue;
            sum += 1. / (otime(o, idx + i) - otime(o, idx + j));
        }
        f0[i] = 1. - 2. * sum * (t - otime(o, idx + i));
    }
    for (int i = 0; i < 4; ++i) {
        h[i] = 1.;
        for (int j = 0; j < 4; ++j) {
            if (j == i) continue;
            h[i] *= (t - otime(o, idx + j)) / (otime(o, idx + i) - otime(o, idx + j));
        }
    }
    V3 p = {0, 0, 0};
    for (int i = 0; i < 4; ++i)
        p = p + (h[i] * h[i]) * (ld3(o.pos, idx + i) * f0[i] + ld3(o.vel, idx + i) * f1[i]);
    *pos = p;
    for (int i = 0; i < 4; ++i) {
        hdot[i] = 0.;
        for (int j = 0; j < 4; ++j) {
            if (j == i) continue;
            double prod = 1. / (otime(o, idx + i) - otime(o, idx + j));
            for (int k = 0; k < 4; ++k) {
                if (k == i || k == j) continue;
                prod *= (t - otime(o, idx + k)) / (otime(o, idx + i) - otime(o, idx + k));
            }
            hdot[i] += prod;
        }
    }
    for (int i = 0; i < 4; ++i) g1[i] = h[i] + 2. * hdot[i] * (t - otime(o, idx + i));
    for (int i = 0; i < 4; ++i) {
        double sum = 0.;
        for (int j = 0; j < 4; ++j) {
            if (j == i) continue;
            sum += 1. / (otime(o, idx + i) - otime(o, idx + j));
        }
        g0[i] = 2. * (f0[i] * hdot[i] - sum * h[i]);
    }
    V3 v = {0, 0, 0};
    for (int i = 0; i < 4; ++i)
        v = v + h[i] * (ld3(o.pos, idx + i) * g0[i] + ld3(o.vel, idx + i) * g1[i]);
    *vel = v;
}

// core/detail/InterpolateOrbit.icc:116-155 (eighth-order Legendre on 9 vectors)
static void orbit_legendre(const I3B_Orbit& o, double t, V3* pos, V3* vel)
{
    int idx = linspace_search(o.t0, o.dt, o.n, t) - 5;
    idx = std::min(std::max(idx, 0), o.n - 9);
    const double trel = 8. * (t - otime(o, idx)) / (otime(o, idx + 8) - otime(o, idx));
    double teller = 1.;
    for (int i = 0; i < 9; ++i) teller *= trel - i;
    if (teller == 0.) {
        const int i = (int) trel;
        *pos = ld3(o.pos, idx + i);
        *vel = ld3(o.vel, idx + i);
        return;
    }
    static const double noemer[9] = {40320.0, -5040.0, 1440.0, -720.0, 576.0,
                                     -720.0,  1440.0,  -5040.0, 40320.0};
    V3 p = {0, 0, 0}, v = {0, 0, 0};
    for (int i = 0; i < 9; ++i) {
        const double coeff = (teller / noemer[i]) / (trel - i);
        p = p + coeff * ld3(o.pos, idx + i);
        v = v + coeff * ld3(o.vel, idx + i);
    }
    *pos = p;
    *vel = v;
}

enum { BORDER_ERROR = 0, BORDER_EXTRAPOLATE = 1, BORDER_FILLNAN = 2 };

// core/detail/InterpolateOrbit.icc:161-193
static int orbit_interpolate(const I3B_Orbit& o, double t, int border, V3* pos, V3* vel)
{
    const int need = o.method == I3B_ORBIT_LEGENDRE ? 9 : 4; // core/Orbit.h minStateVecs
    if (o.n < need) return I3B_ORBIT_INTERP_SIZE_ERROR;
    const double tstart = otime(o, 0), tend = otime(o, o.n - 1);
    if (t < tstart || t > tend) {
        if (border == BORDER_FILLNAN) {
            *pos = kNaN3;
            *vel = kNaN3;
        }
        if (border != BORDER_EXTRAPOLATE) return I3B_ORBIT_INTERP_DOMAIN_ERROR;
    }
    switch (o.method) {
    case I3B_ORBIT_HERMITE: orbit_hermite(o, t, pos, vel); return I3B_SUCCESS;
    case I3B_ORBIT_LEGENDRE: orbit_legendre(o, t, pos, vel); return I3B_SUCCESS;
    default: return I3B_ORBIT_INTERP_UNKNOWN_METHOD;
    }
}

// ---- ellipsoid (WGS84) ---------------------------------------------------

// core/Ellipsoid.h:99-102 (rEast)
static inline double r_east(double lat)
{
    return kA / std::sqrt(1.0 - (kE2 * std::pow(std::sin(lat), 2)));
}

// core/Ellipsoid.h:177-190
static V3 llh_to_xyz(V3 llh)
{
    const double re = r_east(llh.y);
    V3 r;
    r.x = (re + llh.z) * std::cos(llh.y) * std::cos(llh.x);
    r.y = (re + llh.z) * std::cos(llh.y) * std::sin(llh.x);
    r.z = ((re * (1.0 - kE2)) + llh.z) * std::sin(llh.y);
    return r;
}

// core/Ellipsoid.h:196-224 (Vermeille 2002)
static V3 xyz_to_llh(V3 p3)
{
    const double e4 = kE2 * kE2;
    const double a2 = kA * kA;
    const double p = (std::pow(p3.x, 2) + std::pow(p3.y, 2)) / a2;
    const double q = ((1. - kE2) * std::pow(p3.z, 2)) / a2;
    const double r = (p + q - e4) / 6.;
    const double s = (e4 * p * q) / (4. * std::pow(r, 3));
    const double t = std::pow(1. + s + std::sqrt(s * (2. + s)), (1. / 3.));
    const double u = r * (1. + t + (1. / t));
    const double rv = std::sqrt(std::pow(u, 2) + (e4 * q));
    const double w = (kE2 * (u + rv - q)) / (2. * rv);
    const double k = std::sqrt(u + rv + std::pow(w, 2)) - w;
    const double d = (k * std::sqrt(std::pow(p3.x, 2) + std::pow(p3.y, 2))) / (k + kE2);
    V3 llh;
    llh.y = std::atan2(p3.z, d);
    llh.x = std::atan2(p3.y, p3.x);
    llh.z = ((k + kE2 - 1.) * std::sqrt(std::pow(d, 2) + std::pow(p3.z, 2))) / k;
    return llh;
}

// core/Ellipsoid.h:152-158
static inline V3 n_vector(double lon, double lat)
{
    const double clat = std::cos(lat);
    return {clat * std::cos(lon), clat * std::sin(lon), std::sin(lat)};
}

// ---- Brent ----------------------------------------------------------------

static inline bool opposite_sign(double a, double b) { return std::signbit(a) ^ std::signbit(b); }

// math/RootFind1dBracket.icc:57-216
template<class F>
static int brent(double a, double b, F f, const double tol, double* root)
{
    constexpr double eps = std::numeric_limits<double>::epsilon();
    if (tol < 0.0) return I3B_INVALID_TOLERANCE;
    double c, d, e, fa, fb, fc, p, q, r, s, tol1;
    fa = f(a);
    if (fa == 0.0) {
        *root = a;
        return I3B_SUCCESS;
    }
    fb = f(b);
    if (fb == 0.0) {
        *root = b;
        return I3B_SUCCESS;
    }
    if (!opposite_sign(fa, fb)) return I3B_INVALID_INTERVAL;
    c = a;
    fc = fa;
    e = d = b - a;
    tol1 = tol > 0.0 ? tol : eps;
    const int maxiter = 3 * (int) std::ceil(std::log2(std::abs((a - b) / tol1)));
    for (int it = 0; it < maxiter; ++it) {
        if (std::abs(fc) < std::abs(fb)) {
            a = b; b = c; c = a;
            fa = fb; fb = fc; fc = fa;
        }
        tol1 = 2 * eps * std::abs(b) + 0.5 * tol;
        const double xm = 0.5 * (c - b);
        if ((std::abs(xm) <= tol1) || (fb == 0.0)) {
            *root = b;
            return I3B_SUCCESS;
        }
        if ((std::abs(e) < tol1) || (std::abs(fa) <= std::abs(fb))) {
            e = d = xm;
        } else {
            s = fb / fa;
            if (a == c) {
                p = 2 * xm * s;
                q = 1.0 - s;
            } else {
                q = fa / fc;
                r = fb / fc;
                p = s * (2 * xm * q * (q - r) - (b - a) * (r - 1.0));
                q = (q - 1.0) * (r - 1.0) * (s - 1.0);
            }
            if (p > 0.0) q = -q; else p = -p;
            s = e;
            e = d;
            if (((2 * p) >= (3 * xm * q - std::abs(tol1 * q))) ||
                (p >= std::abs(0.5 * s * q))) {
                e = d = xm;
            } else {
                d = p / q;
            }
        }
        a = b;
        fa = fb;
        if (std::abs(d) <= tol1) {
            b = (xm <= 0.0) ? b - tol1 : b + tol1;
        } else {
            b = b + d;
        }
        fb = f(b);
        if (!opposite_sign(fb, fc)) {
            c = a;
            fc = fa;
            e = d = b - a;
        }
    }
    *root = b;
    return I3B_FAILED_TO_CONVERGE;
}

// ---- geometry ---------------------------------------------------------------

// geometry/detail/Rdr2Geo.icc:175-242 via geometry/rdr2geo_roots.cpp:14-27.
// The orbit is sampled with border mode Error: the reference throws OutOfRange
// (core/Orbit.cpp:78-83); here that surfaces as I3B_EXC_OUT_OF_RANGE.
static int rdr2geo_bracket(double aztime, double slant_range, double doppler,
                           const I3B_Orbit& orbit, const I3B_DEM& dem, double wavelength,
                           int side, const I3B_Rdr2GeoBracketParams& prm, V3* xyz)
{
    V3 radar, velocity;
    const int st = orbit_interpolate(orbit, aztime, BORDER_ERROR, &radar, &velocity);
    if (st != I3B_SUCCESS) return I3B_EXC_OUT_OF_RANGE;
    const double speed = norm(velocity);
    const V3 along = velocity / speed;
    const V3 right = unit(cross(along, radar));
    const V3 down = cross(along, right);
    const V3 horizontal = (side == I3B_LOOK_RIGHT) ? right : -1.0 * right;
    const double sin_squint = doppler * wavelength / (2 * speed);
    const double cos_squint = std::sqrt(1.0 - sin_squint * sin_squint);
    const V3 center = radar + (sin_squint * slant_range) * along;
    const double radius = cos_squint * slant_range;
    auto get_xyz = [&](double look) {
        return center + (radius * std::sin(look)) * horizontal +
               (radius * std::cos(look)) * down;
    };
    auto dh = [&](double look) {
        const V3 llh = xyz_to_llh(get_xyz(look));
        return llh.z - dem_interp_lonlat(dem, llh.x, llh.y);
    };
    const double tol_look = prm.tol_height / radius;
    double look = 0.0;
    const int err = brent(prm.look_min, prm.look_max, dh, tol_look, &look);
    if (err != I3B_SUCCESS) return err;
    *xyz = get_xyz(look);
    return I3B_SUCCESS;
}

// geometry/detail/Geo2Rdr.icc:185-238 via geometry/geo2rdr_roots.cpp:16-25
static int geo2rdr_bracket(V3 x, const I3B_Orbit& orbit, const I3B_LUT2d& dop,
                           double wavelength, int side, const I3B_Geo2RdrBracketParams& prm,
                           double* aztime, double* range)
{
    const double orbit_start = otime(orbit, 0), orbit_end = otime(orbit, orbit.n - 1);
    double t0, t1;
    if (prm.has_time_start) t0 = prm.time_start;
    else t0 = dop.have_data ? std::max(orbit_start, dop.ystart) : orbit_start;
    if (prm.has_time_end) t1 = prm.time_end;
    else {
        if (dop.have_data) {
            const double dop_end = dop.ystart + dop.dy * (dop.length - 1);
            t1 = std::min(orbit_end, dop_end);
        } else t1 = orbit_end;
    }
    V3 xp, v, r;
    auto doppler_error = [&](double t) {
        orbit_interpolate(orbit, t, BORDER_FILLNAN, &xp, &v);
        r = x - xp;
        const double rnorm = norm(r);
        const double fd = lut2d_eval(dop, t, rnorm);
        return 2.0 / wavelength * dot(v, r) / rnorm - fd;
    };
    const int err = brent(t0, t1, doppler_error, prm.tol_aztime, aztime);
    if (err != I3B_SUCCESS) return err;
    orbit_interpolate(orbit, *aztime, BORDER_FILLNAN, &xp, &v);
    r = x - xp;
    *range = norm(r);
    if ((side == I3B_LOOK_RIGHT) ^ (dot(cross(r, v), xp) > 0)) return I3B_WRONG_LOOK_SIDE;
    return I3B_SUCCESS;
}

// focus/BistaticDelay.icc:10-17
static inline double bistatic_delay(V3 p, V3 v, V3 x)
{
    const V3 r = x - p;
    return 2. * (dot(r, v) - kC * norm(r)) / (dot(v, v) - (kC * kC));
}

// focus/DryTroposphereModel.icc:10-29
static double dry_tropo_tsx(V3 p, V3 llh)
{
    constexpr double ZPD = 2.3, H = 6000.;
    const V3 x = llh_to_xyz(llh);
    const V3 r_hat = unit(p - x);
    const V3 n_hat = unit(n_vector(llh.x, llh.y));
    const double cos_theta = dot(r_hat, n_hat);
    return 2. * ZPD * std::exp(-llh.z / H) / (kC * cos_theta);
}

// ---- interpolation kernels (Kernel<float>) --------------------------------

// math/Sinc.icc:69-91
template<typename T>
static inline T sinc(T t)
{
    const T eps1 = std::sqrt(std::numeric_limits<T>::epsilon());
    const T eps2 = std::sqrt(eps1);
    const T x = T(M_PI) * std::abs(t);
    if (x < eps2) {
        T out = 1;
        if (x > eps1) out -= x * x / T(6);
        return out;
    }
    return std::sin(x) / x;
}

// core/Kernels.icc:29-52 (Knab 1983 sampling window x sinc), evaluated in T
template<typename T>
static inline T knab(double t, double halfwidth, double bandwidth)
{
    const T st = sinc<T>(T(t));
    const T hw = T(halfwidth), bw = T(bandwidth), tt = T(t);
    const T c = M_PI * hw * (1.0 - bw);
    const T tf = tt / hw;
    const std::complex<T> y = std::sqrt((std::complex<T>) (1.0 - tf * tf));
    const T window = std::real(std::cosh(c * y) / std::cosh(c));
    return window * st;
}

struct KernelEval {
    I3B_Kernel k;
    double halfwidth;
    int width_taps; // ceil(width)
    float one_dx;   // TabulatedKernel::_1_dx stored as T=float (Kernels.h:145)
    int imax;
    float cheb_scale;

    explicit KernelEval(const I3B_Kernel& kk) : k(kk)
    {
        if (k.kind == I3B_KERNEL_LINEAR) k.width = 2.0; // Kernels.h:51
        halfwidth = std::fabs(k.width / 2.0);            // Kernels.h:27
        width_taps = (int) std::ceil(halfwidth * 2);
        one_dx = 0.f;
        imax = 0;
        cheb_scale = 0.f;
        if (k.kind == I3B_KERNEL_TABULATED) { // Kernels.icc:114-137
            imax = k.n - 2;
            const double dx = halfwidth / (k.n - 1.0);
            one_dx = (float) (1.0 / dx);
        } else if (k.kind == I3B_KERNEL_CHEBY) { // Kernels.icc:171
            cheb_scale = (float) (4.0 / k.width);
        }
    }

    float operator()(double t) const
    {
        switch (k.kind) {
        case I3B_KERNEL_BARTLETT:
        case I3B_KERNEL_LINEAR: { // Kernels.icc:15-23
            const double t2 = std::fabs(t / halfwidth);
            if (t2 > 1.0) return 0.f;
            return (float) (1.0 - t2);
        }
        case I3B_KERNEL_KNAB: return knab<float>(t, halfwidth, k.bandwidth);
        case I3B_KERNEL_TABULATED: { // Kernels.icc:139-154
            const double ax = std::abs(t);
            if (ax > halfwidth) return 0.f;
            const double axn = ax * one_dx;
            int i = (int) std::floor(axn);
            i = std::min(i, imax);
            return (float) (k.data[i] + (axn - i) * (k.data[i + 1] - k.data[i]));
        }
        case I3B_KERNEL_CHEBY: { // Kernels.icc:191-211 (Clenshaw, in float)
            const double ax = std::abs(t);
            if (ax > halfwidth) return 0.f;
            const float q = (float) ((ax * cheb_scale) - 1.f);
            const float twoq = 2.f * q;
            float bk = 0, bk1 = 0, bk2 = 0;
            for (int i = k.n - 1; i > 0; --i) {
                bk = k.data[i] + twoq * bk1 - bk2;
                bk2 = bk1;
                bk1 = bk;
            }
            return k.data[0] + q * bk1 - bk2;
        }
        default: return std::numeric_limits<float>::quiet_NaN();
        }
    }
};

// core/Interp1d.icc:7-22 with detail/Interp1d.h:22-38 (coeffs), :54-80
// (zero-padded window), :90-107 (inner product in complex<float>, index order)
static std::complex<float> interp1d(const KernelEval& kern, const std::complex<float>* x,
                                    size_t length, double t)
{
    const int width = kern.width_taps;
    float cbuf[64];
    std::complex<float> dbuf[64];
    std::vector<float> cheap;
    std::vector<std::complex<float>> dheap;
    float* coeffs = cbuf;
    std::complex<float>* block = dbuf;
    if (width > 64) {
        cheap.resize(width);
        dheap.resize(width);
        coeffs = cheap.data();
        block = dheap.data();
    }
    long i0 = (width % 2 == 0) ? (long) std::ceil(t) : (long) std::round(t);
    const long low = i0 - width / 2;
    for (int i = 0; i < width; ++i) coeffs[i] = kern((double) (i + low) - t);
    const long high = low + width;
    const std::complex<float>* px;
    if (low >= 0 && high < (long) length) {
        px = &x[low];
    } else {
        for (int i = 0; i < width; ++i) {
            const long j = low + i;
            block[i] = (j >= 0 && j < (long) length) ? x[j] : std::complex<float>(0);
        }
        px = block;
    }
    std::complex<float> sum = 0;
    for (int i = 0; i < width; ++i) sum += coeffs[i] * px[i];
    return sum;
}

// focus/Backproject.cpp:30-63
static std::complex<float> sum_coherent(const std::complex<float>* data, double swst,
                                        double dtau, int nr, const V3* pos, const V3* vel,
                                        V3 x, double fc, double tau_atm,
                                        const KernelEval& kernel, int kstart, int kstop)
{
    std::complex<double> sum(0., 0.);
    for (int k = kstart; k < kstop; ++k) {
        const double tau = tau_atm + bistatic_delay(pos[k], vel[k], x);
        const std::complex<float>* line = &data[size_t(k) * nr];
        const double u = (tau - swst) / dtau;
        std::complex<double> s = interp1d(kernel, line, nr, u);
        const double phi = 2. * M_PI * fc * tau;
        s *= std::complex<double>(std::cos(phi), std::sin(phi));
        sum += s;
    }
    return std::complex<float>(sum);
}

static thread_local std::string g_err;

// focus/Backproject.cpp:65-212.  `pp_out` (optional) receives
// sum(kstop-kstart), the work unit of the bench metric.
static int backproject(const I3B_BackprojectArgs& a, double* pp_out)
{
    const float nan = std::numeric_limits<float>::quiet_NaN();
    if (!(a.dry_tropo_model == I3B_TROPO_NODELAY || a.dry_tropo_model == I3B_TROPO_TSX)) {
        g_err = "unexpected dry troposphere model"; // :78-83
        return I3B_EXC_INVALID_ARGUMENT;
    }
    const I3B_RadarGeometry &og = a.out_geometry, &ig = a.in_geometry;
    if (og.ref_epoch_sec != ig.ref_epoch_sec || og.ref_epoch_frac != ig.ref_epoch_frac) {
        g_err = "input reference epoch must match output reference epoch"; // :88-92
        return I3B_EXC_RUNTIME_ERROR;
    }
    // :95-98 (container/RadarGeometry.icc:28-58)
    const double in_t0 = ig.grid.sensing_start, in_dt = 1.0 / ig.grid.prf;
    const int in_lines = (int) ig.grid.length;
    const double out_t0 = og.grid.sensing_start, out_dt = 1.0 / og.grid.prf;
    const int out_lines = (int) og.grid.length, out_width = (int) og.grid.width;

    // :101-106 platform position & velocity at each pulse (border mode Error)
    std::vector<V3> pos(in_lines), vel(in_lines);
    // (ABI 3 extension, not in the reference: explicit pulse times; see the header)
    const double* T = a.pulse_times;
    for (int i = 0; i < in_lines; ++i) {
        const double t = T ? T[i] : in_t0 + i * in_dt;
        if (orbit_interpolate(ig.orbit, t, BORDER_ERROR, &pos[i], &vel[i]) != I3B_SUCCESS) {
            g_err = "orbit interpolation outside of orbit domain";
            return I3B_EXC_OUT_OF_RANGE;
        }
    }
    // :109-112
    const double swst = 2. * ig.grid.starting_range / kC;
    const double dtau = 2. * ig.grid.range_pixel_spacing / kC;
    const int nr = (int) ig.grid.width;
    const double wvl = kC / a.fc; // :119
    const KernelEval kernel(a.kernel);
    const auto* in = reinterpret_cast<const std::complex<float>*>(a.in);
    auto* out = reinterpret_cast<std::complex<float>*>(a.out);

    bool all_converged = true;
    int thrown = 0;
    double pp = 0.0;
#pragma omp parallel for collapse(2) reduction(+ : pp)
    for (int j = 0; j < out_lines; ++j) {
        for (int i = 0; i < out_width; ++i) {
            const size_t pix = size_t(j) * out_width + i;
            V3 x, llh;
            {
                const double t = out_t0 + j * out_dt;
                const double r = og.grid.starting_range + i * og.grid.range_pixel_spacing;
                const double fD = lut2d_eval(og.doppler, t, r); // :134
                const int st = rdr2geo_bracket(t, r, fD, og.orbit, a.dem, wvl,
                                               og.grid.look_side, a.rdr2geo, &x); // :136-139
                if (st == I3B_EXC_OUT_OF_RANGE) {
                    thrown = st;
                    continue;
                }
                const bool converged = (st == I3B_SUCCESS);
                // the reference converts x unconditionally (:141); when the solver
                // failed x is indeterminate there, so skip the conversion
                llh = converged ? xyz_to_llh(x) : kNaN3;
                if (a.height) a.height[pix] = (float) llh.z; // :143-145
                if (!converged) {                            // :146-153
                    all_converged = false;
                    out[pix] = {nan, nan};
                    if (a.height) a.height[pix] = nan;
                    continue;
                }
            }
            double t, r;
            {
                const int st = geo2rdr_bracket(x, ig.orbit, ig.doppler, wvl,
                                               ig.grid.look_side, a.geo2rdr, &t, &r); // :161-165
                if (st != I3B_SUCCESS) { // :167-171
                    all_converged = false;
                    out[pix] = {nan, nan};
                    continue;
                }
            }
            V3 p, v; // :175-176 (border mode Error; t is inside the orbit here)
            orbit_interpolate(ig.orbit, t, BORDER_ERROR, &p, &v);
            const double l = wvl * r * (norm(p) / norm(x)) / (2. * a.ds); // :180
            const double cpi = l / norm(v);                               // :183
            const double tstart = t - 0.5 * cpi, tstop = t + 0.5 * cpi;
            int kstart = (int) std::floor((tstart - in_t0) / in_dt); // :190-193
            int kstop = (int) std::ceil((tstop - in_t0) / in_dt);
            if (T) {
                // last pulse at or before tstart / first pulse at or after tstop: what floor and
                // ceil above select on a uniform grid
                kstart = (int) (std::upper_bound(T, T + in_lines, tstart) - T) - 1;
                kstop = (int) (std::lower_bound(T, T + in_lines, tstop) - T);
            }
            kstart = std::max(kstart, 0);
            kstop = std::min(kstop, in_lines);
            double tau_atm = 0.; // :196-199
            if (a.dry_tropo_model == I3B_TROPO_TSX) tau_atm = dry_tropo_tsx(p, llh);
            out[pix] = sum_coherent(in, swst, dtau, nr, pos.data(), vel.data(), x, a.fc,
                                    tau_atm, kernel, kstart, kstop); // :202-204
            if (kstop > kstart) pp += double(kstop - kstart);
        }
    }
    if (pp_out) *pp_out = pp;
    if (thrown) {
        g_err = "orbit interpolation outside of orbit domain";
        return thrown;
    }
    return all_converged ? I3B_SUCCESS : I3B_FAILED_TO_CONVERGE; // :208-211
}

} // namespace tdbp_oracle

using namespace tdbp_oracle;

extern "C" {

const char* tdbp_oracle_last_error() { return g_err.c_str(); }
const char* tdbp_oracle_kind() { return "port"; }

int tdbp_oracle_backproject(const I3B_BackprojectArgs* a) { return backproject(*a, nullptr); }

int tdbp_oracle_backproject_pp(const I3B_BackprojectArgs* a, double* pixel_pulses)
{
    return backproject(*a, pixel_pulses);
}

double tdbp_oracle_bistatic_delay(const double* p, const double* v, const double* x)
{
    return bistatic_delay(ld3(p, 0), ld3(v, 0), ld3(x, 0));
}

int tdbp_oracle_orbit_interpolate(const I3B_Orbit* o, double t, int border_mode, double* pos,
                                  double* vel)
{
    V3 p = {0, 0, 0}, v = {0, 0, 0};
    const int st = orbit_interpolate(*o, t, border_mode, &p, &v);
    pos[0] = p.x; pos[1] = p.y; pos[2] = p.z;
    vel[0] = v.x; vel[1] = v.y; vel[2] = v.z;
    if (st != I3B_SUCCESS && border_mode == BORDER_ERROR) return I3B_EXC_OUT_OF_RANGE;
    return st;
}

int tdbp_oracle_kernel_eval(const I3B_Kernel* k, const double* t, int n, float* out)
{
    const KernelEval ke(*k);
    for (int i = 0; i < n; ++i) out[i] = ke(t[i]);
    return 0;
}

// TabulatedKernel<float>(KnabKernel<double>(width, bandwidth), n): Kernels.icc:114-137
int tdbp_oracle_tabulate_knab(double width, double bandwidth, int n, float* table)
{
    const double hw = std::fabs(width / 2.0);
    const double dx = hw / (n - 1.0);
    for (int i = 0; i < n; ++i) table[i] = (float) knab<double>(i * dx, hw, bandwidth);
    return 0;
}

// ChebyKernel<float>(KnabKernel<double>(width, bandwidth), n): Kernels.icc:156-189
int tdbp_oracle_cheby_knab(double width, double bandwidth, int n, float* coeffs)
{
    const double hw = std::fabs(width / 2.0);
    std::vector<float> q(n), fx(n);
    const float scale = (float) (4.0 / width);
    for (int i = 0; i < n; ++i) {
        q[i] = (float) (M_PI * (2.0 * i + 1.0) / (2.0 * n));
        const float x = (float) ((std::cos(q[i]) + 1.0) / scale);
        fx[i] = (float) knab<double>(x, hw, bandwidth);
    }
    for (int i = 0; i < n; ++i) {
        coeffs[i] = 0.0f;
        for (int j = 0; j < n; ++j) {
            const float w = std::cos(i * q[j]);
            coeffs[i] += w * fx[j];
        }
        coeffs[i] *= 2.0 / n;
    }
    coeffs[0] *= 0.5;
    return 0;
}

int tdbp_oracle_interp1d(const I3B_Kernel* k, const float* data, size_t n, const double* t,
                         int nt, float* out)
{
    const KernelEval ke(*k);
    const auto* z = reinterpret_cast<const std::complex<float>*>(data);
    for (int i = 0; i < nt; ++i) {
        const std::complex<float> s = interp1d(ke, z, n, t[i]);
        out[2 * i] = s.real();
        out[2 * i + 1] = s.imag();
    }
    return 0;
}

// returns 1 when converged, like geometry/rdr2geo_roots.cpp:26
int tdbp_oracle_rdr2geo_bracket(double t, double r, double fd, const I3B_Orbit* o,
                                const I3B_DEM* d, double wvl, int side,
                                const I3B_Rdr2GeoBracketParams* p, double* xyz)
{
    V3 x = {0, 0, 0};
    const int st = rdr2geo_bracket(t, r, fd, *o, *d, wvl, side, *p, &x);
    xyz[0] = x.x; xyz[1] = x.y; xyz[2] = x.z;
    if (st < 0) return st;
    return st == I3B_SUCCESS;
}

int tdbp_oracle_geo2rdr_bracket(const double* x, const I3B_Orbit* o, const I3B_LUT2d* l,
                                double wvl, int side, const I3B_Geo2RdrBracketParams* p,
                                double* t, double* r)
{
    return geo2rdr_bracket(ld3(x, 0), *o, *l, wvl, side, *p, t, r) == I3B_SUCCESS;
}

void tdbp_oracle_xyz_to_llh(const double* x, double* llh)
{
    const V3 r = xyz_to_llh(ld3(x, 0));
    llh[0] = r.x; llh[1] = r.y; llh[2] = r.z;
}

void tdbp_oracle_llh_to_xyz(const double* llh, double* x)
{
    const V3 r = llh_to_xyz(ld3(llh, 0));
    x[0] = r.x; x[1] = r.y; x[2] = r.z;
}

double tdbp_oracle_dry_tropo_tsx(const double* p, const double* llh)
{
    return dry_tropo_tsx(ld3(p, 0), ld3(llh, 0));
}

int tdbp_oracle_brent(double a, double b, double (*f)(double, void*), void* ctx, double tol,
                      double* root)
{
    return brent(a, b, [&](double x) { return f(x, ctx); }, tol, root);
}

double tdbp_oracle_lut2d_eval(const I3B_LUT2d* l, double y, double x)
{
    return lut2d_eval(*l, y, x);
}

int tdbp_oracle_project_forward(int epsg, double lon, double lat, double* xy)
{
    return proj::forward(epsg, lon, lat, &xy[0], &xy[1]);
}

double tdbp_oracle_dem_interp(const I3B_DEM* d, double lon, double lat)
{
    return dem_interp_lonlat(*d, lon, lat);
}

} // extern "C"
