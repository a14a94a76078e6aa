// TEST INFRASTRUCTURE -- CPU oracle for the TDBP parity tests.  Not product
// code: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may build or call anything under oracle/.
//
// Restated 2-D samplers used by both oracle builds (the "port" in
// tdbp_oracle.cpp and the shim classes the reference sources are compiled
// against in oracle/_ref).  The reference implementations need
// Eigen/GDAL/pyre and cannot be compiled here, so these follow them
// statement by statement:
//
//   bilinear   cxx/isce3/core/BilinearInterpolator.cpp:13-48
//   bicubic    cxx/isce3/core/BicubicInterpolator.cpp:35-62
//   biquintic  cxx/isce3/core/Spline2dInterpolator.cpp:32-116 (order 6)
//   nearest    cxx/isce3/core/NearestNeighborInterpolator.cpp:13-23
//   sinc       cxx/isce3/core/Sinc2dInterpolator.cpp:13-133 (8 taps, 8192 sub-sample rows)
//   LUT2d      cxx/isce3/core/LUT2d.cpp:127-160, LUT2d.h:84-95
//   DEM        cxx/isce3/geometry/DEMInterpolator.cpp:592-659,
//              cxx/isce3/core/Projections.h:127-133 (LonLat::forward)
//
// Arithmetic type U matches the reference instantiation: float for the DEM
// raster (Matrix<float>), double for LUT2d<double>.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "../include/isce3_b200_backproject.h"

namespace tdbp_oracle {

template<typename U>
struct Grid2d {
    const U* data;
    long rows, cols;
    U operator()(long r, long c) const { return data[r * cols + c]; }
};

// BilinearInterpolator.cpp:13-48
template<typename U>
inline U bilinear(double x, double y, const Grid2d<U>& z)
{
    const int x1 = (int) std::floor(x), x2 = (int) std::ceil(x);
    const int y1 = (int) std::floor(y), y2 = (int) std::ceil(y);
    const U q11 = z(y1, x1), q12 = z(y2, x1), q21 = z(y1, x2), q22 = z(y2, x2);
    if (y1 == y2 && x1 == x2) return q11;
    if (y1 == y2)
        return U((x2 - x) / (x2 - x1)) * q11 + U((x - x1) / (x2 - x1)) * q21;
    if (x1 == x2)
        return U((y2 - y) / (y2 - y1)) * q11 + U((y - y1) / (y2 - y1)) * q12;
    const U den = U((x2 - x1) * (y2 - y1));
    return (q11 * U((x2 - x) * (y2 - y))) / den +
           (q21 * U((x - x1) * (y2 - y))) / den +
           (q12 * U((x2 - x) * (y - y1))) / den +
           (q22 * U((x - x1) * (y - y1))) / den;
}

// BicubicInterpolator.cpp:29-33 (uniform Catmull-Rom between p1 and p2)
template<typename U>
inline U catmull_rom(U p0, U p1, U p2, U p3, double tf)
{
    const double tc = 1. - tf;
    return (U(tf) * (p2 - p0 * U(tc * tc) +
                     (p2 * U(tc * 3. + 1.) - p3 * U(tc)) * U(tf)) +
            p1 * U(tf * tf * (tf * 3. - 5.) + 2.)) / U(2.);
}

// BicubicInterpolator.cpp:38-57
template<typename U>
inline U bicubic(double x, double y, const Grid2d<U>& z)
{
    const int x0 = (int) std::floor(x), y0 = (int) std::floor(y);
    U rowv[4];
    for (int i = -1; i < 3; ++i)
        rowv[i + 1] = catmull_rom<U>(z(y0 + i, x0 - 1), z(y0 + i, x0),
                                     z(y0 + i, x0 + 1), z(y0 + i, x0 + 2), x - x0);
    return catmull_rom<U>(rowv[0], rowv[1], rowv[2], rowv[3], y - y0);
}

// Spline2dInterpolator.cpp:95-116 (_initSpline) and :74-93 (_spline)
template<typename U, int N>
inline void spline_init(const U* Y, U* R, U* Q)
{
    Q[0] = U(0);
    R[0] = U(0);
    for (int i = 1; i < N - 1; ++i) {
        const U p = U(1.0) / (U(0.5) * Q[i - 1] + U(2.0));
        Q[i] = U(-0.5) * p;
        R[i] = (U(3.0) * (Y[i + 1] - U(2.0) * Y[i] + Y[i - 1]) - U(0.5) * R[i - 1]) * p;
    }
    R[N - 1] = U(0);
    for (int i = N - 2; i > 0; --i) R[i] = Q[i] * R[i + 1] + R[i];
}

template<typename U, int N>
inline U spline_eval(double x, const U* Y, const U* R)
{
    const U denom = U(6.0);
    if (x < 1.0) return Y[0] + U(x - 1.0) * (Y[1] - Y[0] - (R[1] / denom));
    if (x > N) return Y[N - 1] + U(x - N) * (Y[N - 1] - Y[N - 2] + (R[N - 2] / denom));
    const int j = (int) std::floor(x);
    const U xx = U(x - j);
    const U t0 = Y[j] - Y[j - 1] - (R[j - 1] / U(3.0)) - (R[j] / denom);
    const U t1 = xx * ((R[j - 1] / U(2.0)) + (xx * ((R[j] - R[j - 1]) / denom)));
    return Y[j - 1] + (xx * (t0 + t1));
}

// Spline2dInterpolator.cpp:32-72 with _order = 6 (createInterpolator default,
// core/Interpolator.h:205-214; LUT2d.cpp:181)
template<typename U>
inline U biquintic(double x, double y, const Grid2d<U>& z)
{
    constexpr int N = 6;
    const int nx = (int) z.cols, ny = (int) z.rows;
    int i0 = (int) y, j0 = (int) x; // even order: plain truncation
    i0 = i0 - (N / 2) + 1;
    j0 = j0 - (N / 2) + 1;
    U A[N], R[N], Q[N], HC[N];
    for (int i = 0; i < N; ++i) {
        const int indi = std::min(std::max(i0 + i, 0), ny - 2);
        for (int j = 0; j < N; ++j) {
            const int indj = std::min(std::max(j0 + j, 0), nx - 2);
            A[j] = z(indi + 1, indj + 1);
        }
        spline_init<U, N>(A, R, Q);
        HC[i] = spline_eval<U, N>(x - j0, A, R);
    }
    spline_init<U, N>(HC, R, Q);
    return spline_eval<U, N>(y - i0, HC, R);
}

// NearestNeighborInterpolator.cpp:13-23
template<typename U>
inline U nearest(double x, double y, const Grid2d<U>& z)
{
    return z((long) std::round(y), (long) std::round(x));
}

// Sinc2dInterpolator.cpp: constructor :13-40 with _sinc_coef :116-133 (beta 1, pedestal 0,
// weighted), as LUT2d / DEMInterpolator create it: createInterpolator(SINC_METHOD, 6, SINC_LEN = 8,
// SINC_SUB = 8192) (LUT2d.cpp:184-188, Interpolator.h:205-217, Constants.h:32-35)
inline const std::vector<double>& sinc_table()
{
    static const std::vector<double> table = [] {
        const int len = 8, sub = 8192, n = len * sub;
        std::vector<double> filter(n), t(n);
        const double wgthgt = 0.5, soff = (n - 1.) / 2.;
        for (int i = 0; i < n; ++i) {
            const double wgt = (1. - wgthgt) + (wgthgt * std::cos((M_PI * (i - soff)) / soff));
            const double s = std::floor(i - soff) / (1. * sub);
            const double fct = (s != 0.) ? (std::sin(M_PI * s) / (M_PI * s)) : 1.;
            filter[i] = fct * wgt;
        }
        for (int i = 0; i < sub; ++i) {
            double ssum = 0.0;
            for (int j = 0; j < len; ++j) ssum += filter[i + sub * j];
            for (int j = 0; j < len; ++j) t[(size_t) i * len + j] = filter[i + sub * j] / ssum;
        }
        return t;
    }();
    return table;
}

// Sinc2dInterpolator.cpp:45-101 (interp_impl, _sinc_eval_2d)
template<typename U>
inline U sinc2d(double x, double y, const Grid2d<U>& z)
{
    const int len = 8, half = 4, sub = 8192;
    const int ix = (int) std::floor(x), iy = (int) std::floor(y);
    const double fx = x - ix, fy = y - iy;
    if (ix < half - 1 || ix > z.cols - half - 1) return U(0);
    if (iy < half - 1 || iy > z.rows - half - 1) return U(0);
    const int xx = ix + half, yy = iy + half;
    const int ifx = std::min(std::max(0, int(fx * sub)), sub - 1);
    const int ify = std::min(std::max(0, int(fy * sub)), sub - 1);
    const double* k = sinc_table().data();
    U ret(0);
    for (int i = 0; i < len; ++i)
        for (int j = 0; j < len; ++j)
            ret += z(yy - i, xx - j) * static_cast<U>(k[(size_t) ify * len + i]) *
                   static_cast<U>(k[(size_t) ifx * len + j]);
    return ret;
}

#ifdef TDBP_USE_REFERENCE_INTERPOLATORS
} // namespace tdbp_oracle
#include <isce3/core/Interpolator.h>
namespace tdbp_oracle {
// oracle/_ref build: the 2-D interpolation itself is the REFERENCE's
// (core/{Bilinear,Bicubic,Spline2d,NearestNeighbor}Interpolator.cpp compiled unchanged; spline
// order 6 as createInterpolator's default, core/Interpolator.h:203-221).  The restated
// versions above are what the port oracle uses; tests compare the two builds.
template<typename U>
inline U interp2d(int method, double x, double y, const Grid2d<U>& z)
{
    using Map = Eigen::Map<const isce3::core::EArray2D<U>>;
    const Map m(z.data, z.rows, z.cols);
    static const isce3::core::BilinearInterpolator<U> bilinear_i;
    static const isce3::core::BicubicInterpolator<U> bicubic_i;
    static const isce3::core::Spline2dInterpolator<U> biquintic_i(6);
    static const isce3::core::NearestNeighborInterpolator<U> nearest_i;
    static const isce3::core::Sinc2dInterpolator<U> sinc_i(isce3::core::SINC_LEN, isce3::core::SINC_SUB);
    switch (method) {
    case I3B_INTERP_SINC: return sinc_i.interpolate(x, y, m);
    case I3B_INTERP_BICUBIC: return bicubic_i.interpolate(x, y, m);
    case I3B_INTERP_BIQUINTIC: return biquintic_i.interpolate(x, y, m);
    case I3B_INTERP_NEAREST: return nearest_i.interpolate(x, y, m);
    default: return bilinear_i.interpolate(x, y, m);
    }
}
#else
template<typename U>
inline U interp2d(int method, double x, double y, const Grid2d<U>& z)
{
    switch (method) {
    case I3B_INTERP_SINC: return sinc2d<U>(x, y, z);
    case I3B_INTERP_BICUBIC: return bicubic<U>(x, y, z);
    case I3B_INTERP_BIQUINTIC: return biquintic<U>(x, y, z);
    case I3B_INTERP_NEAREST: return nearest<U>(x, y, z);
    default: return bilinear<U>(x, y, z); // createInterpolator fallback
    }
}
#endif

// LUT2d.h:84-95
inline bool lut2d_contains(const I3B_LUT2d& l, double y, double x)
{
    if (!l.have_data) return true;
    const double i = (x - l.xstart) / l.dx;
    const double j = (y - l.ystart) / l.dy;
    return (i >= 0.0 && i <= l.width - 1.0) && (j >= 0.0 && j <= l.length - 1.0);
}

// LUT2d.cpp:127-160.  The reference logs through a pyre error channel when
// bounds_error is set and the point is outside, then clamps and evaluates;
// *out_of_bounds reports that condition to the caller instead.
inline double lut2d_eval(const I3B_LUT2d& l, double y, double x, bool* out_of_bounds = nullptr)
{
    if (!l.have_data) return l.ref_value;
    double xi = (x - l.xstart) / l.dx;
    double yi = (y - l.ystart) / l.dy;
    if (l.bounds_error && !lut2d_contains(l, y, x) && out_of_bounds) *out_of_bounds = true;
    xi = std::min(std::max(xi, 0.0), l.width - 1.0);
    yi = std::min(std::max(yi, 0.0), l.length - 1.0);
    const Grid2d<double> g {l.data, (long) l.length, (long) l.width};
    return interp2d<double>(l.method, xi, yi, g);
}

// DEMInterpolator.cpp:617-659
inline double dem_interp_xy(const I3B_DEM& d, double x, double y)
{
    if (!d.have_raster) return d.ref_height;
    if (d.epsg == 4326 && (x > 360 || x < -360)) x = std::fmod(x, 360);
    if (d.epsg == 4326 && x < -180) x += 360;
    if (d.epsg == 4326 && x - 360 >= d.xstart) {
        x -= 360;
    } else if (d.epsg == 4326 && x < d.xstart && x + 360 >= d.xstart) {
        x += 360;
    } else if (x < d.xstart) {
        return d.ref_height;
    }
    const double row = (y - d.ystart) / d.dy;
    const double col = (x - d.xstart) / d.dx;
    const int irow = (int) std::floor(row);
    const int icol = (int) std::floor(col);
    if (irow < 2 || irow >= (int) (d.length - 1)) return d.ref_height;
    if (icol < 2 || icol >= (int) (d.width - 1)) return d.ref_height;
    const Grid2d<float> g {d.data, (long) d.length, (long) d.width};
    return interp2d<float>(d.method, col, row, g);
}

// ---- map projections (forward only): restatement of cxx/isce3/core/Projections.cpp -------
// createProj :373-402, UTM ctor/forward :84-213, PolarStereo :247-297, CEA :324-360.
// The oracle/_ref build does NOT use these: its DEMInterpolator shim calls the reference's
// own createProj()->forward() (Projections.cpp compiled unchanged).
namespace proj {

// Clenshaw summation of sum_k a[k] sin((k+1) B)  (what Projections.cpp:36-46 computes)
inline double clens(const double* a, int size, double B)
{
    const double two_cos = 2.0 * std::cos(B);
    double u_next = 0.0, u = a[size - 1];
    for (int k = size - 2; k >= 0; --k) {
        const double t = two_cos * u - u_next + a[k];
        u_next = u;
        u = t;
    }
    return std::sin(B) * u;
}

// Complex-argument Clenshaw summation (Projections.cpp:58-82): real and imaginary part of
// sum_k a[k] sin((k+1)(B + i C))
inline double clenS(const double* a, int size, double B, double C, double& R, double& I)
{
    const double sB = std::sin(B), cB = std::cos(B), shC = std::sinh(C), chC = std::cosh(C);
    const double pr = 2.0 * cB * chC, pi = -2.0 * sB * shC; // 2 cos(B + iC)
    double ur = a[size - 1], ui = 0.0, ur_next = 0.0, ui_next = 0.0;
    for (int k = size - 2; k >= 0; --k) {
        const double tr = pr * ur - pi * ui - ur_next + a[k];
        const double ti = pi * ur + pr * ui - ui_next;
        ur_next = ur;
        ui_next = ui;
        ur = tr;
        ui = ti;
    }
    R = sB * chC * ur - cB * shC * ui;
    I = sB * chC * ui + cB * shC * ur;
    return R;
}

inline double pj_tsfn(double phi, double sinphi, double e)
{
    sinphi *= e;
    return std::tan(.5 * ((.5 * M_PI) - phi)) / std::pow((1. - sinphi) / (1. + sinphi), .5 * e);
}

inline double pj_qsfn(double sinphi, double e, double one_es)
{
    const double con = e * sinphi;
    return one_es * ((sinphi / (1. - std::pow(con, 2))) - ((.5 / e) * std::log((1. - con) / (1. + con))));
}

constexpr double kA = 6378137.0, kE2 = 0.006694379990141317;

// returns 0 on success, 1 where the reference's forward() fails, -1 for an unknown code
inline int forward(int epsg, double lon, double lat, double* x, double* y)
{
    if (epsg == 4326) {
        *x = lon * 180.0 / M_PI;
        *y = lat * 180.0 / M_PI;
        return 0;
    }
    if (epsg > 32600 && epsg < 32800) {
        int zone;
        bool isnorth;
        if (epsg <= 32660) { zone = epsg - 32600; isnorth = true; }
        else if (epsg > 32700 && epsg <= 32760) { zone = epsg - 32700; isnorth = false; }
        else return -1;
        const double lon0 = ((zone - 0.5) * (M_PI / 30.)) - M_PI;
        const double f = kE2 / (1. + std::sqrt(1 - kE2));
        const double n = f / (2. - f);
        double cbg[6], gtu[6];
        cbg[0] = n * (-2 + n * ((2. / 3.) + n * ((4. / 3.) + n * ((-82. / 45.) + n * ((32. / 45.) + n * (4642. / 4725.))))));
        cbg[1] = std::pow(n, 2) * ((5. / 3.) + n * ((-16. / 15.) + n * ((-13. / 9.) + n * ((904. / 315.) + n * (-1522. / 945.)))));
        cbg[2] = std::pow(n, 3) * ((-26. / 15.) + n * ((34. / 21.) + n * ((8. / 5.) + n * (-12686. / 2835.))));
        cbg[3] = std::pow(n, 4) * ((1237. / 630.) + n * ((-12. / 5.) + n * (-24832. / 14175.)));
        cbg[4] = std::pow(n, 5) * ((-734. / 315.) + n * (109598. / 31185.));
        cbg[5] = std::pow(n, 6) * (444337. / 155925.);
        const double Qn = (0.9996 / (1. + n)) * (1. + n * n * ((1. / 4.) + n * n * ((1. / 64.) + ((n * n) / 256.))));
        gtu[0] = n * (.5 + n * ((-2. / 3.) + n * ((5. / 16.) + n * ((41. / 180.) + n * ((-127. / 288.) + n * (7891. / 37800.))))));
        gtu[1] = std::pow(n, 2) * ((13. / 48.) + n * ((-3. / 5.) + n * ((557. / 1440.) + n * ((281. / 630.) + n * (-1983433. / 1935360.)))));
        gtu[2] = std::pow(n, 3) * ((61. / 240.) + n * ((-103. / 140.) + n * ((15061. / 26880.) + n * (167603. / 181440.))));
        gtu[3] = std::pow(n, 4) * ((49561. / 161280.) + n * ((-179. / 168.) + n * (6601661. / 7257600.)));
        gtu[4] = std::pow(n, 5) * ((34729. / 80640.) + n * (-3418889. / 1995840.));
        gtu[5] = std::pow(n, 6) * (212378941. / 319334400.);
        const double Z = clens(cbg, 6, 0.);
        const double Zb = -Qn * (Z + clens(gtu, 6, 2 * Z));
        const double gauss = clens(cbg, 6, 2. * lat) + lat;
        const double lam = lon - lon0;
        double Cn = std::atan2(std::sin(gauss), std::cos(lam) * std::cos(gauss));
        double Ce = std::atan2(std::sin(lam) * std::cos(gauss),
                               std::hypot(std::sin(gauss), std::cos(gauss) * std::cos(lam)));
        Ce = std::asinh(std::tan(Ce));
        double dCn, dCe;
        Cn += clenS(gtu, 6, 2 * Cn, 2 * Ce, dCn, dCe);
        Ce += dCe;
        if (std::fabs(Ce) > 2.623395162778) return 1;
        *x = (Qn * Ce * kA) + 500000.;
        *y = (((Qn * Cn) + Zb) * kA) + (isnorth ? 0. : 10000000.);
        return 0;
    }
    if (epsg == 3031 || epsg == 3413) {
        const bool isnorth = epsg == 3413;
        const double lat_ts = (isnorth ? 70. : 71.) * M_PI / 180.;
        const double lon0 = isnorth ? -45. * (M_PI / 180.) : 0.;
        const double e = std::sqrt(kE2);
        double akm1 = std::cos(lat_ts) / pj_tsfn(lat_ts, std::sin(lat_ts), e);
        akm1 *= kA / std::sqrt(1. - (std::pow(e, 2) * std::pow(std::sin(lat_ts), 2)));
        const double lam = lon - lon0;
        const double phi = lat * (isnorth ? 1. : -1.);
        const double temp = akm1 * pj_tsfn(phi, std::sin(phi), e);
        *x = temp * std::sin(lam);
        *y = -temp * std::cos(lam) * (isnorth ? 1. : -1.);
        return 0;
    }
    if (epsg == 6933) {
        const double lat_ts = M_PI / 6.;
        const double k0 = std::cos(lat_ts) / std::sqrt(1. - (kE2 * std::pow(std::sin(lat_ts), 2)));
        const double e = std::sqrt(kE2), one_es = 1. - kE2;
        *x = k0 * lon * kA;
        *y = (.5 * kA * pj_qsfn(std::sin(lat), e, one_es)) / k0;
        return 0;
    }
    return -1;
}

} // namespace proj

// DEMInterpolator.cpp:592-611: project to the raster's CRS, then interpolateXY.
inline double dem_interp_lonlat(const I3B_DEM& d, double lon, double lat)
{
    if (!d.have_raster) return d.ref_height;
    double x, y;
    if (proj::forward(d.epsg, lon, lat, &x, &y) != 0) return d.ref_height;
    return dem_interp_xy(d, x, y);
}

} // namespace tdbp_oracle
