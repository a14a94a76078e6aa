// FP64 device geometry for the per-pixel target solve and the per-pulse table.
// Behavioural references (what must be matched, not how):
//   orbit        cxx/isce3/core/detail/InterpolateOrbit.icc:15-109,116-155,161-193
//   ellipsoid    cxx/isce3/core/Ellipsoid.h:99-102,152-158,177-224
//   Brent        cxx/isce3/math/RootFind1dBracket.icc:57-216
//   rdr2geo      cxx/isce3/geometry/detail/Rdr2Geo.icc:175-242
//   geo2rdr      cxx/isce3/geometry/detail/Geo2Rdr.icc:185-238
//   LUT2d / DEM  cxx/isce3/core/LUT2d.cpp:127-160, geometry/DEMInterpolator.cpp:592-659,
//                core/{Bilinear,Bicubic,Spline2d,NearestNeighbor}Interpolator.cpp
//   tropo        cxx/isce3/focus/DryTroposphereModel.icc:10-29
// Everything is POD + templates: no device-side new, no virtual dispatch (the
// reference's device twins use both, cuda/geometry/gpuDEMInterpolator.cu:69-90).
#pragma once
#include <cfloat>
#include <cmath>

#include "common.cuh"

namespace i3b {

enum { BORDER_ERROR = 0, BORDER_EXTRAPOLATE = 1, BORDER_FILLNAN = 2 };

__device__ inline D3 ld3(const double* __restrict__ p, int i)
{
    return {p[3 * i], p[3 * i + 1], p[3 * i + 2]};
}

__device__ inline D3 nan3()
{
    const double q = nan("");
    return {q, q, q};
}

// ---- orbit -----------------------------------------------------------------

__device__ inline int linspace_search(double first, double spacing, int size, double val)
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

// Cubic Hermite through 4 state vectors (positions and velocities):
// interpolateOrbitHermite, InterpolateOrbit.icc:15-109.  The state vectors are uniformly
// spaced (Orbit holds a Linspace), so with u_k = (t - t_k)/dt the reference's basis
//   h_i = prod_{j!=i} (t - t_j)/(t_i - t_j),  sum_i = sum_{j!=i} 1/(t_i - t_j), ...
// reduces to products of the u_k with small rational constants: the same polynomial, no
// divisions (the reference form costs 36 FP64 divisions per call, and the call sits inside
// the geo2rdr root finder).  Differs from the literal form by rounding only (~1e-15 rel.).
__device__ inline void orbit_hermite(const DevOrbit& o, double t, D3* pos, D3* vel)
{
    int idx = linspace_search(o.t0, o.dt, o.n, t) - 2;
    idx = min(max(idx, 0), o.n - 4);
    const double inv_dt = 1.0 / o.dt;
    double u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = (t - (o.t0 + (idx + i) * o.dt)) * inv_dt;
    // c_i = 1 / prod_{j != i} (i - j),  S_i = sum_{j != i} 1 / (i - j)
    const double c[4] = {-1.0 / 6.0, 0.5, -0.5, 1.0 / 6.0};
    const double S[4] = {-11.0 / 6.0, -0.5, 0.5, 11.0 / 6.0};
    const double u01 = u[0] * u[1], u23 = u[2] * u[3];
    // products of the other three u's, and sums of their pairwise products
    const double P3[4] = {u[1] * u23, u[0] * u23, u01 * u[3], u01 * u[2]};
    const double E2[4] = {fma(u[1], u[2] + u[3], u23), fma(u[0], u[2] + u[3], u23),
                          fma(u[3], u[0] + u[1], u01), fma(u[2], u[0] + u[1], u01)};
    D3 p = {0, 0, 0}, v = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double h = c[i] * P3[i];
        const double hd_dt = c[i] * E2[i];          // hdot_i * dt
        const double f0 = fma(-2.0 * S[i], u[i], 1.0);
        const double f1 = u[i] * o.dt;
        const double g1 = fma(2.0 * hd_dt, u[i], h);
        const double g0 = 2.0 * inv_dt * (f0 * hd_dt - S[i] * h);
        const D3 P = ld3(o.pos, idx + i), V = ld3(o.vel, idx + i);
        p = p + (h * h) * (P * f0 + V * f1);
        v = v + h * (P * g0 + V * g1);
    }
    *pos = p;
    *vel = v;
}

// Eighth-order Legendre (Lagrange on 9 equispaced vectors).
__device__ inline void orbit_legendre(const DevOrbit& o, double t, D3* pos, D3* vel)
{
    int idx = linspace_search(o.t0, o.dt, o.n, t) - 5;
    idx = min(max(idx, 0), o.n - 9);
    const double ta = o.t0 + idx * o.dt, tb = o.t0 + (idx + 8) * o.dt;
    const double trel = 8. * (t - ta) / (tb - ta);
    double teller = 1.;
#pragma unroll
    for (int i = 0; i < 9; ++i) teller *= trel - i;
    if (teller == 0.) {
        const int i = (int) trel;
        *pos = ld3(o.pos, idx + i);
        *vel = ld3(o.vel, idx + i);
        return;
    }
    const double noemer[9] = {40320.0, -5040.0, 1440.0, -720.0, 576.0,
                              -720.0,  1440.0,  -5040.0, 40320.0};
    D3 p = {0, 0, 0}, v = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const double coeff = (teller / noemer[i]) / (trel - i);
        p = p + coeff * ld3(o.pos, idx + i);
        v = v + coeff * ld3(o.vel, idx + i);
    }
    *pos = p;
    *vel = v;
}

// LEG = false: the caller knows (at compile time) that no orbit on its path is Legendre
// I3B_OUTLINE_MORE = 1: the ellipsoid conversion and the orbit interpolation as real functions
// too (see the note at the raster samplers)
#ifndef I3B_OUTLINE_MORE
#define I3B_OUTLINE_MORE 0
#endif
#if I3B_OUTLINE_MORE
#define I3B_MAYBE_NOINLINE __noinline__
#else
#define I3B_MAYBE_NOINLINE
#endif

template<bool LEG = true>
__device__ I3B_MAYBE_NOINLINE inline int orbit_interpolate(const DevOrbit& o, double t, int border, D3* pos, D3* vel)
{
    const int method = LEG ? o.method : (int) I3B_ORBIT_HERMITE;
    const int need = method == I3B_ORBIT_LEGENDRE ? 9 : 4;
    if (o.n < need) return I3B_ORBIT_INTERP_SIZE_ERROR;
    const double tstart = o.t0, tend = o.t0 + (o.n - 1) * o.dt;
    if (t < tstart || t > tend) {
        if (border == BORDER_FILLNAN) {
            *pos = nan3();
            *vel = nan3();
        }
        if (border != BORDER_EXTRAPOLATE) return I3B_ORBIT_INTERP_DOMAIN_ERROR;
    }
    if (method == I3B_ORBIT_HERMITE) {
        orbit_hermite(o, t, pos, vel);
        return I3B_SUCCESS;
    }
    if (LEG && method == I3B_ORBIT_LEGENDRE) {
        orbit_legendre(o, t, pos, vel);
        return I3B_SUCCESS;
    }
    return I3B_ORBIT_INTERP_UNKNOWN_METHOD;
}

// ---- WGS84 ellipsoid ------------------------------------------------------

__device__ inline D3 llh_to_xyz(D3 llh)
{
    double slat, clat, slon, clon;
    sincos(llh.y, &slat, &clat);
    sincos(llh.x, &slon, &clon);
    const double re = kA / sqrt(1.0 - kE2 * slat * slat);
    return {(re + llh.z) * clat * clon, (re + llh.z) * clat * slon,
            (re * (1.0 - kE2) + llh.z) * slat};
}

// Vermeille (2002) closed form.
__device__ I3B_MAYBE_NOINLINE inline D3 xyz_to_llh(D3 p3)
{
    const double e4 = kE2 * kE2, a2 = kA * kA;
    const double rho2 = p3.x * p3.x + p3.y * p3.y;
    const double p = rho2 / a2;
    const double q = (1. - kE2) * (p3.z * p3.z) / a2;
    const double r = (p + q - e4) / 6.;
    const double s = (e4 * p * q) / (4. * r * r * r);
    const double t = cbrt(1. + s + sqrt(s * (2. + s)));
    const double u = r * (1. + t + (1. / t));
    const double rv = sqrt(u * u + e4 * q);
    const double w = (kE2 * (u + rv - q)) / (2. * rv);
    const double k = sqrt(u + rv + w * w) - w;
    const double d = (k * sqrt(rho2)) / (k + kE2);
    D3 llh;
    llh.y = atan2(p3.z, d);
    llh.x = atan2(p3.y, p3.x);
    llh.z = ((k + kE2 - 1.) * sqrt(d * d + p3.z * p3.z)) / k;
    return llh;
}

// Height above the ellipsoid only (same closed form; skips the two atan2 of the angles).
__device__ inline double xyz_to_height(D3 p3)
{
    const double e4 = kE2 * kE2, a2 = kA * kA;
    const double rho2 = p3.x * p3.x + p3.y * p3.y;
    const double p = rho2 / a2;
    const double q = (1. - kE2) * (p3.z * p3.z) / a2;
    const double r = (p + q - e4) / 6.;
    const double s = (e4 * p * q) / (4. * r * r * r);
    const double t = cbrt(1. + s + sqrt(s * (2. + s)));
    const double u = r * (1. + t + (1. / t));
    const double rv = sqrt(u * u + e4 * q);
    const double w = (kE2 * (u + rv - q)) / (2. * rv);
    const double k = sqrt(u + rv + w * w) - w;
    const double d = (k * sqrt(rho2)) / (k + kE2);
    return ((k + kE2 - 1.) * sqrt(d * d + p3.z * p3.z)) / k;
}

__device__ inline D3 n_vector(double lon, double lat)
{
    double slat, clat, slon, clon;
    sincos(lat, &slat, &clat);
    sincos(lon, &slon, &clon);
    return {clat * clon, clat * slon, slat};
}

// ---- 2-D samplers -----------------------------------------------------------

// CLAMP: indices are clamped to the array (edge replication).  Used for LUT2d, whose
// coordinates are only clamped to [0, n-1] (LUT2d.cpp:151-152) while the bicubic / spline
// stencils reach up to 3 samples beyond that -- the reference reads outside its matrix
// there (host UB); on a device that is garbage or a fault.  DEM sampling keeps the
// reference's own [2, n-1) margin test and needs no clamp.
template<typename U, bool CLAMP = false>
struct Grid2d {
    const U* __restrict__ data;
    int rows, cols;
    __device__ U operator()(int r, int c) const
    {
        if (CLAMP) {
            r = min(max(r, 0), rows - 1);
            c = min(max(c, 0), cols - 1);
        }
        return data[(size_t) r * cols + c];
    }
};

template<typename U, class G>
__device__ inline U bilinear(double x, double y, const G& z)
{
    const int x1 = (int) floor(x), x2 = (int) ceil(x);
    const int y1 = (int) floor(y), y2 = (int) ceil(y);
    const U q11 = z(y1, x1), q12 = z(y2, x1), q21 = z(y1, x2), q22 = z(y2, x2);
    if (y1 == y2 && x1 == x2) return q11;
    if (y1 == y2) return U((x2 - x) / (x2 - x1)) * q11 + U((x - x1) / (x2 - x1)) * q21;
    if (x1 == x2) return U((y2 - y) / (y2 - y1)) * q11 + U((y - y1) / (y2 - y1)) * q12;
    const U den = U((x2 - x1) * (y2 - y1));
    return (q11 * U((x2 - x) * (y2 - y))) / den + (q21 * U((x - x1) * (y2 - y))) / den +
           (q12 * U((x2 - x) * (y - y1))) / den + (q22 * U((x - x1) * (y - y1))) / den;
}

template<typename U>
__device__ inline U catmull_rom(U p0, U p1, U p2, U p3, double tf)
{
    const double tc = 1. - tf;
    return (U(tf) * (p2 - p0 * U(tc * tc) + (p2 * U(tc * 3. + 1.) - p3 * U(tc)) * U(tf)) +
            p1 * U(tf * tf * (tf * 3. - 5.) + 2.)) / U(2.);
}

template<typename U, class G>
__device__ inline U bicubic(double x, double y, const G& z)
{
    const int x0 = (int) floor(x), y0 = (int) floor(y);
    U rowv[4];
#pragma unroll
    for (int i = -1; i < 3; ++i)
        rowv[i + 1] = catmull_rom<U>(z(y0 + i, x0 - 1), z(y0 + i, x0), z(y0 + i, x0 + 1),
                                     z(y0 + i, x0 + 2), x - x0);
    return catmull_rom<U>(rowv[0], rowv[1], rowv[2], rowv[3], y - y0);
}

// Natural cubic spline through N points, second derivatives by the usual
// tridiagonal sweep; "biquintic" in the reference is this with N = 6.
template<typename U, int N>
__device__ inline void spline_init(const U* Y, U* R, U* Q)
{
    Q[0] = U(0);
    R[0] = U(0);
#pragma unroll
    for (int i = 1; i < N - 1; ++i) {
        const U p = U(1.0) / (U(0.5) * Q[i - 1] + U(2.0));
        Q[i] = U(-0.5) * p;
        R[i] = (U(3.0) * (Y[i + 1] - U(2.0) * Y[i] + Y[i - 1]) - U(0.5) * R[i - 1]) * p;
    }
    R[N - 1] = U(0);
#pragma unroll
    for (int i = N - 2; i > 0; --i) R[i] = Q[i] * R[i + 1] + R[i];
}

template<typename U, int N>
__device__ inline U spline_eval(double x, const U* Y, const U* R)
{
    const U denom = U(6.0);
    if (x < 1.0) return Y[0] + U(x - 1.0) * (Y[1] - Y[0] - (R[1] / denom));
    if (x > N) return Y[N - 1] + U(x - N) * (Y[N - 1] - Y[N - 2] + (R[N - 2] / denom));
    const int j = (int) floor(x);
    const U xx = U(x - j);
    // static unrolled select keeps Y/R in registers (no local-memory indexing)
    U yj = Y[1], yjm = Y[0], rj = R[1], rjm = R[0];
#pragma unroll
    for (int q = 2; q < N; ++q)
        if (j == q) {
            yj = Y[q];
            yjm = Y[q - 1];
            rj = R[q];
            rjm = R[q - 1];
        }
    if (j >= N) { // x == N exactly: reference indexes Y[N] (out of range); clamp
        yj = Y[N - 1];
        yjm = Y[N - 2];
        rj = R[N - 1];
        rjm = R[N - 2];
    }
    const U t0 = yj - yjm - (rjm / U(3.0)) - (rj / denom);
    const U t1 = xx * ((rjm / U(2.0)) + (xx * ((rj - rjm) / denom)));
    return yjm + (xx * (t0 + t1));
}

template<typename U, class G>
__device__ inline U biquintic(double x, double y, const G& z)
{
    constexpr int N = 6;
    int i0 = (int) y, j0 = (int) x;
    i0 = i0 - (N / 2) + 1;
    j0 = j0 - (N / 2) + 1;
    U A[N], R[N], Q[N], HC[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const int indi = min(max(i0 + i, 0), z.rows - 2);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const int indj = min(max(j0 + j, 0), z.cols - 2);
            A[j] = z(indi + 1, indj + 1);
        }
        spline_init<U, N>(A, R, Q);
        HC[i] = spline_eval<U, N>(x - j0, A, R);
    }
    spline_init<U, N>(HC, R, Q);
    return spline_eval<U, N>(y - i0, HC, R);
}

// Sinc2dInterpolator::interp_impl / _sinc_eval_2d (core/Sinc2dInterpolator.cpp:45-101): 8 x 8
// taps from the table row nearest below the fractional offset; 0 within half a kernel of the
// array edges.
template<typename U, class G>
__device__ inline U sinc2d(double x, double y, const G& z, const double* __restrict__ kern)
{
    constexpr int LEN = 8, HALF = 4, SUB = 8192;
    const int ix = (int) floor(x), iy = (int) floor(y);
    const double fx = x - ix, fy = y - iy;
    if (ix < HALF - 1 || ix > z.cols - HALF - 1) return U(0);
    if (iy < HALF - 1 || iy > z.rows - HALF - 1) return U(0);
    const int xx = ix + HALF, yy = iy + HALF;
    const int ifx = min(max(0, (int) (fx * SUB)), SUB - 1), ify = min(max(0, (int) (fy * SUB)), SUB - 1);
    const double* kx = kern + (size_t) ifx * LEN;
    const double* ky = kern + (size_t) ify * LEN;
    U ret(0);
    for (int i = 0; i < LEN; ++i) {
        const U wy = U(ky[i]);
#pragma unroll
        for (int j = 0; j < LEN; ++j) ret += z(yy - i, xx - j) * wy * U(kx[j]);
    }
    return ret;
}

template<typename U, class G>
__device__ inline U interp2d(int method, double x, double y, const G& z, const double* sinc = nullptr)
{
    if (method == I3B_INTERP_SINC && sinc) return sinc2d<U, G>(x, y, z, sinc);
    switch (method) {
    case I3B_INTERP_BICUBIC: return bicubic<U, G>(x, y, z);
    case I3B_INTERP_BIQUINTIC: return biquintic<U, G>(x, y, z);
    case I3B_INTERP_NEAREST: return z((int) round(y), (int) round(x));
    default: return bilinear<U, G>(x, y, z);
    }
}

// The 2-D samplers of a raster (five interpolation methods each, the spline and sinc ones
// hundreds of instructions) are real functions: the solvers that call them from inside a
// root finder were instruction-fetch bound with them inlined at every call site (raster-DEM
// target solve: 38 k instructions, 14 warps stalled on "no instruction" per issue, ncu round 2).
__device__ __noinline__ inline double lut_sample_outlined(const double* data, int length, int width, int method,
                                                          double xi, double yi, const double* sinc)
{
    const Grid2d<double, true> g {data, length, width};
    return interp2d<double, Grid2d<double, true>>(method, xi, yi, g, sinc);
}

__device__ __noinline__ inline double dem_sample_outlined(const float* data, int length, int width, int method,
                                                          double col, double row, const double* sinc)
{
    const Grid2d<float> g {data, length, width};
    return interp2d<float, Grid2d<float>>(method, col, row, g, sinc);
}

// LUT = false: the caller knows that no Doppler LUT on its path holds data
template<bool LUT = true>
__device__ inline double lut2d_eval(const DevLUT2d& l, double y, double x)
{
    if (!LUT || !l.have_data) return l.ref_value;
    double xi = (x - l.xstart) / l.dx;
    double yi = (y - l.ystart) / l.dy;
    xi = fmin(fmax(xi, 0.0), l.width - 1.0);
    yi = fmin(fmax(yi, 0.0), l.length - 1.0);
    return lut_sample_outlined(l.data, l.length, l.width, l.method, xi, yi, l.sinc);
}

// LUT2d::contains (core/LUT2d.h:84-95)
__device__ inline bool lut2d_contains(const DevLUT2d& l, double y, double x)
{
    if (!l.have_data) return true;
    const double i = (x - l.xstart) / l.dx, j = (y - l.ystart) / l.dy;
    return (i >= 0.0 && i <= l.width - 1.0) && (j >= 0.0 && j <= l.length - 1.0);
}

// DEMInterpolator::interpolateLonLat -> interpolateXY (DEMInterpolator.cpp:592-659): project
// to the raster's CRS, then sample; the longitude wrapping applies to EPSG:4326 only.
__device__ inline double dem_interp_lonlat(const DevDEM& d, double lon, double lat)
{
    if (!d.have_raster) return d.ref_height;
    double x, y;
    // (the reference ignores forward()'s status and samples an unset point: out of the raster)
    if (proj_forward(d.proj, lon, lat, &x, &y) != 0) return d.ref_height;
    if (d.proj.kind == PROJ_LONLAT) {
        if (x > 360 || x < -360) x = fmod(x, 360.);
        if (x < -180) x += 360;
        if (x - 360 >= d.xstart) {
            x -= 360;
        } else if (x < d.xstart && x + 360 >= d.xstart) {
            x += 360;
        } else if (x < d.xstart) {
            return d.ref_height;
        }
    } else if (x < d.xstart) {
        return d.ref_height;
    }
    const double row = (y - d.ystart) / d.dy;
    const double col = (x - d.xstart) / d.dx;
    const int irow = (int) floor(row), icol = (int) floor(col);
    if (irow < 2 || irow >= d.length - 1) return d.ref_height;
    if (icol < 2 || icol >= d.width - 1) return d.ref_height;
    return dem_sample_outlined(d.data, d.length, d.width, d.method, col, row, d.sinc);
}

// ---- Brent's bracketing root finder ---------------------------------------------

__device__ inline bool opposite_sign(double a, double b) { return signbit(a) != signbit(b); }

// (The function is evaluated at ONE place in the code -- the initial end-point evaluations and
// the one per iteration share it -- so that a large `f`, inlined, appears once.  Arithmetic and
// control flow are those of the reference's loop, math/detail/RootFind1dBracket.icc.)
template<class F>
__device__ inline int brent(double a, double b, F f, const double tol, double* root)
{
    if (tol < 0.0) return I3B_INVALID_TOLERANCE;
    double c = 0.0, d = 0.0, e = 0.0, fa = 0.0, fb = 0.0, fc = 0.0, p, q, r, s, tol1;
    int maxiter = 0, it = 0, stage = 0; // stage 0: f(a) pending, 1: f(b) pending, 2: iterating
    double x = a;
#pragma unroll 1
    for (;;) {
        const double fx = f(x);
        if (stage == 0) {
            fa = fx;
            if (fa == 0.0) {
                *root = a;
                return I3B_SUCCESS;
            }
            stage = 1;
            x = b;
            continue;
        }
        if (stage == 1) {
            fb = fx;
            if (fb == 0.0) {
                *root = b;
                return I3B_SUCCESS;
            }
            if (!opposite_sign(fa, fb)) return I3B_INVALID_INTERVAL;
            c = a;
            fc = fa;
            e = d = b - a;
            tol1 = tol > 0.0 ? tol : DBL_EPSILON;
            maxiter = 3 * (int) ceil(log2(fabs((a - b) / tol1)));
            stage = 2;
        } else {
            fb = fx;
            if (!opposite_sign(fb, fc)) {
                c = a;
                fc = fa;
                e = d = b - a;
            }
            ++it;
        }
        if (it >= maxiter) break;
        if (fabs(fc) < fabs(fb)) {
            a = b; b = c; c = a;
            fa = fb; fb = fc; fc = fa;
        }
        tol1 = 2 * DBL_EPSILON * fabs(b) + 0.5 * tol;
        const double xm = 0.5 * (c - b);
        if ((fabs(xm) <= tol1) || (fb == 0.0)) {
            *root = b;
            return I3B_SUCCESS;
        }
        if ((fabs(e) < tol1) || (fabs(fa) <= fabs(fb))) {
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
            if (((2 * p) >= (3 * xm * q - fabs(tol1 * q))) || (p >= fabs(0.5 * s * q))) {
                e = d = xm;
            } else {
                d = p / q;
            }
        }
        a = b;
        fa = fb;
        if (fabs(d) <= tol1) b = (xm <= 0.0) ? b - tol1 : b + tol1;
        else b = b + d;
        x = b;
    }
    *root = b;
    return I3B_FAILED_TO_CONVERGE;
}

// ---- rdr2geo / geo2rdr (bracketing solvers) -----------------------------------------

// Target ECEF on the DEM for (aztime, range, doppler).  Returns I3B_SUCCESS, a soft
// ErrorCode, or I3B_EXC_OUT_OF_RANGE when aztime is outside the orbit (the CPU
// reference throws there, core/Orbit.cpp:78-83).
template<bool RASTER = true, bool LEG = true>
__device__ inline int rdr2geo_bracket(double aztime, double slant_range, double doppler,
                                      const DevOrbit& orbit, const DevDEM& dem, double wavelength,
                                      int side, const I3B_Rdr2GeoBracketParams& prm, D3* xyz)
{
    D3 radar, velocity;
    if (orbit_interpolate<LEG>(orbit, aztime, BORDER_ERROR, &radar, &velocity) != I3B_SUCCESS)
        return I3B_EXC_OUT_OF_RANGE;
    const double speed = norm(velocity);
    const D3 along = velocity / speed;
    const D3 right = unit(cross(along, radar));
    const D3 down = cross(along, right);
    const D3 horizontal = (side == I3B_LOOK_RIGHT) ? right : -1.0 * right;
    const double sin_squint = doppler * wavelength / (2 * speed);
    const double cos_squint = sqrt(1.0 - sin_squint * sin_squint);
    const D3 center = radar + (sin_squint * slant_range) * along;
    const double radius = cos_squint * slant_range;
    auto get_xyz = [&](double look) {
        double sl, cl;
        sincos(look, &sl, &cl);
        return center + (radius * sl) * horizontal + (radius * cl) * down;
    };
    const double tol_look = prm.tol_height / radius;
    double look = 0.0;
    if (!RASTER || !dem.have_raster) {
        // Constant-height DEM: the height error is monotonic in the look angle, so the root in
        // [look_min, look_max] is unique.  Bracket it tightly around the spherical-Earth
        // solution first (same root finder, same tolerance: the result is the root to within
        // tol_look either way); the full interval of the reference is the fallback.
        auto dh_flat = [&](double look_) { return xyz_to_height(get_xyz(look_)) - dem.ref_height; };
        const double sinpsi = radar.z / norm(radar);   // geocentric latitude of the platform
        const double re = kA * sqrt((1.0 - kE2) / (1.0 - kE2 * (1.0 - sinpsi * sinpsi))) + dem.ref_height;
        const double cguess = (re * re - dot(center, center) - radius * radius) /
                              (2.0 * radius * dot(radar, down));
        if (fabs(cguess) < 1.0) {
            double g = acos(cguess);
            // Newton on the height error from that guess: d(height)/d(look) is the ellipsoid
            // normal at the point dotted with the tangent of the range/Doppler circle.  Accepted
            // only when the residual height proves the look angle is within tol_look / 2 of the
            // root (|dh| <= |dh/dlook| * tol_look / 2), i.e. the root finder's own accuracy.
#pragma unroll 1
            for (int it = 0; it < 4; ++it) {
                double sl, cl;
                sincos(g, &sl, &cl);
                const D3 X = center + (radius * sl) * horizontal + (radius * cl) * down;
                const double e = xyz_to_height(X) - dem.ref_height;
                const D3 n = unit(D3 {X.x, X.y, X.z / (1.0 - kE2)});
                const double slope = dot(n, (radius * cl) * horizontal - (radius * sl) * down);
                if (!(fabs(slope) > 1e-3 * radius)) break; // grazing geometry: leave it to Brent
                if (fabs(e) <= 0.5 * tol_look * fabs(slope) && g >= prm.look_min && g <= prm.look_max) {
                    *xyz = X;
                    return I3B_SUCCESS;
                }
                g -= e / slope;
                if (!(g == g)) break;
            }
            g = acos(cguess);
            const double lo = fmax(g - 0.02, prm.look_min), hi = fmin(g + 0.02, prm.look_max);
            if (lo < hi && brent(lo, hi, dh_flat, tol_look, &look) == I3B_SUCCESS) {
                *xyz = get_xyz(look);
                return I3B_SUCCESS;
            }
        }
        const int err = brent(prm.look_min, prm.look_max, dh_flat, tol_look, &look);
        if (err != I3B_SUCCESS) return err;
        *xyz = get_xyz(look);
        return I3B_SUCCESS;
    }
    auto dh = [&](double look_) {
        const D3 llh = xyz_to_llh(get_xyz(look_));
        return llh.z - dem_interp_lonlat(dem, llh.x, llh.y);
    };
    // (A narrower bracket was tried for rasters -- the look angles at which the circle crosses
    // the raster's height range, licensed by a measured gradient bound that rules out layover:
    // Brent needs ~8 evaluations from the full interval and the bracket's own set-up costs as
    // much as it saves; 41.8 ms against 33.9 ms on the C4 frame.  The reference's interval it is.)
    const int err = brent(prm.look_min, prm.look_max, dh, tol_look, &look);
    if (err != I3B_SUCCESS) return err;
    *xyz = get_xyz(look);
    return I3B_SUCCESS;
}

// `t_guess`: where the root is expected (the output line's time when both geometries share
// the orbit).  The Doppler error is first bracketed in a few seconds around it -- same root
// finder and tolerance as the reference, so the answer is the root to within tol_aztime
// either way -- and the reference's full interval is the fallback (also taken when the
// guess is NaN).
template<bool LEG = true, bool LUT = true>
__device__ inline int geo2rdr_bracket(D3 x, const DevOrbit& orbit, const DevLUT2d& dop,
                                      double wavelength, int side,
                                      const I3B_Geo2RdrBracketParams& prm, double* aztime,
                                      double* range, double t_guess = nan(""))
{
    const double orbit_start = orbit.t0, orbit_end = orbit.t0 + (orbit.n - 1) * orbit.dt;
    double t0, t1;
    const bool have_lut = LUT && dop.have_data;
    if (prm.has_time_start) t0 = prm.time_start;
    else t0 = have_lut ? fmax(orbit_start, dop.ystart) : orbit_start;
    if (prm.has_time_end) t1 = prm.time_end;
    else t1 = have_lut ? fmin(orbit_end, dop.ystart + dop.dy * (dop.length - 1)) : orbit_end;
    D3 xp, v, r;
    auto doppler_error = [&](double t) {
        orbit_interpolate<LEG>(orbit, t, BORDER_FILLNAN, &xp, &v);
        r = x - xp;
        const double rnorm = norm(r);
        const double fd = lut2d_eval<LUT>(dop, t, rnorm);
        return 2.0 / wavelength * dot(v, r) / rnorm - fd;
    };
    int err = I3B_INVALID_INTERVAL;
    if (t_guess == t_guess) {
        // When input and output geometry share orbit and Doppler model the guess IS the root
        // (geo2rdr inverts rdr2geo): a sign change across [guess - tol/2, guess + tol/2] proves
        // the root lies within tol/2 of the guess -- the accuracy the root finder itself
        // stops at -- for two evaluations instead of a search.
        const double hw = 0.5 * prm.tol_aztime;
        if (hw > 0.0 && t_guess - hw >= t0 && t_guess + hw <= t1) {
            double fv[2];
#pragma unroll 1
            for (int i = 0; i < 2; ++i) fv[i] = doppler_error(i ? t_guess + hw : t_guess - hw);
            if (fv[0] == fv[0] && fv[1] == fv[1] && opposite_sign(fv[0], fv[1])) {
                *aztime = t_guess;
                err = I3B_SUCCESS;
            }
        }
    }
    if (err != I3B_SUCCESS) {
        double lo = t0, hi = t1;
        bool narrow = false;
        if (t_guess == t_guess) {
            const double l4 = fmax(t_guess - 4.0, t0), h4 = fmin(t_guess + 4.0, t1);
            if (l4 < h4) {
                lo = l4;
                hi = h4;
                narrow = true;
            }
        }
#pragma unroll 1
        for (;;) { // (one call site of the root finder: the Doppler error is inlined once)
            err = brent(lo, hi, doppler_error, prm.tol_aztime, aztime);
            if (err == I3B_SUCCESS || !narrow) break;
            lo = t0;
            hi = t1;
            narrow = false;
        }
    }
    if (err != I3B_SUCCESS) return err;
    orbit_interpolate<LEG>(orbit, *aztime, BORDER_FILLNAN, &xp, &v);
    r = x - xp;
    *range = norm(r);
    const bool positive = dot(cross(r, v), xp) > 0;
    if ((side == I3B_LOOK_RIGHT) ^ positive) return I3B_WRONG_LOOK_SIDE;
    return I3B_SUCCESS;
}

__device__ inline double dry_tropo_tsx(D3 p, D3 llh)
{
    constexpr double ZPD = 2.3, H = 6000.;
    const D3 x = llh_to_xyz(llh);
    const D3 r_hat = unit(p - x);
    const D3 n_hat = unit(n_vector(llh.x, llh.y));
    return 2. * ZPD * exp(-llh.z / H) / (kC * dot(r_hat, n_hat));
}

} // namespace i3b
