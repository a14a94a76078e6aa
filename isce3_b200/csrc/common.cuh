// Shared device-side types of the B200 TDBP backend.
//
// HBM layout:
//   rc        complex64 [pulses resident][rc_pitch]   range-compressed lines, pitch % 2 == 0
//   pulse     PulseRec  [pulses]                       per-pulse orbit-derived constants (80 B)
//   pv        double    [pulses][6]                    per-pulse position, velocity (generic kernel)
//   pix       PixelRec  [out pixels]                   per-pixel target solve result (40 B)
//   acc       double2   [out pixels]                   complex<double> image accumulator
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/isce3_b200_backproject.h"
#include "projections.cuh"

namespace i3b {

constexpr double kC = 299792458.0;           // cxx/isce3/core/Constants.h:50
constexpr double kA = 6378137.0;             // Constants.h:41
constexpr double kE2 = 0.006694379990141317; // Constants.h:44

struct D3 {
    double x, y, z;
};
__host__ __device__ inline D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__host__ __device__ inline D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__host__ __device__ inline D3 operator*(double s, D3 a) { return {s * a.x, s * a.y, s * a.z}; }
__host__ __device__ inline D3 operator*(D3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__host__ __device__ inline D3 operator/(D3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
__host__ __device__ inline double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ inline double norm(D3 a) { return sqrt(dot(a, a)); }
__host__ __device__ inline D3 unit(D3 a) { return a / norm(a); }
__host__ __device__ inline D3 cross(D3 a, D3 b)
{
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// Device views of the flat descriptors (pointers are device pointers).
struct DevOrbit {
    double t0, dt;
    int n, method;
    const double* pos; // [n][3]
    const double* vel;
};

struct DevLUT2d {
    int have_data, bounds_error, method;
    int length, width;
    double ref_value, xstart, ystart, dx, dy;
    const double* data;
    const double* sinc; // method SINC: normalised kernel table [SINC_SUB][SINC_LEN] (sinc_table.h)
};

struct DevDEM {
    int have_raster, epsg, method;
    int length, width;
    double ref_height, xstart, ystart, dx, dy;
    const float* data;
    DevProj proj; // forward projection of the raster's CRS (set up on the host)
    const double* sinc; // method SINC: normalised kernel table [SINC_SUB][SINC_LEN]
};

struct DevKernel {
    int kind, n, taps; // taps = ceil(width)
    int imax;          // TABULATED: n - 2
    double halfwidth, bandwidth;
    float one_dx;     // TABULATED: (float)(1/dx)   (core/Kernels.h:145 stores T=float)
    float cheb_scale; // CHEBY: (float)(4/width)
    const float* data;
};

struct Linspace {
    double first, spacing;
    int size;
    __host__ __device__ double operator[](int i) const { return first + i * spacing; }
};

// Result of the per-pixel target solve (Backproject.cpp:128-199), 40 B.
struct PixelRec {
    double x, y, z; // target ECEF (m)
    double tau_atm; // dry-troposphere two-way delay (s)
    int kstart, kstop; // coherent integration bounds; kstart == kstop == -1: failed pixel
};

// Per-pulse constants for the fast kernel (80 B, 16-B aligned for LDS.128):
//   |x - p|^2 = xx + pp + x.m2p                      (m2p = -2 p)
//   fc*tau    = fc*tau_atm + E + x.vB + Cs*|x - p|   (cycles)
// with A = 2/(v.v - c^2), vB = fc*A*v, E = -fc*A*(p.v), Cs = -fc*A*c
// (cxx/isce3/focus/BistaticDelay.icc:10-17 rearranged; all FP64).
struct __align__(16) PulseRec {
    double m2px, m2py, m2pz, pp;
    double vBx, vBy, vBz, E;
    double Cs, pad;
};

// Status words written by kernels.
struct DevStatus {
    int soft_error;  // last non-success isce3 ErrorCode raised by a pixel
    int hard_error;  // I3B_EXC_* (orbit domain error under border mode Error)
    int window_overflow; // fast kernel: a gather fell outside its staged tile
    int kmin, kmax;      // pulse span [kmin, kmax) needed by the solved pixels
    int premature;       // a row-wavefront launch met a tile whose pulses had not all landed
    unsigned long long pixel_pulses; // sum (kstop - kstart)
};

} // namespace i3b
