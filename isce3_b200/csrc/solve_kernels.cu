// Per-pulse table, per-pixel target solve, generic accumulation and finalisation
// kernels of the B200 TDBP backend (sm_100a).  All geometry here is FP64 and is
// O(pulses) or O(pixels); the O(pixels x pulses) product kernel is accumulate_fast.cu.
#include "geometry.cuh"
#include "kernels.cuh"
#include "launch.h"

#include <algorithm>
#include <climits>

namespace i3b {

// ---- per-pulse table ---------------------------------------------------------
// Replaces the host loop Backproject.cpp:101-106 (orbit border mode Error).
__global__ void pulse_table_kernel(DevOrbit orbit, Linspace in_time, const double* __restrict__ in_times,
                                   double fc, PulseRec* pulse, double* pv, DevStatus* status)
{
    // pulse[] is indexed from -kPulsePadLo (the pointer passed in is already offset)
    const int k = (int) (blockIdx.x * blockDim.x + threadIdx.x) - kPulsePadLo;
    if (k >= in_time.size + kPulsePadHi) return;
    const bool in_grid = k >= 0 && k < in_time.size;
    D3 p, v;
    // pulses of the input grid: border mode Error like Backproject.cpp:101-106; padding
    // entries: smooth orbit extrapolation (never contributes to a pixel)
    const double tk = in_times ? in_times[k] : in_time[k];
    const int st = orbit_interpolate(orbit, tk, in_grid ? BORDER_ERROR : BORDER_EXTRAPOLATE, &p, &v);
    if (st != I3B_SUCCESS) {
        if (in_grid) status->hard_error = I3B_EXC_OUT_OF_RANGE;
        p = v = nan3();
    }
    if (in_grid) {
        pv[6 * k + 0] = p.x; pv[6 * k + 1] = p.y; pv[6 * k + 2] = p.z;
        pv[6 * k + 3] = v.x; pv[6 * k + 4] = v.y; pv[6 * k + 5] = v.z;
    }
    const double A = 2.0 / (dot(v, v) - kC * kC);
    PulseRec r;
    r.m2px = -2.0 * p.x; r.m2py = -2.0 * p.y; r.m2pz = -2.0 * p.z;
    r.pp = dot(p, p);
    const double fA = fc * A;
    r.vBx = fA * v.x; r.vBy = fA * v.y; r.vBz = fA * v.z;
    r.E = -fA * dot(p, v);
    r.Cs = -fA * kC;
    r.pad = 0.0;
    pulse[k] = r;
}

// ---- per-pixel target solve -------------------------------------------------------
// One thread per output pixel: Backproject.cpp:128-199 (rdr2geo_bracket on the output
// geometry, LLH, geo2rdr_bracket on the input geometry, CPI bounds, dry-troposphere
// delay) fused into one kernel that writes a 40-byte record per pixel instead of the
// reference CUDA path's ~136 B of FP64 side tables (cuda/focus/Backproject.cu:526-640).
__global__ void tile_info_init_kernel(TileInfo* tiles, int n)
{
    const int i = (int) (blockIdx.x * blockDim.x + threadIdx.x);
    if (i < n) tiles[i] = TileInfo {INT_MAX, INT_MIN, 0, 0};
}

// 6 CTAs/SM (<= 85 registers, a few spills): the kernel is latency-bound (FP64 transcendentals,
// DEM loads), more resident warps beat fewer spills -- 154 registers / 3 CTAs was 32 % slower
// on raster DEMs and 16 % on flat ones.
//
// One instantiation per configuration class -- raster or constant-height DEM, any Legendre
// orbit or Hermite only, any Doppler LUT with data or none: the flags are template arguments
// of the solvers, so each instantiation carries only the samplers / orbit interpolators / LUT
// code it can reach.  (One kernel for everything was fetch-bound:
// `no_instruction` 3.5 warps per issue cycle, profiles/r01_ncu_target_solve.md.)
template<bool RASTER, bool LEGENDRE, bool LUT>
__global__ void __launch_bounds__(128, 6)
target_solve_kernel(SolveParams P, PixelRec* __restrict__ pix, float* __restrict__ height,
                    TileInfo* __restrict__ tiles, DevStatus* status)
{
    const long long tid = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long npix = (long long) P.out_lines * P.out_width;
    int kstart = -1, kstop = -1;
    if (tid < npix) {
        const int j = (int) (tid / P.out_width) + P.line0;
        const int i = (int) (tid % P.out_width);
        const double t = P.out_time[j];
        const double r = P.out_range[i];
        const double fD = lut2d_eval<LUT>(P.out_doppler, t, r);
        // LUT2d with bounds_error (LUT2d.cpp:143-150): the CPU reference raises through its
        // error channel, the reference CUDA path silently returns ref_value
        // (gpuLUT2d.cu:178-181).  Here the lookup is clamped like bounds_error=False and the
        // call reports the soft code OutOfBoundsLookup (checked at the solutions, not at the
        // probes of the root finders).
        if (LUT && P.out_doppler.bounds_error && !lut2d_contains(P.out_doppler, t, r))
            status->soft_error = I3B_OUT_OF_BOUNDS_LOOKUP;
        PixelRec rec;
        rec.x = rec.y = rec.z = nan("");
        rec.tau_atm = 0.0;
        float h = nanf("");
        D3 x;
        int st = rdr2geo_bracket<RASTER, LEGENDRE>(t, r, fD, P.out_orbit, P.dem, P.wvl, P.out_side, P.r2g, &x);
        if (st == I3B_EXC_OUT_OF_RANGE) {
            status->hard_error = st;
        } else if (st != I3B_SUCCESS) {
            status->soft_error = I3B_FAILED_TO_CONVERGE;
        } else {
            const D3 llh = xyz_to_llh(x);
            h = (float) llh.z;
            double tc, rc;
            st = geo2rdr_bracket<LEGENDRE, LUT>(x, P.in_orbit, P.in_doppler, P.wvl, P.in_side, P.g2r, &tc, &rc, t);
            if (st != I3B_SUCCESS) {
                status->soft_error = I3B_FAILED_TO_CONVERGE;
            } else {
                if (LUT && P.in_doppler.bounds_error && !lut2d_contains(P.in_doppler, tc, rc))
                    status->soft_error = I3B_OUT_OF_BOUNDS_LOOKUP;
                D3 p, v;
                orbit_interpolate<LEGENDRE>(P.in_orbit, tc, BORDER_FILLNAN, &p, &v);
                const double l = P.wvl * rc * (norm(p) / norm(x)) / (2. * P.ds);
                const double cpi = l / norm(v);
                const double tstart = tc - 0.5 * cpi, tstop = tc + 0.5 * cpi;
                const double t0 = P.in_time.first, dt = P.in_time.spacing;
                kstart = (int) floor((tstart - t0) / dt);
                kstop = (int) ceil((tstop - t0) / dt);
                if (P.in_times) {
                    // explicit pulse times: last pulse at or before tstart, first pulse at or
                    // after tstop (what floor / ceil select on a uniform grid); binary searches
                    const double* T = P.in_times;
                    const int n = P.in_time.size;
                    int lo = 0, hi = n; // upper_bound(tstart)
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (T[mid] <= tstart) lo = mid + 1;
                        else hi = mid;
                    }
                    kstart = lo - 1;
                    lo = 0, hi = n; // lower_bound(tstop)
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (T[mid] < tstop) lo = mid + 1;
                        else hi = mid;
                    }
                    kstop = lo;
                }
                kstart = max(kstart, 0);
                kstop = min(kstop, P.in_time.size);
                double tau_atm = 0.;
                if (P.tropo == I3B_TROPO_TSX) tau_atm = dry_tropo_tsx(p, llh);
                rec.x = x.x; rec.y = x.y; rec.z = x.z;
                rec.tau_atm = tau_atm;
                if (kstop < kstart) kstop = kstart; // empty window -> zero pixel (CPU loop runs 0 times)
            }
        }
        rec.kstart = kstart;
        rec.kstop = kstop;
        pix[tid] = rec;
        if (height) height[tid] = h;
    }
    // block-level reduction of the pulse span and the work count
    int kmin = (kstart >= 0 && kstop > kstart) ? kstart : INT_MAX;
    int kmax = (kstart >= 0 && kstop > kstart) ? kstop : INT_MIN;
    // per-tile summary: one atomic set per group of lanes that share a tile
    {
        int tile = -1;
        if (tid < npix) {
            const int jl = (int) (tid / P.out_width), i = (int) (tid % P.out_width);
            tile = (jl / P.tile_az) * P.tiles_rg + i / P.tile_rg;
        }
        // lanes that share a tile reduce among themselves (hardware warp reduce over the peer
        // mask: correct for any grouping, e.g. warps that straddle two lines or two tiles)
        const unsigned peers = __match_any_sync(0xffffffffu, tile);
        const int tmin = __reduce_min_sync(peers, kmin);
        const int tmax = __reduce_max_sync(peers, kmax);
        const int tbad = (int) __reduce_or_sync(peers, (tid < npix && kstart < 0) ? 1u : 0u);
        const int leader = __ffs(peers) - 1;
        if (tile >= 0 && (int) (threadIdx.x & 31) == leader) {
            if (tmin != INT_MAX) atomicMin(&tiles[tile].kmin, tmin);
            if (tmax != INT_MIN) atomicMax(&tiles[tile].kmax, tmax);
            if (tbad) atomicOr(&tiles[tile].bad, 1);
        }
    }
    unsigned long long pp = (kstart >= 0) ? (unsigned long long) (kstop - kstart) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
        pp += __shfl_xor_sync(0xffffffffu, pp, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (kmin != INT_MAX) atomicMin(&status->kmin, kmin);
        if (kmax != INT_MIN) atomicMax(&status->kmax, kmax);
        if (pp) atomicAdd(&status->pixel_pulses, pp);
    }
}

// ---- generic accumulation (exact kernel evaluation, FP64 throughout) ---------------
// sumCoherent, Backproject.cpp:30-63, with the CPU path's zero-padded edge windows
// (core/detail/Interp1d.h:54-80).  Used for kernels the fast path cannot represent
// (Bartlett/Linear kinks, unsupported tap counts) and as an on-device cross-check.
__global__ void __launch_bounds__(256)
accumulate_generic_kernel(AccumParams P, const PixelRec* __restrict__ pix,
                          const double* __restrict__ pv, const float2* __restrict__ rc,
                          double2* __restrict__ acc)
{
    const long long tid = (long long) blockIdx.x * blockDim.x + threadIdx.x + (long long) P.line_begin * P.out_width;
    const long long tid_end = P.line_end > 0 ? min((long long) P.line_end * P.out_width, P.npix) : P.npix;
    if (tid >= tid_end) return;
    if (P.tile_mask) {
        const int j = (int) (tid / P.out_width), i = (int) (tid % P.out_width);
        if (!P.tile_mask[(j / P.tile_az) * P.tiles_rg + (i / P.tile_rg)].bad) return;
    }
    const PixelRec rec = pix[tid];
    if (rec.kstart < 0) return;
    const int k0 = max(rec.kstart, P.k_begin), k1 = min(rec.kstop, P.k_end);
    if (k1 <= k0) return;
    const D3 x = {rec.x, rec.y, rec.z};
    const double tau_atm = rec.tau_atm;
    const int W = P.kernel.taps;
    double sr = 0., si = 0.;
    for (int k = k0; k < k1; ++k) {
        const D3 p = {pv[6 * k], pv[6 * k + 1], pv[6 * k + 2]};
        const D3 v = {pv[6 * k + 3], pv[6 * k + 4], pv[6 * k + 5]};
        const D3 rr = x - p;
        const double tau = tau_atm + 2. * (dot(rr, v) - kC * norm(rr)) / (dot(v, v) - (kC * kC));
        const double u = (tau - P.swst) / P.dtau;
        const long long i0 = (W % 2 == 0) ? (long long) ceil(u) : (long long) round(u);
        const long long low = i0 - W / 2;
        const float2* line = rc + (size_t) (k - P.rc_k0) * P.rc_pitch;
        float ar = 0.f, ai = 0.f;
        for (int m = 0; m < W; ++m) {
            const long long jj = low + m;
            const float w = kernel_eval(P.kernel, (double) jj - u);
            if (jj >= 0 && jj < P.nr) {
                const float2 d = line[jj];
                ar += w * d.x;
                ai += w * d.y;
            }
        }
        // carrier phase: cycles reduced to [-1/2, 1/2] in FP64, sin / cos of the remainder in
        // FP32 (|error| ~ 1e-7 rad; the FP64 sincospi cost a quarter of this kernel)
        const double cyc = P.fc * tau;
        float sphi, cphi;
        sincospif(2.f * (float) (cyc - rint(cyc)), &sphi, &cphi);
        sr += (double) ar * cphi - (double) ai * sphi;
        si += (double) ar * sphi + (double) ai * cphi;
    }
    double2 a = acc[tid];
    a.x += sr;
    a.y += si;
    acc[tid] = a;
}

// ---- finalisation: complex<double> accumulator -> complex64, NaN for failed pixels ---
// Optional output encoding of the workflow's writer fused in (focus.py:899-925): multiply
// every range column by a complex phasor (deramp / scale), zero the low mantissa bits
// (isce3/core/types.py:116-171).
__global__ void finalize_kernel(long long npix, int out_width, const PixelRec* __restrict__ pix,
                                const double2* __restrict__ acc, float2* __restrict__ out,
                                const float2* __restrict__ range_cor, unsigned mantissa_mask)
{
    const long long tid = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= npix) return;
    float2 o;
    if (pix[tid].kstart < 0) {
        o.x = o.y = nanf("");
    } else {
        const double2 a = acc[tid];
        o.x = (float) a.x;
        o.y = (float) a.y;
        if (range_cor) {
            // complex64 * complex64 in float, products rounded separately like the host's
            const float2 c = range_cor[tid % out_width];
            const float re = __fsub_rn(__fmul_rn(o.x, c.x), __fmul_rn(o.y, c.y));
            const float im = __fadd_rn(__fmul_rn(o.x, c.y), __fmul_rn(o.y, c.x));
            o.x = re;
            o.y = im;
        }
        o.x = __uint_as_float(__float_as_uint(o.x) & mantissa_mask);
        o.y = __uint_as_float(__float_as_uint(o.y) & mantissa_mask);
    }
    out[tid] = o;
}

// ---- batch geometry (the solvers as a reusable device API) ------------------------------
// One thread per point; same device functions as the target solve above.  These replace, for
// array callers, what the reference exposes per thread in cuda/geometry/gpuGeometry.cu:57-68
// (rdr2geo_bracket) and :166-181 (geo2rdr_bracket) on top of device-side `new` + virtual
// dispatch (gpuDEMInterpolator.cu:69-90): here the DEM / LUT / orbit are POD descriptors.
__global__ void __launch_bounds__(128)
rdr2geo_batch_kernel(DevOrbit orbit, DevDEM dem, double wvl, int side, I3B_Rdr2GeoBracketParams prm,
                     long long n, const double* __restrict__ aztime, const double* __restrict__ range,
                     const double* __restrict__ doppler, double* __restrict__ xyz,
                     int* __restrict__ status, DevStatus* dev_status)
{
    const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    D3 x = nan3();
    int st = rdr2geo_bracket(aztime[i], range[i], doppler ? doppler[i] : 0.0, orbit, dem, wvl, side, prm, &x);
    if (st == I3B_EXC_OUT_OF_RANGE) {
        // (the CPU reference throws OutOfRange here, core/Orbit.cpp:78-83; per point: a domain error code)
        st = I3B_ORBIT_INTERP_DOMAIN_ERROR;
        x = nan3();
    } else if (st != I3B_SUCCESS) {
        x = nan3();
    }
    xyz[3 * i] = x.x;
    xyz[3 * i + 1] = x.y;
    xyz[3 * i + 2] = x.z;
    if (status) status[i] = st;
    if (st != I3B_SUCCESS) dev_status->soft_error = st;
}

__global__ void __launch_bounds__(128)
geo2rdr_batch_kernel(DevOrbit orbit, DevLUT2d dop, double wvl, int side, I3B_Geo2RdrBracketParams prm,
                     long long n, const double* __restrict__ xyz, double* __restrict__ aztime,
                     double* __restrict__ range, int* __restrict__ status, DevStatus* dev_status)
{
    const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const D3 x = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    double t = nan(""), r = nan("");
    const int st = geo2rdr_bracket(x, orbit, dop, wvl, side, prm, &t, &r);
    if (st != I3B_SUCCESS) {
        t = r = nan("");
        dev_status->soft_error = st;
    }
    aztime[i] = t;
    range[i] = r;
    if (status) status[i] = st;
}

void launch_rdr2geo_batch(const DevOrbit& orbit, const DevDEM& dem, double wvl, int side,
                          const I3B_Rdr2GeoBracketParams& prm, long long n, const double* aztime,
                          const double* range, const double* doppler, double* xyz, int* status,
                          DevStatus* dev_status, cudaStream_t s)
{
    if (n <= 0) return;
    rdr2geo_batch_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, s>>>(orbit, dem, wvl, side, prm, n, aztime, range,
                                                                     doppler, xyz, status, dev_status);
}

void launch_geo2rdr_batch(const DevOrbit& orbit, const DevLUT2d& dop, double wvl, int side,
                          const I3B_Geo2RdrBracketParams& prm, long long n, const double* xyz,
                          double* aztime, double* range, int* status, DevStatus* dev_status, cudaStream_t s)
{
    if (n <= 0) return;
    geo2rdr_batch_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, s>>>(orbit, dop, wvl, side, prm, n, xyz, aztime, range,
                                                                     status, dev_status);
}

// ---- launchers -------------------------------------------------------------------------

void launch_pulse_table(const DevOrbit& orbit, Linspace in_time, const double* in_times, double fc,
                        PulseRec* pulse, double* pv, DevStatus* status, cudaStream_t s)
{
    const int n = in_time.size + kPulsePadLo + kPulsePadHi;
    pulse_table_kernel<<<(n + 127) / 128, 128, 0, s>>>(orbit, in_time, in_times, fc, pulse, pv, status);
}

void launch_target_solve(const SolveParams& P, PixelRec* pix, float* height, TileInfo* tiles,
                         int n_tiles, DevStatus* status, cudaStream_t s)
{
    tile_info_init_kernel<<<(n_tiles + 255) / 256, 256, 0, s>>>(tiles, n_tiles);
    const long long npix = (long long) P.out_lines * P.out_width;
    if (npix == 0) return;
    const unsigned grid = (unsigned) ((npix + 127) / 128);
    const bool raster = P.dem.have_raster != 0;
    const bool legendre = P.out_orbit.method == I3B_ORBIT_LEGENDRE || P.in_orbit.method == I3B_ORBIT_LEGENDRE;
    const bool lut = P.out_doppler.have_data || P.in_doppler.have_data;
    const int cls = (raster ? 4 : 0) | (legendre ? 2 : 0) | (lut ? 1 : 0);
#define I3B_SOLVE_CASE(C, R, L, U) \
    case C: target_solve_kernel<R, L, U><<<grid, 128, 0, s>>>(P, pix, height, tiles, status); break
    switch (cls) {
        I3B_SOLVE_CASE(0, false, false, false);
        I3B_SOLVE_CASE(1, false, false, true);
        I3B_SOLVE_CASE(2, false, true, false);
        I3B_SOLVE_CASE(3, false, true, true);
        I3B_SOLVE_CASE(4, true, false, false);
        I3B_SOLVE_CASE(5, true, false, true);
        I3B_SOLVE_CASE(6, true, true, false);
        I3B_SOLVE_CASE(7, true, true, true);
    }
#undef I3B_SOLVE_CASE
}

void launch_accumulate_generic(const AccumParams& P, const PixelRec* pix, const double* pv,
                               const float2* rc, double2* acc, cudaStream_t s)
{
    const long long first = (long long) P.line_begin * P.out_width;
    const long long last = P.line_end > 0 ? std::min((long long) P.line_end * P.out_width, P.npix) : P.npix;
    if (last <= first) return;
    const unsigned grid = (unsigned) ((last - first + 255) / 256);
    accumulate_generic_kernel<<<grid, 256, 0, s>>>(P, pix, pv, rc, acc);
}

void launch_finalize(long long npix, int out_width, const PixelRec* pix, const double2* acc, float2* out,
                     const float2* range_cor, int mantissa_nbits, cudaStream_t s)
{
    const unsigned grid = (unsigned) ((npix + 255) / 256);
    const unsigned mask = (mantissa_nbits > 0 && mantissa_nbits < 23) ? (0xFFFFFFFFu << (23 - mantissa_nbits))
                                                                      : 0xFFFFFFFFu;
    finalize_kernel<<<grid, 256, 0, s>>>(npix, out_width, pix, acc, out, range_cor, mask);
}

} // namespace i3b
