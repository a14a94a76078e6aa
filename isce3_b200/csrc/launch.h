// Host-visible kernel parameter blocks and launchers (internal to the library).
#pragma once
#include "common.cuh"

namespace i3b {

// Per fast-kernel tile summary written by the target solve: pulse span of the tile's pixels
// and whether it holds a failed pixel (those tiles go to the generic kernel).  Lets an
// accumulation launch skip tiles with no pulse in its range before touching pixel records.
struct TileInfo {
    int kmin, kmax; // over pixels with a non-empty aperture; kmin >= kmax: nothing to integrate
    int bad;        // a pixel of the tile failed its geometry solve
    int pad;
};

struct SolveParams {
    Linspace out_time, out_range, in_time;
    // explicit pulse times (non-uniform PRF): device array indexed like the pulse table, from
    // -kPulsePadLo to n_pulses + kPulsePadHi (extrapolated past both ends); null: in_time
    const double* in_times;
    DevOrbit out_orbit, in_orbit;
    DevLUT2d out_doppler, in_doppler;
    DevDEM dem;
    I3B_Rdr2GeoBracketParams r2g;
    I3B_Geo2RdrBracketParams g2r;
    double wvl, ds;
    int out_side, in_side, tropo;
    int line0;     // first output line of this shard
    int out_lines; // lines in this shard
    int out_width;
    int tile_az, tile_rg, tiles_rg; // fast-kernel tile grid of the shard
};

struct AccumParams {
    long long npix;
    int out_lines, out_width; // of the shard
    int nr;                   // range samples per input line
    int rc_pitch;             // float2 elements per staged line
    int rc_k0;                // pulse index of staged line 0
    int rc_rows;              // staged lines
    int n_pulses;             // pulses in the input grid (length of the pulse tables)
    int k_begin, k_end;       // pulses to integrate in this launch
    double fc, swst, dtau;
    double spacing_ratio;     // output / input range pixel spacing
    DevKernel kernel;         // data pointer: device memory
    // generic kernel only: when non-null, process just the pixels whose fast-kernel tile
    // is flagged bad (tiles holding failed pixels are skipped by the fast kernel)
    const TileInfo* tile_mask;
    int tile_az, tile_rg, tiles_rg;
    // output lines [line_begin, line_end) of the shard are processed by this launch
    // (line_end == 0: all); line_begin is a multiple of tile_az
    int line_begin, line_end;
    // pulses < k_landed are on the device (row-wavefront launches of the one-shot call, which
    // integrate whole apertures of the rows whose pulses have landed): a tile that needs more
    // sets DevStatus::premature and is left alone.  0: everything in [k_begin, k_end) is there.
    int k_landed;
    // non-uniform pulse times (fast kernel; null: uniform): tn[k] = t_k * nominal PRF, xi[k] =
    // (float) (tn[k] - tn[segment base of k]); device arrays indexed like the pulse table
    const double* tn;
    const float* xi;
    int seg; // pulses per geometry segment of the fast kernel (fast_segment())
};

void launch_pulse_table(const DevOrbit& orbit, Linspace in_time, const double* in_times, double fc,
                        PulseRec* pulse, double* pv, DevStatus* status, cudaStream_t s);
void launch_target_solve(const SolveParams& P, PixelRec* pix, float* height, TileInfo* tiles,
                         int n_tiles, DevStatus* status, cudaStream_t s);
void launch_accumulate_generic(const AccumParams& P, const PixelRec* pix, const double* pv,
                               const float2* rc, double2* acc, cudaStream_t s);
void launch_finalize(long long npix, int out_width, const PixelRec* pix, const double2* acc, float2* out,
                     const float2* range_cor, int mantissa_nbits, cudaStream_t s);

// batch geometry (solve_kernels.cu): all pointers are device pointers
void launch_rdr2geo_batch(const DevOrbit& orbit, const DevDEM& dem, double wvl, int side,
                          const I3B_Rdr2GeoBracketParams& prm, long long n, const double* aztime,
                          const double* range, const double* doppler, double* xyz, int* status,
                          DevStatus* dev_status, cudaStream_t s);
void launch_geo2rdr_batch(const DevOrbit& orbit, const DevLUT2d& dop, double wvl, int side,
                          const I3B_Geo2RdrBracketParams& prm, long long n, const double* xyz,
                          double* aztime, double* range, int* status, DevStatus* dev_status, cudaStream_t s);

// fast path (accumulate_fast.cu)
// `host_kernel`: same kernel with its data pointer in HOST memory (polynomial fitting).
bool fast_supported(const DevKernel& host_kernel, char* why, size_t why_len);
int launch_accumulate_fast(const AccumParams& P, const DevKernel& host_kernel,
                           const PixelRec* pix, const PulseRec* pulse, const float2* rc,
                           double2* acc, const TileInfo* tiles, DevStatus* status,
                           cudaStream_t s);
int fast_fit(const DevKernel& host_kernel, I3B_TapPolyFit* fit, char* why, size_t why_len);
int fast_tiles(int out_lines, int out_width);
void fast_tile_shape(int* tile_az, int* tile_rg);
// Pulses per staged tile.  Tiles sit on absolute multiples of this; an accumulation launch
// that is not the last one of a call must END on such a multiple, so that every FP32 tile sum
// holds the same pulses whatever the launch partition (bit-reproducible output).
int fast_pulse_tile();
// Pulses per geometry segment of the fast kernel for a scene (64 or 128; segments sit on
// absolute multiples of it): radar wavelength, PRF, fastest platform speed, nearest range.
int fast_segment(double wavelength, double prf, double v_max, double r_min);
// The pulse table holds records for pulses [-kPulsePadLo, n_pulses + kPulsePadHi): the entries
// outside the input grid are orbit EXTRAPOLATIONS (smooth continuation), used by the fast
// kernel's segment-boundary evaluations and by staged tiles that run over the ends.
constexpr int kPulsePadLo = 128;
constexpr int kPulsePadHi = 288;

int measure_peaks(int device, I3B_Peaks* out);

} // namespace i3b
