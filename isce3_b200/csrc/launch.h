// Host-visible kernel parameter blocks and launchers (internal to the library).
#pragma once
#include "common.cuh"

namespace i3b {

struct SolveParams {
    Linspace out_time, out_range, in_time;
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
    // is flagged (tiles holding failed pixels are skipped by the fast kernel)
    const unsigned char* tile_mask;
    int tile_az, tile_rg, tiles_rg;
};

void launch_pulse_table(const DevOrbit& orbit, Linspace in_time, double fc, PulseRec* pulse,
                        double* pv, DevStatus* status, cudaStream_t s);
void launch_target_solve(const SolveParams& P, PixelRec* pix, float* height, DevStatus* status,
                         cudaStream_t s);
void launch_accumulate_generic(const AccumParams& P, const PixelRec* pix, const double* pv,
                               const float2* rc, double2* acc, cudaStream_t s);
void launch_finalize(long long npix, const PixelRec* pix, const double2* acc, float2* out,
                     cudaStream_t s);

// fast path (accumulate_fast.cu)
// `host_kernel`: same kernel with its data pointer in HOST memory (polynomial fitting).
bool fast_supported(const DevKernel& host_kernel, char* why, size_t why_len);
int launch_accumulate_fast(const AccumParams& P, const DevKernel& host_kernel,
                           const PixelRec* pix, const PulseRec* pulse, const float2* rc,
                           double2* acc, unsigned char* tile_generic, DevStatus* status,
                           cudaStream_t s);
int fast_fit(const DevKernel& host_kernel, I3B_TapPolyFit* fit, char* why, size_t why_len);
int fast_tiles(int out_lines, int out_width);
void fast_tile_shape(int* tile_az, int* tile_rg);
// The pulse table holds records for pulses [-kPulsePadLo, n_pulses + kPulsePadHi): the entries
// outside the input grid are orbit EXTRAPOLATIONS (smooth continuation), used by the fast
// kernel's segment-boundary evaluations and by staged tiles that run over the ends.
constexpr int kPulsePadLo = 128;
constexpr int kPulsePadHi = 288;

int measure_peaks(int device, I3B_Peaks* out);

} // namespace i3b
