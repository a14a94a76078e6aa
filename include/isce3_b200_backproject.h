/* isce3_b200_backproject.h -- C-ABI of the B200-native time-domain backprojection
 * (TDBP) backend.  Plain pointers and sizes only; no C++ / CUDA / torch types.
 *
 * This is the drop-in boundary for ONE path of isce-framework/isce3:
 *
 *   isce3::cuda::focus::backproject(...)   cxx/isce3/cuda/focus/Backproject.h:78-88
 *   isce3::focus::backproject(...)         cxx/isce3/focus/Backproject.h:37-47
 *   isce3.cuda.focus.backproject(...)      python/extensions/pybind_isce3/cuda/focus/Backproject.cpp:25-117
 *
 * A thin C++ adapter (shown in INTEGRATION.md) flattens the isce3 value types
 * (RadarGeometry, Orbit, LUT2d, DEMInterpolator, Kernel<float>, bracket params)
 * into the descriptors below and calls i3b_backproject().  The library never
 * throws across this boundary: every entry point returns an int status and
 * i3b_last_error() returns the message of the last failure on this thread.
 *
 * All arrays are caller-owned, row-major, host memory (pageable is fine).  The
 * callee owns every device allocation and frees it before returning (one-shot
 * call) or at i3b_plan_destroy() / i3b_blocks_destroy() (resident objects).  A
 * caller may opt in to a per-process cache that keeps freed buffers for the
 * next call: i3b_set_device_memory_pool().
 */
#ifndef ISCE3_B200_BACKPROJECT_H
#define ISCE3_B200_BACKPROJECT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define I3B_ABI_VERSION 3

/* ---- status codes --------------------------------------------------------
 * 0..11 mirror isce3::error::ErrorCode (cxx/isce3/error/ErrorCode.h:8-21): they
 * are "soft" per-pixel geometry outcomes -- the call completed, failed pixels
 * are (NaN,NaN) in `out` and NaN in `height` (Backproject.cpp:146-153,167-171).
 * Negative codes stand for the exceptions the reference throws
 * (cxx/isce3/except/Error.h:20-38, cuda/except/Error.h:33-66); the adapter
 * rethrows them as the matching isce3::except type.                        */
enum {
    I3B_SUCCESS = 0,
    I3B_ORBIT_INTERP_SIZE_ERROR = 1,
    I3B_ORBIT_INTERP_DOMAIN_ERROR = 2,
    I3B_ORBIT_INTERP_UNKNOWN_METHOD = 3,
    I3B_OUT_OF_BOUNDS_DEM = 4,
    I3B_INVALID_DEM = 5,
    I3B_FAILED_TO_CONVERGE = 6,
    I3B_WRONG_LOOK_SIDE = 7,
    I3B_OUT_OF_BOUNDS_LOOKUP = 8,
    I3B_NULL_DEREFERENCE = 9,
    I3B_INVALID_TOLERANCE = 10,
    I3B_INVALID_INTERVAL = 11,

    I3B_EXC_INVALID_ARGUMENT = -1, /* isce3::except::InvalidArgument        */
    I3B_EXC_RUNTIME_ERROR = -2,    /* isce3::except::RuntimeError (epochs,
                                      unsupported kernel)                   */
    I3B_EXC_DOMAIN_ERROR = -3,     /* isce3::except::DomainError (batch<1)  */
    I3B_EXC_OVERFLOW_ERROR = -4,   /* isce3::except::OverflowError          */
    I3B_EXC_OUT_OF_RANGE = -5,     /* isce3::except::OutOfRange (orbit)     */
    I3B_EXC_CUDA_ERROR = -6,       /* isce3::cuda::except::CudaError        */
    I3B_EXC_NO_DEVICE = -7,        /* no usable sm_100 device: there is NO
                                      CPU fallback in this library          */
    I3B_EXC_LENGTH_ERROR = -8      /* isce3::except::LengthError (RangeComp
                                      batch > maxbatch)                     */
};

/* isce3::core::LookSide (cxx/isce3/core/LookSide.h:13-17) */
enum { I3B_LOOK_LEFT = 1, I3B_LOOK_RIGHT = -1 };

/* isce3::core::OrbitInterpMethod (cxx/isce3/core/Orbit.h) */
enum { I3B_ORBIT_HERMITE = 0, I3B_ORBIT_LEGENDRE = 1 };

/* isce3::core::dataInterpMethod (cxx/isce3/core/Constants.h) subset usable by
 * LUT2d / DEMInterpolator on this path.                                    */
enum {
    I3B_INTERP_SINC = 0, /* Sinc2dInterpolator(SINC_LEN = 8, SINC_SUB = 8192) */
    I3B_INTERP_BILINEAR = 1,
    I3B_INTERP_BICUBIC = 2,
    I3B_INTERP_NEAREST = 3,
    I3B_INTERP_BIQUINTIC = 4
};

/* isce3::focus::DryTroposphereModel (cxx/isce3/focus/DryTroposphereModel.h:15-21) */
enum { I3B_TROPO_NODELAY = 0, I3B_TROPO_TSX = 1 };

/* isce3::core::Kernel<float> dynamic types accepted by the reference CUDA path
 * (cuda/focus/Backproject.cu:715-752).                                     */
enum {
    I3B_KERNEL_BARTLETT = 0,
    I3B_KERNEL_LINEAR = 1,
    I3B_KERNEL_KNAB = 2,
    I3B_KERNEL_TABULATED = 3,
    I3B_KERNEL_CHEBY = 4
};

/* isce3::product::RadarGridParameters (product/RadarGridParameters.h:43-163),
 * already re-based to the orbit reference epoch as RadarGeometry's ctor does
 * (container/RadarGeometry.icc:7-26).                                      */
typedef struct {
    double sensing_start;       /* s since ref epoch: t of line 0            */
    double prf;                 /* Hz; line spacing is 1/prf                 */
    double starting_range;      /* m                                         */
    double range_pixel_spacing; /* m                                         */
    double wavelength;          /* m (informational; the path uses c/fc)     */
    int64_t length;             /* azimuth lines                             */
    int64_t width;              /* range samples                             */
    int32_t look_side;          /* I3B_LOOK_*                                */
    int32_t _pad;
} I3B_RadarGrid;

/* isce3::core::Orbit (core/Orbit.h:193-199): uniformly sampled state vectors */
typedef struct {
    double t0;         /* s since ref epoch of state vector 0              */
    double dt;         /* s                                                */
    int32_t n;         /* number of state vectors                          */
    int32_t method;    /* I3B_ORBIT_*                                      */
    const double* pos; /* [n][3] ECEF m                                    */
    const double* vel; /* [n][3] ECEF m/s                                  */
} I3B_Orbit;

/* isce3::core::LUT2d<double> (core/LUT2d.h:54-95, core/LUT2d.cpp:127-160).
 * have_data == 0 reproduces the default-constructed LUT (eval == ref_value).*/
typedef struct {
    int32_t have_data;
    int32_t bounds_error;
    int32_t method; /* I3B_INTERP_* */
    int32_t _pad;
    int64_t length; /* rows  (y = azimuth time) */
    int64_t width;  /* cols  (x = slant range)  */
    double ref_value;
    double xstart, ystart, dx, dy;
    const double* data; /* [length][width] */
} I3B_LUT2d;

/* isce3::container::RadarGeometry (container/RadarGeometry.h:16-64) */
typedef struct {
    I3B_RadarGrid grid;
    I3B_Orbit orbit;
    I3B_LUT2d doppler;
    /* orbit reference epoch, only compared for equality between the input
     * and output geometry (Backproject.cpp:88-92): whole seconds since
     * 1970-01-01T00:00:00 and the fractional second.                       */
    int64_t ref_epoch_sec;
    double ref_epoch_frac;
} I3B_RadarGeometry;

/* isce3::geometry::DEMInterpolator (geometry/DEMInterpolator.h:31-52,
 * DEMInterpolator.cpp:592-659).  have_raster == 0 -> constant ref_height.  */
typedef struct {
    int32_t have_raster;
    int32_t epsg;   /* 4326 supported for rasters; any WGS84-based code for
                       constant-height DEMs                                 */
    int32_t method; /* I3B_INTERP_* */
    int32_t _pad;
    int64_t length, width;
    double ref_height;
    double xstart, ystart, dx, dy; /* pixel-centre coordinates of data[0][0] */
    const float* data;             /* [length][width] */
} I3B_DEM;

/* isce3::core::Kernel<float> (core/Kernels.h:18-172) */
typedef struct {
    int32_t kind;     /* I3B_KERNEL_* */
    int32_t n;        /* table length (TABULATED) or #coeffs (CHEBY)        */
    double width;     /* kernel.width() = 2*halfwidth                       */
    double bandwidth; /* KNAB only                                          */
    const float* data; /* table[n] (TABULATED, samples on [0,halfwidth]) or
                          Chebyshev coefficients[n] (CHEBY); else NULL      */
} I3B_Kernel;

/* isce3::geometry::detail::Rdr2GeoBracketParams (geometry/detail/Rdr2Geo.h:88-97) */
typedef struct {
    double tol_height; /* default 1e-5 m   */
    double look_min;   /* default 0        */
    double look_max;   /* default pi/2     */
} I3B_Rdr2GeoBracketParams;

/* isce3::geometry::detail::Geo2RdrBracketParams (geometry/detail/Geo2Rdr.h:57-68) */
typedef struct {
    double tol_aztime; /* default 1e-7 s */
    int32_t has_time_start, has_time_end;
    double time_start, time_end;
} I3B_Geo2RdrBracketParams;

/* Everything isce3::cuda::focus::backproject takes, flattened.             */
typedef struct {
    uint32_t abi_version; /* = I3B_ABI_VERSION */
    uint32_t flags;       /* I3B_FLAG_* */

    /* Host arrays (pageable or page-locked; page-locked ones make every copy asynchronous:
     * the upload then hides behind the target solve and the accumulation, the results leave
     * the device while the last rows are still being summed), or device arrays with
     * I3B_FLAG_DEVICE_POINTERS / I3B_FLAG_DEVICE_INPUT.                                   */
    float* out;      /* complex64 [out.length][out.width] interleaved re,im  */
    const float* in; /* complex64 [in.length][in.width] range-compressed     */
    float* height;   /* float32 [out.length][out.width] or NULL              */

    I3B_RadarGeometry out_geometry;
    I3B_RadarGeometry in_geometry;
    I3B_DEM dem;
    double fc; /* centre frequency, Hz */
    double ds; /* desired azimuth resolution, m */
    I3B_Kernel kernel;
    int32_t dry_tropo_model; /* I3B_TROPO_* */
    int32_t batch;           /* pulses per H2D slab (>= 1), reference default 1024 */
    I3B_Rdr2GeoBracketParams rdr2geo;
    I3B_Geo2RdrBracketParams geo2rdr;

    /* multi-GPU extension: the output grid is cut into contiguous azimuth
     * blocks, one per listed device, each fed only the pulses its block's
     * apertures cover; no inter-GPU exchange.  n_devices == 0 -> the
     * current device only (reference behaviour, focus.py:1589-1595).       */
    int32_t n_devices;
    const int32_t* devices;

    /* output-encoding extension (ABI 2): what the workflow's writer does to every focused
     * block on the host before storing it (nisar/workflows/focus.py:899-925), fused into
     * the kernel that converts the accumulator to complex64.  Both optional.            */
    const float* range_cor;  /* complex64 [out.width] or NULL: out[j][i] *= range_cor[i]
                                (range deramp / scale phasors, focus.py:899-900)       */
    int32_t mantissa_nbits;  /* 0: keep all 23 mantissa bits; 1..23: zero the least
                                significant ones like truncate_mantissa(z, n)
                                (python/packages/isce3/core/types.py:116-171)          */
    int32_t _pad2;

    /* non-uniform pulse timing extension (ABI 3).  The reference API can only describe
     * uniformly spaced pulses (RadarGeometry::sensingTime() is a Linspace,
     * container/RadarGeometry.icc:28-42), so the workflow first RESAMPLES dithered-PRF raw
     * data onto a uniform grid (nisar/workflows/focus.py:973-1061) -- a full pass over the
     * swath that time-domain backprojection does not need: it only wants to know where the
     * platform was at each pulse.
     *   pulse_times: azimuth time (s since the reference epoch) of every input line
     *   [in_geometry.grid.length], strictly increasing; NULL: the uniform grid
     *   sensing_start + k / prf.  The platform state of pulse k is the input orbit at
     *   pulse_times[k]; a pixel integrates the pulses from the last one at or before the start
     *   of its coherent processing interval up to (not including) the first one at or after
     *   its end -- for uniform times exactly the floor / ceil of Backproject.cpp:186-193.
     *   in_geometry.grid.prf stays the nominal PRF.                                       */
    const double* pulse_times;
} I3B_BackprojectArgs;

enum {
    I3B_FLAG_NONE = 0,
    /* force the generic accumulation kernel (evaluates the interpolation
     * kernel exactly as core/Kernels.icc does) instead of the fast one     */
    I3B_FLAG_FORCE_GENERIC = 1u << 0,
    /* `in` / `out` / `height` are DEVICE pointers on the current device
     * (no staging copies); single device only                              */
    I3B_FLAG_DEVICE_POINTERS = 1u << 1,
    /* only `in` is a DEVICE pointer on the current device (e.g. the output of
     * i3b_rangecomp_execute_to_device): the swath never crosses the host link;
     * `out` / `height` stay host arrays; single device only                 */
    I3B_FLAG_DEVICE_INPUT = 1u << 2
};

/* Per-call measurements, filled by i3b_last_stats() for the last successful
 * i3b_backproject()/i3b_plan_execute() on this thread.  Durations are CUDA
 * event times on the launching stream.                                     */
typedef struct {
    double pixel_pulses;      /* sum over pixels of (kstop - kstart)        */
    double ms_total;          /* wall time of the whole call                */
    double ms_h2d;            /* host->device staging of `in`               */
    double ms_target_solve;   /* per-pixel rdr2geo/geo2rdr kernel           */
    double ms_accumulate;     /* sum of accumulation-kernel launches        */
    double ms_d2h;            /* device->host of out/height                 */
    int32_t accumulate_launches;
    int32_t total_launches;   /* all kernels of this library in the call    */
    int32_t used_fast_kernel; /* 1 = fast path, 0 = generic                 */
    int32_t taps;             /* ceil(kernel.width)                         */
    int64_t h2d_bytes, d2h_bytes;
    int32_t pulse_first, pulse_last; /* [first,last) pulses actually used   */
    int32_t n_devices;
    int32_t fast_variant;     /* fast path: >= 0 build-time specialised kernel
                                 (baked coefficient table), -1 the general
                                 constant-bank kernel                       */
} I3B_Stats;

/* Device microbenchmarks used as roofline denominators (bench harness).    */
typedef struct {
    double fp32_tflops; /* FFMA throughput, 2 flop per FMA                  */
    double fp64_tflops; /* DFMA throughput                                  */
    double sfu_gops;    /* MUFU.SIN/COS ops per ns                          */
    double sm_mhz;      /* SM clock implied by a clock64()/event pair       */
    int32_t sm_count;
    int32_t _pad;
} I3B_Peaks;

/* Per-tap polynomial form of the interpolation kernel used by the fast
 * accumulation kernel (host-side fit, no device needed): for tap m of K and
 * the centred fractional sample offset f in [-1/2, 1/2),
 *   w_m(f)     = E_m(f^2) + f*O_m(f^2),      m < ceil(K/2)
 *   w_{K-1-m}  = E_m(f^2) - f*O_m(f^2)
 * even[m][i] multiplies f^(2i), odd[m][i] multiplies f^(2i+1).  max_err is the
 * largest |polynomial - kernel(x)| over the taps (kernel evaluated exactly as
 * Kernel<float>::operator(), core/Kernels.icc); `supported` is 0 when the fast
 * kernel cannot take this kernel (reason in i3b_last_error()).             */
#define I3B_FIT_MAX_TAP_PAIRS 17
#define I3B_FIT_MAX_COEF 4
typedef struct {
    int32_t taps, degree;
    int32_t supported;
    int32_t imm_variant; /* >= 0: index of the build-time specialised kernel
                            whose baked coefficients equal this fit         */
    double max_err;
    float even[I3B_FIT_MAX_TAP_PAIRS][I3B_FIT_MAX_COEF];
    float odd[I3B_FIT_MAX_TAP_PAIRS][I3B_FIT_MAX_COEF];
    int32_t pair_degree[I3B_FIT_MAX_TAP_PAIRS]; /* degree used for tap pair m (<= degree):
                            outer pairs carry small weights and need fewer terms       */
    int32_t _pad;
} I3B_TapPolyFit;

/* ---- entry points -------------------------------------------------------- */

/* Blocking one-shot call; replaces isce3::cuda::focus::backproject
 * (cuda/focus/Backproject.h:78-88).  Returns an I3B_* status.              */
int i3b_backproject(const I3B_BackprojectArgs* args);

/* Resident variant for callers that keep the range-compressed swath in HBM
 * (bench `value`; a future RangeComp-on-GPU caller): create uploads `in` and
 * the geometry, execute runs target solve + accumulation on device, download
 * copies out/height back.                                                  */
typedef struct I3B_Plan I3B_Plan;
int i3b_plan_create(const I3B_BackprojectArgs* args, I3B_Plan** plan);
int i3b_plan_execute(I3B_Plan* plan);
int i3b_plan_download(I3B_Plan* plan, float* out, float* height);
int i3b_plan_destroy(I3B_Plan* plan);

/* Blocks API: ONE swath, MANY output blocks.  The workflow cuts the image
 * into blocks and calls backproject once per block with the whole swath pointer
 * (nisar/workflows/focus.py:726-783, :1988-2007); i3b_blocks_create uploads the swath, DEM,
 * LUTs and per-pulse tables ONCE to every listed device, i3b_blocks_run focuses any number of
 * blocks from there -- free devices take the next block -- writing each into its own host
 * array, exactly what the per-block calls would have produced.
 *   args: as for i3b_backproject; `out` / `height` are ignored, out_geometry supplies the
 *         orbit and Doppler LUT of the output grids (its grid is the whole image: range_cor,
 *         if given, holds that image's columns).
 *   grids[i]: output sub-grid of block i (sensing_start / starting_range already those of
 *         the block, as RadarGridParameters slicing gives them); out[i]: complex64
 *         [length][width]; height: NULL or height[i] float32 [length][width] (entries may be
 *         NULL).                                                                        */
typedef struct I3B_Blocks I3B_Blocks;
int i3b_blocks_create(const I3B_BackprojectArgs* args, I3B_Blocks** blocks);
int i3b_blocks_run(I3B_Blocks* blocks, int32_t n, const I3B_RadarGrid* grids, float* const* out,
                   float* const* height);
int i3b_blocks_destroy(I3B_Blocks* blocks);

int i3b_last_stats(I3B_Stats* stats);
const char* i3b_last_error(void);
const char* i3b_version(void);
int i3b_device_count(void);
/* The calling thread's current CUDA device (-1: none).  Every entry point of this library
 * leaves it as it found it (the reference never changes it either: the caller selects the
 * device once, focus.py:1592-1593).                                          */
int i3b_current_device(void);
int i3b_measure_peaks(int device, I3B_Peaks* peaks);
/* Device memory between calls.  Default: none is kept -- every device allocation of a call is
 * back with the driver when it returns (the reference's contract).  keep_mb < 0: keep freed
 * buffers in a per-process cache for the next call, without limit (a workflow focusing block
 * after block; avoids ~6 GB of cudaMalloc / cudaFree per call); keep_mb > 0: keep at most that
 * many MiB; 0: back to the default (and release what is cached).  Environment variable
 * I3B_POOL_KEEP_MB sets the initial value.                                           */
int i3b_set_device_memory_pool(int64_t keep_mb);
/* Return device memory cached by earlier calls to the driver (all devices). */
int i3b_release_device_memory(void);
/* Host-only diagnostic: the polynomial fit the fast kernel would use.      */
int i3b_fit_tap_polynomials(const I3B_Kernel* kernel, I3B_TapPolyFit* fit);

/* ---- per-point geometry as a batch API ---------------------------
 * The two bracketing solvers of the path for ARRAYS of points, on the current device: what the
 * other GPU consumers of the reference (cuda/geometry/gpuTopo.cu:185, gpuGeo2rdr.cu:28,
 * cuda/geocode/Geocode.cu:87) call per thread through cuda/geometry/gpuGeometry.cu:57-68 and
 * :166-181.  Same semantics as isce3::geometry::rdr2geo_bracket (geometry/rdr2geo_roots.cpp:14-27)
 * and geo2rdr_bracket (geometry/geo2rdr_roots.cpp:16-25) per point; points that fail are NaN.
 * Arrays are host arrays (device arrays on the current device with I3B_FLAG_DEVICE_POINTERS).
 * `status` (optional) receives the per-point ErrorCode; the call returns the last non-success
 * ErrorCode of any point (0 if all converged) or a negative I3B_EXC_* code.              */
int i3b_rdr2geo_bracket_batch(const I3B_Orbit* orbit, const I3B_DEM* dem, double wavelength,
                              int32_t look_side, const I3B_Rdr2GeoBracketParams* params, int64_t n,
                              const double* aztime, const double* slant_range,
                              const double* doppler /* [n] or NULL = zero Doppler */,
                              double* xyz /* [n][3] ECEF m */, int32_t* status /* [n] or NULL */,
                              uint32_t flags);
int i3b_geo2rdr_bracket_batch(const I3B_Orbit* orbit, const I3B_LUT2d* doppler, double wavelength,
                              int32_t look_side, const I3B_Geo2RdrBracketParams* params, int64_t n,
                              const double* xyz /* [n][3] ECEF m */, double* aztime, double* slant_range,
                              int32_t* status /* [n] or NULL */, uint32_t flags);

/* ---- range compression (the step that produces `in`) ---------
 * isce3::focus::RangeComp (cxx/isce3/focus/RangeComp.h:13-116, RangeComp.cpp):
 * frequency-domain convolution of each input line with the time-reversed complex
 * conjugate of the chirp.  The reference has no CUDA twin of this class; the FFTs
 * here are cuFFT.                                                              */
enum { I3B_RANGECOMP_FULL = 0, I3B_RANGECOMP_VALID = 1, I3B_RANGECOMP_SAME = 2 }; /* RangeComp::Mode */
typedef struct I3B_RangeComp I3B_RangeComp;
/* RangeComp::RangeComp(chirp, inputsize, maxbatch, mode); chirp = complex64[chirp_size] */
int i3b_rangecomp_create(const float* chirp, int chirp_size, int input_size, int max_batch,
                         int mode, I3B_RangeComp** rc);
/* fftSize(), outputSize(), firstValidSample() */
int i3b_rangecomp_query(const I3B_RangeComp* rc, int* fft_size, int* output_size,
                        int* first_valid_sample);
/* RangeComp::rangecompress(out, in, batch): in complex64 [batch][input_size] ->
 * out complex64 [batch][output_size]; host pointers, or device pointers on the
 * current device with I3B_FLAG_DEVICE_POINTERS (output stays in HBM for a
 * following i3b_backproject with the same flag).                              */
int i3b_rangecomp_execute(I3B_RangeComp* rc, float* out, const float* in, int batch,
                          uint32_t flags);
/* Radiometric corrections the workflow applies to every range-compressed block on the host
 * (nisar/workflows/focus.py:1956-1975), fused into the pass that writes the output:
 *   out[b][j] = rc[b][j] * column_scale[j] / interp(slant_ranges[j]; pattern_ranges, patterns[b])
 * column_scale: the per-column factors multiplied together -- baseband shift phasors
 * `deramp_rc` and range-loss `slant_ranges / ref_range` -- complex64 [output_size] or NULL;
 * the dynamic antenna pattern of each line, given on the coarse axis pattern_ranges, is
 * interpolated like numpy.interp (linear, clamped) onto slant_ranges.  Set once with
 * i3b_rangecomp_set_scaling (NULL clears); the per-line pattern samples travel with each
 * execute call (complex64 [batch][n_pattern], NULL: no pattern division for that call).    */
typedef struct {
    const float* column_scale;    /* complex64 [output_size] or NULL                      */
    const double* slant_ranges;   /* [output_size]: slant range of every output sample    */
    const double* pattern_ranges; /* [n_pattern], increasing                              */
    int32_t n_pattern;            /* 0: no antenna-pattern division                       */
    int32_t _pad;
} I3B_RangeCompScaling;
int i3b_rangecomp_set_scaling(I3B_RangeComp* rc, const I3B_RangeCompScaling* scaling);
int i3b_rangecomp_execute_scaled(I3B_RangeComp* rc, float* out, const float* in, int batch,
                                 uint32_t flags, const float* patterns);
int i3b_rangecomp_execute_to_device_scaled(I3B_RangeComp* rc, const float* in, int64_t lines,
                                           const float* patterns, float** dev_out);
/* Range-compress `lines` host lines (any number: processed in chunks of maxbatch) into ONE
 * newly allocated device array complex64 [lines][output_size], returned in *dev_out and owned
 * by the caller (i3b_device_free); feed it to i3b_backproject with I3B_FLAG_DEVICE_INPUT.   */
int i3b_rangecomp_execute_to_device(I3B_RangeComp* rc, const float* in, int64_t lines,
                                    float** dev_out);
int i3b_device_free(void* device_pointer);
int i3b_device_to_host(void* dst, const void* device_src, size_t bytes);
double i3b_rangecomp_last_device_ms(const I3B_RangeComp* rc);
const char* i3b_rangecomp_last_error(void);
int i3b_rangecomp_destroy(I3B_RangeComp* rc);

#ifdef __cplusplus
}
#endif
#endif /* ISCE3_B200_BACKPROJECT_H */
