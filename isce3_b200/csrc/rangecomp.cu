// Range compression on the GPU: the step that produces the backprojection's input
// (SURVEY.md 8f rank 1).  Behavioural reference: isce3::focus::RangeComp,
// cxx/isce3/focus/RangeComp.cpp:10-132 (matched filter = time-reversed conjugate chirp,
// zero-pad to nextFastPower(chirp + input - 1), FFT, multiply, inverse FFT, crop by mode),
// nextFastPower cxx/isce3/fft/FFTUtil.icc:39-75.  The FFTs are cuFFT (library code, batched
// C2C); padding, spectrum multiply (with the 1/N scale folded into the reference spectrum)
// and cropping are the kernels below.  All of it is HBM-bound streaming.
#include <cuda_runtime.h>
#include <cufft.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/isce3_b200_backproject.h"

namespace i3b {

struct RcError : std::runtime_error {
    int code;
    RcError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define RC_CK(call)                                                                            \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            throw RcError(e__ == cudaErrorNoDevice ? I3B_EXC_NO_DEVICE : I3B_EXC_CUDA_ERROR,   \
                          std::string(#call) + " failed: " + cudaGetErrorString(e__));         \
    } while (0)
#define RC_FFT(call)                                                                           \
    do {                                                                                       \
        cufftResult r__ = (call);                                                              \
        if (r__ != CUFFT_SUCCESS)                                                              \
            throw RcError(I3B_EXC_CUDA_ERROR, std::string(#call) + " failed: cufft status " +  \
                                                      std::to_string((int) r__));              \
    } while (0)

// smallest m = 2^a 3^b 5^c >= n  (FFTUtil.icc:39-75)
static int next_fast_power(int n)
{
    if (n <= 1) return 1;
    const double logn = std::log((double) n);
    const int max5 = (int) std::ceil(logn / std::log(5.0)), max3 = (int) std::ceil(logn / std::log(3.0));
    long long best = INT32_MAX;
    long long n5 = 1;
    for (int x5 = 0; x5 <= max5; ++x5, n5 *= 5) {
        long long n3 = 1;
        for (int x3 = 0; x3 <= max3; ++x3, n3 *= 3) {
            long long m = n5 * n3;
            while (m < n) m *= 2;
            best = std::min(best, m);
        }
    }
    return (int) best;
}

static int output_size(int m, int n, int mode)
{
    switch (mode) {
    case I3B_RANGECOMP_FULL: return m + n - 1;
    case I3B_RANGECOMP_VALID: return std::max(m, n) - std::min(m, n) + 1;
    default: return n;
    }
}

// work[b][i] = in[b][i] for i < n_in, 0 up to nfft  (RangeComp.cpp:98-106)
__global__ void rc_pad_kernel(float2* __restrict__ work, const float2* __restrict__ in, int n_in, int nfft,
                              long long total)
{
    for (long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long) gridDim.x * blockDim.x) {
        const long long b = t / nfft;
        const int i = (int) (t - b * nfft);
        work[t] = i < n_in ? in[b * n_in + i] : make_float2(0.f, 0.f);
    }
}

// work[b][i] *= ref[i]   (ref already carries 1/nfft; RangeComp.cpp:109-116)
__global__ void rc_multiply_kernel(float2* __restrict__ work, const float2* __restrict__ ref, int nfft,
                                   long long total)
{
    for (long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long) gridDim.x * blockDim.x) {
        const float2 r = ref[(int) (t % nfft)], w = work[t];
        work[t] = make_float2(w.x * r.x - w.y * r.y, w.x * r.y + w.y * r.x);
    }
}

// out[b][j] = work[b][offset + j]   (RangeComp.cpp:120-129)
__global__ void rc_crop_kernel(float2* __restrict__ out, const float2* __restrict__ work, int n_out, int nfft,
                               int offset, long long total)
{
    for (long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long) gridDim.x * blockDim.x) {
        const long long b = t / n_out;
        const int j = (int) (t - b * n_out);
        out[t] = work[b * nfft + offset + j];
    }
}

// out[b][j] = work[b][offset + j] * col[j] / interp(pattern[b])[j]: the crop with the
// workflow's radiometric corrections fused in (nisar/workflows/focus.py:1956-1975: baseband
// shift `*= deramp_rc[None, :]`, dynamic antenna pattern `/= np.interp(slant_ranges, pat_ranges,
// patterns[pulse])` per line, range-loss `*= slant_ranges / ref_range`).  col (optional) is the
// product of the per-column factors; pat_idx / pat_w (optional) are np.interp's interval and
// weight of every output column on the pattern's range axis, pattern[b] the line's complex samples.
__global__ void rc_crop_scale_kernel(float2* __restrict__ out, const float2* __restrict__ work, int n_out,
                                     int nfft, int offset, long long total, const float2* __restrict__ col,
                                     const int* __restrict__ pat_idx, const float* __restrict__ pat_w,
                                     const float2* __restrict__ pattern, int n_pat)
{
    for (long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long) gridDim.x * blockDim.x) {
        const long long b = t / n_out;
        const int j = (int) (t - b * n_out);
        float2 v = work[b * nfft + offset + j];
        if (col) {
            const float2 c = col[j];
            v = make_float2(v.x * c.x - v.y * c.y, v.x * c.y + v.y * c.x);
        }
        if (pattern) {
            const int i = pat_idx[j];
            const float w = pat_w[j];
            const float2 p0 = pattern[b * n_pat + i], p1 = pattern[b * n_pat + min(i + 1, n_pat - 1)];
            const float px = p0.x + w * (p1.x - p0.x), py = p0.y + w * (p1.y - p0.y);
            const float d = px * px + py * py;
            v = make_float2((v.x * px + v.y * py) / d, (v.y * px - v.x * py) / d);
        }
        out[t] = v;
    }
}

} // namespace i3b

using namespace i3b;

struct I3B_RangeComp {
    int chirp_size = 0, input_size = 0, fft_size = 0, max_batch = 0, mode = 0, out_size = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    float2 *d_ref = nullptr, *d_work = nullptr, *d_in = nullptr, *d_out = nullptr;
    std::map<int, cufftHandle> plans; // by batch size
    double ms_last = 0.0;
    // fused radiometric corrections (i3b_rangecomp_set_scaling)
    float2* d_col = nullptr;
    int* d_pat_idx = nullptr;
    float* d_pat_w = nullptr;
    float2* d_pattern = nullptr; // [max_batch][n_pat]
    int n_pat = 0;

    cufftHandle plan_for(int batch)
    {
        auto it = plans.find(batch);
        if (it != plans.end()) return it->second;
        cufftHandle h;
        int n[1] = {fft_size};
        RC_FFT(cufftPlanMany(&h, 1, n, nullptr, 1, fft_size, nullptr, 1, fft_size, CUFFT_C2C, batch));
        RC_FFT(cufftSetStream(h, stream));
        plans[batch] = h;
        return h;
    }
    ~I3B_RangeComp()
    {
        cudaSetDevice(device);
        for (auto& kv : plans) cufftDestroy(kv.second);
        cudaFree(d_ref);
        cudaFree(d_work);
        cudaFree(d_in);
        cudaFree(d_out);
        cudaFree(d_col);
        cudaFree(d_pat_idx);
        cudaFree(d_pat_w);
        cudaFree(d_pattern);
        if (stream) cudaStreamDestroy(stream);
    }
};

static thread_local std::string g_rc_error;

template<class F>
static int rc_guarded(F&& f)
{
    try {
        return f();
    } catch (const RcError& e) {
        g_rc_error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        g_rc_error = e.what();
        return I3B_EXC_RUNTIME_ERROR;
    }
}

static unsigned grid_for(long long total) { return (unsigned) std::min<long long>((total + 255) / 256, 148 * 16); }

extern "C" {

const char* i3b_rangecomp_last_error(void) { return g_rc_error.c_str(); }

int i3b_rangecomp_create(const float* chirp, int chirp_size, int input_size, int max_batch, int mode,
                         I3B_RangeComp** out)
{
    return rc_guarded([&]() {
        if (!out) throw RcError(I3B_EXC_INVALID_ARGUMENT, "null handle pointer");
        *out = nullptr;
        if (!chirp || chirp_size < 1) throw RcError(I3B_EXC_INVALID_ARGUMENT, "chirp is empty");
        if (input_size < 1) throw RcError(I3B_EXC_DOMAIN_ERROR, "number of samples must be > 0");
        if (max_batch < 1) throw RcError(I3B_EXC_DOMAIN_ERROR, "max batch size must be > 0");
        if (mode < I3B_RANGECOMP_FULL || mode > I3B_RANGECOMP_SAME)
            throw RcError(I3B_EXC_RUNTIME_ERROR, "unexpected range compression mode");
        if ((long long) chirp_size + input_size - 1 > (1LL << 30))
            throw RcError(I3B_EXC_OVERFLOW_ERROR, "convolution length exceeds the supported FFT size");
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
            throw RcError(I3B_EXC_NO_DEVICE, "no CUDA device available; isce3_b200 has no CPU fallback");
        std::unique_ptr<I3B_RangeComp> rc(new I3B_RangeComp());
        RC_CK(cudaGetDevice(&rc->device));
        rc->chirp_size = chirp_size;
        rc->input_size = input_size;
        rc->max_batch = max_batch;
        rc->mode = mode;
        rc->fft_size = next_fast_power(output_size(chirp_size, input_size, I3B_RANGECOMP_FULL));
        rc->out_size = output_size(chirp_size, input_size, mode);
        RC_CK(cudaStreamCreateWithFlags(&rc->stream, cudaStreamNonBlocking));
        const size_t nfft = (size_t) rc->fft_size;
        RC_CK(cudaMalloc(&rc->d_ref, nfft * sizeof(float2)));
        RC_CK(cudaMalloc(&rc->d_work, nfft * max_batch * sizeof(float2)));
        RC_CK(cudaMalloc(&rc->d_in, (size_t) input_size * max_batch * sizeof(float2)));
        RC_CK(cudaMalloc(&rc->d_out, (size_t) rc->out_size * max_batch * sizeof(float2)));
        // matched filter: time-reversed complex conjugate of the chirp, zero-padded, in the
        // frequency domain, times 1/nfft  (RangeComp.cpp:24-39,108)
        std::vector<float2> ref(nfft, make_float2(0.f, 0.f));
        for (int i = 0; i < chirp_size; ++i) {
            const int s = chirp_size - 1 - i;
            ref[i] = make_float2(chirp[2 * s], -chirp[2 * s + 1]);
        }
        RC_CK(cudaMemcpyAsync(rc->d_ref, ref.data(), nfft * sizeof(float2), cudaMemcpyHostToDevice, rc->stream));
        cufftHandle one = rc->plan_for(1);
        RC_FFT(cufftExecC2C(one, rc->d_ref, rc->d_ref, CUFFT_FORWARD));
        RC_CK(cudaStreamSynchronize(rc->stream));
        RC_CK(cudaMemcpy(ref.data(), rc->d_ref, nfft * sizeof(float2), cudaMemcpyDeviceToHost));
        const float scale = (float) (1.0 / (double) nfft);
        for (auto& z : ref) {
            z.x *= scale;
            z.y *= scale;
        }
        RC_CK(cudaMemcpy(rc->d_ref, ref.data(), nfft * sizeof(float2), cudaMemcpyHostToDevice));
        *out = rc.release();
        return 0;
    });
}

int i3b_rangecomp_query(const I3B_RangeComp* rc, int* fft_size, int* out_size, int* first_valid_sample)
{
    if (!rc) return I3B_EXC_INVALID_ARGUMENT;
    if (fft_size) *fft_size = rc->fft_size;
    if (out_size) *out_size = rc->out_size;
    if (first_valid_sample)
        *first_valid_sample = rc->mode == I3B_RANGECOMP_FULL ? rc->chirp_size - 1
                              : rc->mode == I3B_RANGECOMP_VALID ? 0 : rc->chirp_size / 2;
    return 0;
}

int i3b_rangecomp_set_scaling(I3B_RangeComp* rc, const I3B_RangeCompScaling* sc)
{
    return rc_guarded([&]() {
        if (!rc) throw RcError(I3B_EXC_INVALID_ARGUMENT, "null handle");
        RC_CK(cudaSetDevice(rc->device));
        RC_CK(cudaStreamSynchronize(rc->stream));
        cudaFree(rc->d_col); rc->d_col = nullptr;
        cudaFree(rc->d_pat_idx); rc->d_pat_idx = nullptr;
        cudaFree(rc->d_pat_w); rc->d_pat_w = nullptr;
        cudaFree(rc->d_pattern); rc->d_pattern = nullptr;
        rc->n_pat = 0;
        if (!sc) return 0;
        const int n_out = rc->out_size;
        if (sc->column_scale) {
            RC_CK(cudaMalloc(&rc->d_col, (size_t) n_out * sizeof(float2)));
            RC_CK(cudaMemcpy(rc->d_col, sc->column_scale, (size_t) n_out * sizeof(float2), cudaMemcpyHostToDevice));
        }
        if (sc->n_pattern > 0) {
            if (!sc->slant_ranges || !sc->pattern_ranges)
                throw RcError(I3B_EXC_INVALID_ARGUMENT, "antenna pattern needs slant_ranges and pattern_ranges");
            const int np = sc->n_pattern;
            for (int i = 1; i < np; ++i)
                if (!(sc->pattern_ranges[i] > sc->pattern_ranges[i - 1]))
                    throw RcError(I3B_EXC_INVALID_ARGUMENT, "pattern_ranges must be increasing");
            // numpy.interp: clamped at both ends, linear in between
            std::vector<int> idx(n_out);
            std::vector<float> w(n_out);
            for (int j = 0; j < n_out; ++j) {
                const double x = sc->slant_ranges[j];
                if (!(x > sc->pattern_ranges[0])) {
                    idx[j] = 0;
                    w[j] = 0.f;
                } else if (!(x < sc->pattern_ranges[np - 1])) {
                    idx[j] = np - 1;
                    w[j] = 0.f;
                } else {
                    const int i = (int) (std::upper_bound(sc->pattern_ranges, sc->pattern_ranges + np, x) -
                                         sc->pattern_ranges) - 1;
                    idx[j] = i;
                    w[j] = (float) ((x - sc->pattern_ranges[i]) / (sc->pattern_ranges[i + 1] - sc->pattern_ranges[i]));
                }
            }
            RC_CK(cudaMalloc(&rc->d_pat_idx, (size_t) n_out * sizeof(int)));
            RC_CK(cudaMalloc(&rc->d_pat_w, (size_t) n_out * sizeof(float)));
            RC_CK(cudaMalloc(&rc->d_pattern, (size_t) rc->max_batch * np * sizeof(float2)));
            RC_CK(cudaMemcpy(rc->d_pat_idx, idx.data(), (size_t) n_out * sizeof(int), cudaMemcpyHostToDevice));
            RC_CK(cudaMemcpy(rc->d_pat_w, w.data(), (size_t) n_out * sizeof(float), cudaMemcpyHostToDevice));
            rc->n_pat = np;
        }
        return 0;
    });
}

// crop (+ fused corrections) of `batch` lines; `patterns`: host complex64 [batch][n_pat] or null
static void rc_crop(I3B_RangeComp* rc, float2* d_out, int batch, int offset, const float* patterns, cudaStream_t s)
{
    const int n_out = rc->out_size, nfft = rc->fft_size;
    const long long tout = (long long) batch * n_out;
    const float2* d_pat = nullptr;
    if (rc->n_pat > 0 && patterns) {
        RC_CK(cudaMemcpyAsync(rc->d_pattern, patterns, (size_t) batch * rc->n_pat * sizeof(float2),
                              cudaMemcpyHostToDevice, s));
        d_pat = rc->d_pattern;
    }
    if (rc->d_col || d_pat)
        rc_crop_scale_kernel<<<grid_for(tout), 256, 0, s>>>(d_out, rc->d_work, n_out, nfft, offset, tout, rc->d_col,
                                                           rc->d_pat_idx, rc->d_pat_w, d_pat, rc->n_pat);
    else
        rc_crop_kernel<<<grid_for(tout), 256, 0, s>>>(d_out, rc->d_work, n_out, nfft, offset, tout);
}

int i3b_rangecomp_execute(I3B_RangeComp* rc, float* out, const float* in, int batch, uint32_t flags)
{
    return i3b_rangecomp_execute_scaled(rc, out, in, batch, flags, nullptr);
}

int i3b_rangecomp_execute_scaled(I3B_RangeComp* rc, float* out, const float* in, int batch, uint32_t flags,
                                 const float* patterns)
{
    return rc_guarded([&]() {
        if (!rc || !out || !in) throw RcError(I3B_EXC_INVALID_ARGUMENT, "null argument");
        if (batch > rc->max_batch) throw RcError(I3B_EXC_LENGTH_ERROR, "batch size exceeds max batch");
        if (batch < 1) return 0;
        RC_CK(cudaSetDevice(rc->device));
        cudaStream_t s = rc->stream;
        const bool dev = (flags & I3B_FLAG_DEVICE_POINTERS) != 0;
        const int n_in = rc->input_size, n_out = rc->out_size, nfft = rc->fft_size;
        const float2* d_in = reinterpret_cast<const float2*>(in);
        float2* d_out = reinterpret_cast<float2*>(out);
        cudaEvent_t e0, e1;
        RC_CK(cudaEventCreate(&e0));
        RC_CK(cudaEventCreate(&e1));
        if (!dev) {
            RC_CK(cudaMemcpyAsync(rc->d_in, in, (size_t) batch * n_in * sizeof(float2), cudaMemcpyHostToDevice, s));
            d_in = rc->d_in;
            d_out = rc->d_out;
        }
        RC_CK(cudaEventRecord(e0, s));
        const long long tot = (long long) batch * nfft;
        rc_pad_kernel<<<grid_for(tot), 256, 0, s>>>(rc->d_work, d_in, n_in, nfft, tot);
        cufftHandle plan = rc->plan_for(batch);
        RC_FFT(cufftExecC2C(plan, rc->d_work, rc->d_work, CUFFT_FORWARD));
        rc_multiply_kernel<<<grid_for(tot), 256, 0, s>>>(rc->d_work, rc->d_ref, nfft, tot);
        RC_FFT(cufftExecC2C(plan, rc->d_work, rc->d_work, CUFFT_INVERSE));
        const int offset = rc->mode == I3B_RANGECOMP_FULL ? 0
                           : rc->mode == I3B_RANGECOMP_VALID ? rc->chirp_size - 1 : rc->chirp_size / 2;
        // NOTE Valid mode with chirp longer than input: offset follows the reference literally
        rc_crop(rc, d_out, batch, offset, patterns, s);
        RC_CK(cudaGetLastError());
        RC_CK(cudaEventRecord(e1, s));
        if (!dev)
            RC_CK(cudaMemcpyAsync(out, rc->d_out, (size_t) batch * n_out * sizeof(float2), cudaMemcpyDeviceToHost, s));
        RC_CK(cudaStreamSynchronize(s));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        rc->ms_last = ms;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return 0;
    });
}

int i3b_rangecomp_execute_to_device(I3B_RangeComp* rc, const float* in, int64_t lines, float** dev_out)
{
    return i3b_rangecomp_execute_to_device_scaled(rc, in, lines, nullptr, dev_out);
}

int i3b_rangecomp_execute_to_device_scaled(I3B_RangeComp* rc, const float* in, int64_t lines,
                                           const float* patterns, float** dev_out)
{
    return rc_guarded([&]() {
        if (!rc || !in || !dev_out) throw RcError(I3B_EXC_INVALID_ARGUMENT, "null argument");
        *dev_out = nullptr;
        if (lines < 1) throw RcError(I3B_EXC_DOMAIN_ERROR, "number of lines must be > 0");
        RC_CK(cudaSetDevice(rc->device));
        cudaStream_t s = rc->stream;
        const int n_in = rc->input_size, n_out = rc->out_size, nfft = rc->fft_size;
        float2* d_all = nullptr;
        RC_CK(cudaMalloc(&d_all, (size_t) lines * n_out * sizeof(float2)));
        const int offset = rc->mode == I3B_RANGECOMP_FULL ? 0
                           : rc->mode == I3B_RANGECOMP_VALID ? rc->chirp_size - 1 : rc->chirp_size / 2;
        double ms_total = 0.0;
        try {
            cudaEvent_t e0, e1;
            RC_CK(cudaEventCreate(&e0));
            RC_CK(cudaEventCreate(&e1));
            for (int64_t l0 = 0; l0 < lines; l0 += rc->max_batch) {
                const int batch = (int) std::min<int64_t>(rc->max_batch, lines - l0);
                RC_CK(cudaMemcpyAsync(rc->d_in, in + 2 * (size_t) l0 * n_in, (size_t) batch * n_in * sizeof(float2),
                                      cudaMemcpyHostToDevice, s));
                RC_CK(cudaEventRecord(e0, s));
                const long long tot = (long long) batch * nfft;
                rc_pad_kernel<<<grid_for(tot), 256, 0, s>>>(rc->d_work, rc->d_in, n_in, nfft, tot);
                cufftHandle plan = rc->plan_for(batch);
                RC_FFT(cufftExecC2C(plan, rc->d_work, rc->d_work, CUFFT_FORWARD));
                rc_multiply_kernel<<<grid_for(tot), 256, 0, s>>>(rc->d_work, rc->d_ref, nfft, tot);
                RC_FFT(cufftExecC2C(plan, rc->d_work, rc->d_work, CUFFT_INVERSE));
                rc_crop(rc, d_all + (size_t) l0 * n_out, batch, offset,
                        patterns ? patterns + 2 * (size_t) l0 * rc->n_pat : nullptr, s);
                RC_CK(cudaGetLastError());
                RC_CK(cudaEventRecord(e1, s));
                RC_CK(cudaStreamSynchronize(s)); // d_in is reused by the next chunk
                float ms = 0.f;
                cudaEventElapsedTime(&ms, e0, e1);
                ms_total += ms;
            }
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
        } catch (...) {
            cudaFree(d_all);
            throw;
        }
        rc->ms_last = ms_total;
        *dev_out = reinterpret_cast<float*>(d_all);
        return 0;
    });
}

int i3b_device_free(void* p)
{
    return cudaFree(p) == cudaSuccess ? 0 : I3B_EXC_CUDA_ERROR;
}

int i3b_device_to_host(void* dst, const void* src, size_t bytes)
{
    return cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : I3B_EXC_CUDA_ERROR;
}

double i3b_rangecomp_last_device_ms(const I3B_RangeComp* rc) { return rc ? rc->ms_last : 0.0; }

int i3b_rangecomp_destroy(I3B_RangeComp* rc)
{
    delete rc;
    return 0;
}

} // extern "C"
