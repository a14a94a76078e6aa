// C-ABI of the B200 TDBP backend (include/isce3_b200_backproject.h) and its host driver:
// argument validation, device staging, shard-per-GPU execution, statistics.
//
// Replaces the host sequence of isce3::cuda::focus::backproject
// (cxx/isce3/cuda/focus/Backproject.cu:468-754): 8 synchronous kernels + a blocking
// cudaMemcpy per pulse batch there; here one fused target-solve kernel, and pulse slabs
// uploaded on a copy stream while the accumulation kernel of the previous slab runs.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "launch.h"
#include "sinc_table.h"

namespace i3b {

struct ApiError : std::runtime_error {
    int code;
    ApiError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            char buf__[512];                                                                  \
            snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call,                     \
                     cudaGetErrorString(e__), __FILE__, __LINE__);                            \
            throw ApiError(e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver     \
                                   ? I3B_EXC_NO_DEVICE                                        \
                                   : I3B_EXC_CUDA_ERROR,                                      \
                           buf__);                                                            \
        }                                                                                     \
    } while (0)

static thread_local std::string g_last_error;
static thread_local I3B_Stats g_last_stats;

// Device buffers go through a per-process cache.  By DEFAULT it keeps nothing: every allocation
// is back with the driver when a call returns, as the reference's ownership contract says
// (Backproject.cu:691-695).  A caller that focuses block after block (focus.py:1988-2007) opts
// in to keeping them -- i3b_set_device_memory_pool(-1) or I3B_POOL_KEEP_MB=-1 -- and saves the
// cudaMalloc / cudaFree of ~6 GB per call; in a process that has enabled peer access (NCCL,
// multi-GPU) each of those maps / unmaps the allocation on all peers: 850 ms per call measured
// at 2 GPUs.  Buffers are returned to the cache only after the streams that used them have
// been synchronised; i3b_release_device_memory() hands the cached memory back at any time.
class DeviceCache {
public:
    void* get(size_t bytes)
    {
        int device = 0;
        CK(cudaGetDevice(&device));
        bytes = (bytes + kGranule - 1) / kGranule * kGranule;
        {
            std::lock_guard<std::mutex> lock(mtx_);
            int best = -1;
            for (int i = 0; i < (int) free_.size(); ++i) {
                const Block& b = free_[i];
                if (b.device != device || b.bytes < bytes || b.bytes > bytes + bytes / 4 + kGranule) continue;
                if (best < 0 || b.bytes < free_[best].bytes) best = i;
            }
            if (best >= 0) {
                Block b = free_[best];
                free_.erase(free_.begin() + best);
                cached_ -= b.bytes;
                live_.push_back(b);
                return b.p;
            }
        }
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaErrorMemoryAllocation) { // make room and retry once
            cudaGetLastError();
            release_all();
            e = cudaMalloc(&p, bytes);
        }
        CK(e);
        std::lock_guard<std::mutex> lock(mtx_);
        live_.push_back(Block {p, bytes, device});
        return p;
    }
    void put(void* p)
    {
        Block b {nullptr, 0, 0};
        {
            std::lock_guard<std::mutex> lock(mtx_);
            for (size_t i = 0; i < live_.size(); ++i)
                if (live_[i].p == p) {
                    b = live_[i];
                    live_.erase(live_.begin() + i);
                    break;
                }
            if (b.p && cached_ + b.bytes <= keep_limit()) {
                free_.push_back(b);
                cached_ += b.bytes;
                return;
            }
        }
        int cur = 0;
        cudaGetDevice(&cur);
        if (b.p && b.device != cur) cudaSetDevice(b.device);
        cudaFree(p);
        if (b.p && b.device != cur) cudaSetDevice(cur);
    }
    void release_all()
    {
        std::vector<Block> blocks;
        {
            std::lock_guard<std::mutex> lock(mtx_);
            blocks.swap(free_);
            cached_ = 0;
        }
        int cur = 0;
        cudaGetDevice(&cur);
        for (const Block& b : blocks) {
            cudaSetDevice(b.device);
            cudaFree(b.p);
        }
        cudaSetDevice(cur);
    }

private:
    struct Block {
        void* p;
        size_t bytes;
        int device;
    };
    static constexpr size_t kGranule = (size_t) 2 << 20;

public:
    // bytes the cache may hold between calls: 0 (the default, like the reference: everything goes
    // back to the driver before a call returns) unless the caller opted in through
    // i3b_set_device_memory_pool() or I3B_POOL_KEEP_MB (MiB; negative: unlimited)
    static std::atomic<long long>& limit_setting()
    {
        static std::atomic<long long> v {[] {
            if (const char* e = std::getenv("I3B_POOL_KEEP_MB")) {
                const long long mb = std::strtoll(e, nullptr, 10);
                return mb < 0 ? -1LL : mb << 20;
            }
            return 0LL;
        }()};
        return v;
    }

private:
    static size_t keep_limit()
    {
        const long long v = limit_setting().load();
        return v < 0 ? (~(size_t) 0 >> 1) : (size_t) v;
    }
    std::mutex mtx_;
    std::vector<Block> free_, live_;
    size_t cached_ = 0;
};

// Small pinned host blocks (device status words read back while the host keeps working);
// recycled for the life of the process, like the device buffers.
class PinnedPool {
public:
    void* get(size_t bytes)
    {
        {
            std::lock_guard<std::mutex> lock(mtx_);
            for (size_t i = 0; i < free_.size(); ++i)
                if (free_[i].second >= bytes) {
                    void* p = free_[i].first;
                    sizes_.push_back(free_[i]);
                    free_.erase(free_.begin() + i);
                    return p;
                }
        }
        void* p = nullptr;
        CK(cudaHostAlloc(&p, std::max<size_t>(bytes, 256), cudaHostAllocPortable));
        std::lock_guard<std::mutex> lock(mtx_);
        sizes_.push_back({p, std::max<size_t>(bytes, 256)});
        return p;
    }
    void put(void* p)
    {
        std::lock_guard<std::mutex> lock(mtx_);
        for (size_t i = 0; i < sizes_.size(); ++i)
            if (sizes_[i].first == p) {
                free_.push_back(sizes_[i]);
                sizes_.erase(sizes_.begin() + i);
                return;
            }
    }

private:
    std::mutex mtx_;
    std::vector<std::pair<void*, size_t>> free_, sizes_;
};

static PinnedPool& pinned_pool()
{
    static PinnedPool* pool = new PinnedPool();
    return *pool;
}

static DeviceCache& device_cache()
{
    static DeviceCache* cache = new DeviceCache(); // leaked on purpose: no CUDA calls at exit
    return *cache;
}

template<typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release()
    {
        if (p) device_cache().put(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count, cudaStream_t = nullptr)
    {
        release();
        n = count;
        if (count) p = static_cast<T*>(device_cache().get(count * sizeof(T)));
    }
    void upload(const T* src, size_t count, cudaStream_t s)
    {
        alloc(count);
        if (count) CK(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
};

struct Event {
    cudaEvent_t e = nullptr;
    Event() { CK(cudaEventCreate(&e)); }
    ~Event() { if (e) cudaEventDestroy(e); }
    void record(cudaStream_t s) { CK(cudaEventRecord(e, s)); }
};

static float elapsed(const Event& a, const Event& b)
{
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, a.e, b.e));
    return ms;
}

// Deep host copy of the caller's descriptors (a plan may outlive the caller's arrays).
struct HostScene {
    I3B_BackprojectArgs a;
    std::vector<double> out_pos, out_vel, in_pos, in_vel, out_dop, in_dop;
    std::vector<float> dem, kdata, range_cor;
    std::vector<double> pulse_times; // padded: [-kPulsePadLo, n + kPulsePadHi), empty: uniform grid
    std::vector<double> pulse_tn;    // pulse_times * nominal PRF (fast kernel's time axis)
    std::vector<float> pulse_xi;     // pulse_tn minus its value at the pulse's segment base
    std::vector<int> devices;
};

static int scene_segment(const I3B_BackprojectArgs& a)
{
    const I3B_Orbit& o = a.in_geometry.orbit;
    double vmax = 0;
    for (int i = 0; i < o.n; ++i) {
        const double* v = o.vel + 3 * (size_t) i;
        vmax = std::max(vmax, std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]));
    }
    const double r_min = std::min(a.in_geometry.grid.starting_range, a.out_geometry.grid.starting_range);
    return fast_segment(kC / a.fc, a.in_geometry.grid.prf, vmax, r_min);
}

static void validate(const I3B_BackprojectArgs& a)
{
    auto bad = [](const std::string& m) { throw ApiError(I3B_EXC_INVALID_ARGUMENT, m); };
    if (a.abi_version != I3B_ABI_VERSION) bad("ABI version mismatch");
    if (!a.in) bad("input signal data is null");
    if (!(a.dry_tropo_model == I3B_TROPO_NODELAY || a.dry_tropo_model == I3B_TROPO_TSX))
        bad("unexpected dry troposphere model"); // Backproject.cpp:78-83
    if (a.batch < 1) throw ApiError(I3B_EXC_DOMAIN_ERROR, "batch size must be > 0");
    const I3B_RadarGeometry &og = a.out_geometry, &ig = a.in_geometry;
    if (og.ref_epoch_sec != ig.ref_epoch_sec || og.ref_epoch_frac != ig.ref_epoch_frac)
        throw ApiError(I3B_EXC_RUNTIME_ERROR,
                       "input reference epoch must match output reference epoch"); // :88-92
    for (const I3B_RadarGeometry* g : {&og, &ig}) {
        if (g->grid.length < 0 || g->grid.width < 0) bad("negative grid dimension");
        if (g->grid.length > INT_MAX) throw ApiError(I3B_EXC_OVERFLOW_ERROR, "grid length exceeds max int");
        if (g->grid.width > INT_MAX) throw ApiError(I3B_EXC_OVERFLOW_ERROR, "grid width exceeds max int");
        if (!(g->grid.prf > 0)) bad("PRF must be positive");
        if (g->orbit.n < 2 || !g->orbit.pos || !g->orbit.vel) bad("orbit needs state vectors");
        if (g->orbit.method != I3B_ORBIT_HERMITE && g->orbit.method != I3B_ORBIT_LEGENDRE)
            bad("unknown orbit interpolation method");
        if (g->grid.look_side != I3B_LOOK_LEFT && g->grid.look_side != I3B_LOOK_RIGHT)
            bad("invalid look side");
        if (g->doppler.have_data) {
            if (!g->doppler.data || g->doppler.length < 1 || g->doppler.width < 1)
                bad("Doppler LUT has no data");
        }
    }
    if (a.dem.have_raster) {
        if (!a.dem.data || a.dem.length < 4 || a.dem.width < 4) bad("DEM raster too small");
        DevProj pj;
        if (!proj_setup(a.dem.epsg, kA, kE2, &pj))
            bad("unknown EPSG code for a raster DEM: " + std::to_string(a.dem.epsg) +
                " (supported: 4326, UTM 326xx/327xx, 3031, 3413, 6933)");
    }
    if (!(a.fc > 0) || !(a.ds > 0)) bad("fc and ds must be positive");
    if (a.pulse_times) {
        const int64_t n = ig.grid.length;
        for (int64_t k = 0; k < n; ++k) {
            if (!std::isfinite(a.pulse_times[k])) bad("pulse_times must be finite");
            if (k > 0 && !(a.pulse_times[k] > a.pulse_times[k - 1])) bad("pulse_times must be strictly increasing");
        }
    }
    if (a.mantissa_nbits < 0 || a.mantissa_nbits > 23) bad("mantissa_nbits must be in [0, 23]");
    const I3B_Kernel& k = a.kernel;
    switch (k.kind) {
    case I3B_KERNEL_BARTLETT:
    case I3B_KERNEL_LINEAR: break;
    case I3B_KERNEL_KNAB:
        if (!(k.bandwidth > 0.0 && k.bandwidth < 1.0)) throw ApiError(I3B_EXC_RUNTIME_ERROR, "Require 0 < bandwidth < 1");
        break;
    case I3B_KERNEL_TABULATED:
        if (!k.data || k.n < 2) throw ApiError(I3B_EXC_INVALID_ARGUMENT, "Require table size >= 2.");
        break;
    case I3B_KERNEL_CHEBY:
        if (!k.data || k.n < 1) throw ApiError(I3B_EXC_INVALID_ARGUMENT, "Need at least one coefficient.");
        break;
    default: throw ApiError(I3B_EXC_RUNTIME_ERROR, "not implemented"); // Backproject.cu:750-752
    }
    const double width = k.kind == I3B_KERNEL_LINEAR ? 2.0 : k.width;
    if (!(width > 0) || std::ceil(width) > 1024) bad("unsupported kernel width");
    if (a.n_devices < 0 || (a.n_devices > 0 && !a.devices)) bad("bad device list");
    if ((a.flags & (I3B_FLAG_DEVICE_POINTERS | I3B_FLAG_DEVICE_INPUT)) && a.n_devices > 1)
        bad("device-pointer mode is single-device only");
}

static void copy_geometry(const I3B_RadarGeometry& g, std::vector<double>& pos,
                          std::vector<double>& vel, std::vector<double>& dop, I3B_RadarGeometry& out)
{
    out = g;
    pos.assign(g.orbit.pos, g.orbit.pos + 3 * (size_t) g.orbit.n);
    vel.assign(g.orbit.vel, g.orbit.vel + 3 * (size_t) g.orbit.n);
    out.orbit.pos = pos.data();
    out.orbit.vel = vel.data();
    if (g.doppler.have_data) {
        dop.assign(g.doppler.data, g.doppler.data + (size_t) g.doppler.length * g.doppler.width);
        out.doppler.data = dop.data();
    }
}

static DevKernel make_kernel(const I3B_Kernel& k, const float* data)
{
    DevKernel d;
    std::memset(&d, 0, sizeof d);
    d.kind = k.kind;
    d.n = k.n;
    const double width = k.kind == I3B_KERNEL_LINEAR ? 2.0 : k.width; // core/Kernels.h:51
    d.halfwidth = std::fabs(width / 2.0);
    d.taps = (int) std::ceil(d.halfwidth * 2);
    d.bandwidth = k.bandwidth;
    if (k.kind == I3B_KERNEL_TABULATED) {
        d.imax = k.n - 2;
        const double dx = d.halfwidth / (k.n - 1.0);
        d.one_dx = (float) (1.0 / dx);
    } else if (k.kind == I3B_KERNEL_CHEBY) {
        d.cheb_scale = (float) (4.0 / width);
    }
    d.data = data;
    return d;
}

// Everything one GPU holds for its azimuth block of the output grid.
struct Shard {
    int device = 0;
    int line0 = 0, nlines = 0;
    // declared before the buffers (destroyed after them)
    struct Streams {
        cudaStream_t compute = nullptr, copy = nullptr, down = nullptr;
        ~Streams()
        {
            if (compute) cudaStreamDestroy(compute);
            if (copy) cudaStreamDestroy(copy);
            if (down) cudaStreamDestroy(down);
        }
    } streams;
    cudaStream_t& compute = streams.compute;
    cudaStream_t& copy = streams.copy;
    cudaStream_t& down = streams.down; // early device-to-host copies of the one-shot call
    DevBuf<double> out_pos, out_vel, in_pos, in_vel, out_dop, in_dop, pv, times, tn, sinc;
    DevBuf<float> xi;
    DevBuf<float> dem, kdata, height;
    DevBuf<PulseRec> pulse;
    DevBuf<PixelRec> pix;
    DevBuf<double2> acc;
    DevBuf<float2> out, rc, range_cor;
    DevBuf<DevStatus> status;
    DevBuf<TileInfo> tile_info;
    const float2* rc_dev = nullptr; // staged lines (or the caller's device pointer)
    int rc_pitch = 0, rc_k0 = 0, rc_rows = 0;
    bool rc_resident = false;
    SolveParams sp;
    AccumParams ap;
    DevKernel host_kernel;
    bool use_fast = false;
    int fast_variant = -1;
    I3B_Stats stats;
    int status_code = 0;
    std::string error;
    // target solve in flight: its status lands in pinned host memory, `ev_status` fires then
    DevStatus* status_host = nullptr;
    std::unique_ptr<Event> ev_solve0, ev_solve1, ev_status;
    bool solve_pending = false;
    // blocks API: per-pulse tables and the staged swath belong to the device's scene shard
    const PulseRec* pulse_shared = nullptr;
    const double* pv_shared = nullptr;
    const float2* range_cor_view = nullptr; // per-column output phasors of this shard's columns
    // One-shot call with host result arrays: parts of the result leave the device while the
    // accumulation is still running -- the image rows of the last launch's first parts as soon
    // as they are final, and (page-locked arrays) the height layer as soon as the solve is in.
    // Copies into pageable arrays block the calling thread, so they are issued only once every
    // kernel of the call is queued.
    float2* early_out = nullptr;   // this shard's slice of the caller's image (null: no early copies)
    float* early_height = nullptr; // ... of the caller's height layer
    bool early_pinned = false;     // ... and both are page-locked (copies do not block this thread)
    int out_rows_sent = 0;         // image rows [0, out_rows_sent) are on their way to the host
    bool height_sent = false;

    ~Shard()
    {
        if (status_host) pinned_pool().put(status_host);
        // Buffers go back to the cache right after this body, on this thread: make sure
        // nothing is still running on them (error paths may leave work in flight).
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess) cur = -1;
        if (compute || copy) cudaSetDevice(device);
        if (compute) cudaStreamSynchronize(compute);
        if (copy) cudaStreamSynchronize(copy);
        if (down) cudaStreamSynchronize(down);
        if (cur >= 0) cudaSetDevice(cur);
    }
};

} // namespace i3b

using namespace i3b;

struct I3B_Plan {
    HostScene hs;
    std::vector<std::unique_ptr<Shard>> shards;
    bool executed = false; // i3b_plan_execute has produced an image to download
};

namespace i3b {

static void shard_setup(const HostScene& hs, Shard& sh)
{
    const I3B_BackprojectArgs& a = hs.a;
    CK(cudaSetDevice(sh.device));
    // (cudaDeviceGetAttribute, not cudaGetDeviceProperties: the latter takes up to 100 ms)
    int cc_major = 0;
    CK(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, sh.device));
    if (cc_major < 10)
        throw ApiError(I3B_EXC_NO_DEVICE, "device " + std::to_string(sh.device) +
                                                  " is not sm_100-class; isce3_b200 has no fallback path");
    CK(cudaStreamCreateWithFlags(&sh.compute, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&sh.copy, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&sh.down, cudaStreamNonBlocking));
    cudaStream_t s = sh.compute;
    const I3B_RadarGeometry &og = a.out_geometry, &ig = a.in_geometry;
    sh.out_pos.upload(og.orbit.pos, 3 * (size_t) og.orbit.n, s);
    sh.out_vel.upload(og.orbit.vel, 3 * (size_t) og.orbit.n, s);
    sh.in_pos.upload(ig.orbit.pos, 3 * (size_t) ig.orbit.n, s);
    sh.in_vel.upload(ig.orbit.vel, 3 * (size_t) ig.orbit.n, s);
    if (og.doppler.have_data) sh.out_dop.upload(og.doppler.data, (size_t) og.doppler.length * og.doppler.width, s);
    if (ig.doppler.have_data) sh.in_dop.upload(ig.doppler.data, (size_t) ig.doppler.length * ig.doppler.width, s);
    if (a.dem.have_raster) sh.dem.upload(a.dem.data, (size_t) a.dem.length * a.dem.width, s);
    if (a.kernel.data && a.kernel.n > 0) sh.kdata.upload(a.kernel.data, (size_t) a.kernel.n, s);
    if (a.range_cor) sh.range_cor.upload(reinterpret_cast<const float2*>(a.range_cor), (size_t) og.grid.width, s);
    sh.range_cor_view = sh.range_cor.p;

    // 2-D sinc interpolation of a LUT / the DEM: the reference's kernel table (512 KB), once
    const bool any_sinc = (og.doppler.have_data && og.doppler.method == I3B_INTERP_SINC) ||
                          (ig.doppler.have_data && ig.doppler.method == I3B_INTERP_SINC) ||
                          (a.dem.have_raster && a.dem.method == I3B_INTERP_SINC);
    if (any_sinc) {
        static const std::vector<double> table = make_sinc_table();
        sh.sinc.upload(table.data(), table.size(), s);
    }
    const double* sinc_dev = sh.sinc.p;
    auto dev_orbit = [](const I3B_Orbit& o, const double* p, const double* v) {
        return DevOrbit {o.t0, o.dt, o.n, o.method, p, v};
    };
    auto dev_lut = [sinc_dev](const I3B_LUT2d& l, const double* d) {
        DevLUT2d r;
        r.sinc = sinc_dev;
        r.have_data = l.have_data; r.bounds_error = l.bounds_error; r.method = l.method;
        r.length = (int) l.length; r.width = (int) l.width;
        r.ref_value = l.ref_value; r.xstart = l.xstart; r.ystart = l.ystart; r.dx = l.dx; r.dy = l.dy;
        r.data = d;
        return r;
    };
    SolveParams& P = sh.sp;
    P.out_time = Linspace {og.grid.sensing_start, 1.0 / og.grid.prf, (int) og.grid.length};
    P.out_range = Linspace {og.grid.starting_range, og.grid.range_pixel_spacing, (int) og.grid.width};
    P.in_time = Linspace {ig.grid.sensing_start, 1.0 / ig.grid.prf, (int) ig.grid.length};
    P.in_times = nullptr;
    if (!hs.pulse_times.empty()) {
        sh.times.upload(hs.pulse_times.data(), hs.pulse_times.size(), s);
        P.in_times = sh.times.p + kPulsePadLo;
        sh.tn.upload(hs.pulse_tn.data(), hs.pulse_tn.size(), s);
        sh.xi.upload(hs.pulse_xi.data(), hs.pulse_xi.size(), s);
    }
    P.out_orbit = dev_orbit(og.orbit, sh.out_pos.p, sh.out_vel.p);
    P.in_orbit = dev_orbit(ig.orbit, sh.in_pos.p, sh.in_vel.p);
    P.out_doppler = dev_lut(og.doppler, sh.out_dop.p);
    P.in_doppler = dev_lut(ig.doppler, sh.in_dop.p);
    P.dem.have_raster = a.dem.have_raster; P.dem.epsg = a.dem.epsg; P.dem.method = a.dem.method;
    P.dem.length = (int) a.dem.length; P.dem.width = (int) a.dem.width;
    P.dem.ref_height = a.dem.ref_height; P.dem.xstart = a.dem.xstart; P.dem.ystart = a.dem.ystart;
    P.dem.dx = a.dem.dx; P.dem.dy = a.dem.dy; P.dem.data = sh.dem.p;
    P.dem.sinc = sinc_dev;
    proj_setup(a.dem.have_raster ? a.dem.epsg : 4326, kA, kE2, &P.dem.proj);
    P.r2g = a.rdr2geo;
    P.g2r = a.geo2rdr;
    P.wvl = kC / a.fc; // Backproject.cpp:119: wavelength from fc, not from the grid
    P.ds = a.ds;
    P.out_side = og.grid.look_side;
    P.in_side = ig.grid.look_side;
    P.tropo = a.dry_tropo_model;
    P.line0 = sh.line0;
    P.out_lines = sh.nlines;
    P.out_width = (int) og.grid.width;

    const size_t npix = (size_t) sh.nlines * og.grid.width;
    const int n_pulses = (int) ig.grid.length;
    sh.pulse.alloc((size_t) n_pulses + kPulsePadLo + kPulsePadHi, s);
    CK(cudaMemsetAsync(sh.pulse.p, 0, sh.pulse.n * sizeof(PulseRec), s));
    sh.pv.alloc((size_t) 6 * std::max(n_pulses, 1), s);
    sh.status.alloc(1, s);
    sh.pix.alloc(npix, s);
    sh.acc.alloc(npix, s);
    sh.out.alloc(npix, s);
    sh.height.alloc(npix, s);
    sh.tile_info.alloc((size_t) std::max(fast_tiles(sh.nlines, (int) og.grid.width), 1), s);

    AccumParams& A = sh.ap;
    std::memset(&A, 0, sizeof A);
    A.npix = (long long) npix;
    A.out_lines = sh.nlines;
    A.out_width = (int) og.grid.width;
    A.nr = (int) ig.grid.width;
    A.n_pulses = n_pulses;
    A.fc = a.fc;
    A.swst = 2. * ig.grid.starting_range / kC;       // Backproject.cpp:109-112
    A.dtau = 2. * ig.grid.range_pixel_spacing / kC;
    A.spacing_ratio = og.grid.range_pixel_spacing / ig.grid.range_pixel_spacing;
    A.kernel = make_kernel(a.kernel, sh.kdata.p);
    fast_tile_shape(&A.tile_az, &A.tile_rg);
    A.tiles_rg = (A.out_width + A.tile_rg - 1) / A.tile_rg;
    P.tile_az = A.tile_az;
    P.tile_rg = A.tile_rg;
    P.tiles_rg = A.tiles_rg;
    sh.host_kernel = make_kernel(a.kernel, a.kernel.data);
    char why[160] = "";
    I3B_TapPolyFit fit;
    sh.use_fast = !(a.flags & I3B_FLAG_FORCE_GENERIC) && fast_fit(sh.host_kernel, &fit, why, sizeof why);
    // Non-uniform pulse trains: the fast kernel's phase cubic runs over TIME (positions tn / xi
    // of the pulses on the segment's time axis) instead of the pulse index.
    A.seg = scene_segment(a);
    A.tn = hs.pulse_times.empty() ? nullptr : sh.tn.p + kPulsePadLo;
    A.xi = hs.pulse_times.empty() ? nullptr : sh.xi.p + kPulsePadLo;
    sh.fast_variant = sh.use_fast ? fit.imm_variant : -1;
    std::memset(&sh.stats, 0, sizeof sh.stats);
    sh.stats.taps = A.kernel.taps;
}

// pulse table + per-pixel target solve, asynchronous: kernels and the read-back of the status
// words (into pinned memory) are queued on the compute stream; shard_solve_finish() collects.
static void shard_solve_launch(const HostScene& hs, Shard& sh)
{
    CK(cudaSetDevice(sh.device));
    cudaStream_t s = sh.compute;
    if (!sh.status_host) sh.status_host = static_cast<DevStatus*>(pinned_pool().get(sizeof(DevStatus)));
    if (!sh.ev_solve0) {
        sh.ev_solve0.reset(new Event());
        sh.ev_solve1.reset(new Event());
        sh.ev_status.reset(new Event());
    }
    DevStatus init;
    std::memset(&init, 0, sizeof init);
    init.kmin = INT_MAX;
    init.kmax = INT_MIN;
    CK(cudaMemcpyAsync(sh.status.p, &init, sizeof init, cudaMemcpyHostToDevice, s));
    sh.ev_solve0->record(s);
    if (!sh.pulse_shared) {
        launch_pulse_table(sh.sp.in_orbit, sh.sp.in_time, sh.sp.in_times, hs.a.fc, sh.pulse.p + kPulsePadLo, sh.pv.p, sh.status.p, s);
        CK(cudaGetLastError());
    }
    launch_target_solve(sh.sp, sh.pix.p, sh.height.p, sh.tile_info.p, (int) sh.tile_info.n, sh.status.p, s);
    CK(cudaGetLastError());
    sh.ev_solve1->record(s);
    CK(cudaMemcpyAsync(sh.status_host, sh.status.p, sizeof(DevStatus), cudaMemcpyDeviceToHost, s));
    sh.ev_status->record(s);
    sh.solve_pending = true;
}

static bool shard_solve_ready(Shard& sh)
{
    return !sh.solve_pending || cudaEventQuery(sh.ev_status->e) == cudaSuccess;
}

static void shard_solve_finish(Shard& sh)
{
    if (!sh.solve_pending) return;
    CK(cudaSetDevice(sh.device));
    CK(cudaEventSynchronize(sh.ev_status->e));
    sh.solve_pending = false;
    const DevStatus st = *sh.status_host;
    sh.stats.ms_target_solve = elapsed(*sh.ev_solve0, *sh.ev_solve1);
    sh.stats.total_launches += 3;
    sh.stats.pixel_pulses = (double) st.pixel_pulses;
    if (st.hard_error) throw ApiError(st.hard_error, "orbit interpolation outside of orbit domain");
    sh.status_code = st.soft_error;
    if (st.kmin == INT_MAX) { // no pixel integrates anything
        sh.stats.pulse_first = sh.stats.pulse_last = 0;
    } else {
        sh.stats.pulse_first = st.kmin;
        sh.stats.pulse_last = st.kmax;
    }
}

static void shard_solve(const HostScene& hs, Shard& sh)
{
    shard_solve_launch(hs, sh);
    shard_solve_finish(sh);
}

// Pulses the shard's block can possibly integrate, from host-side bounds alone (output line
// times +- half the longest coherent processing interval of Backproject.cpp:176-193, plus the
// beam-centre shift the Doppler LUTs allow): what the one-shot call starts uploading WHILE the
// target solve runs, and how it decides which output rows have all their pulses on the device.
// The solve's exact range is checked against it afterwards (and every tile checks its own).
struct PulseBound {
    bool ok = false;
    double cpi = 0, shift = 0;      // s
    double t_out0 = 0, out_prf = 0; // output line times of the shard: t_out0 + row / out_prf
    double t_in0 = 0, in_prf = 0;
    int n_pulses = 0, nlines = 0;
    // first pulse rows >= row can need / one past the last pulse rows < row_end can need
    int lower(int row) const
    {
        const double lo = (t_out0 + row / out_prf - 0.5 * cpi - shift - t_in0) * in_prf - 64.0;
        return (int) std::max(0.0, std::min(std::floor(lo), (double) n_pulses));
    }
    int upper(int row_end) const
    {
        const double hi = (t_out0 + (row_end - 1) / out_prf + 0.5 * cpi + shift - t_in0) * in_prf + 64.0;
        return (int) std::max(0.0, std::min(std::ceil(hi), (double) n_pulses));
    }
    // leading rows whose pulses all lie below `landed`
    int rows_ready(int landed) const
    {
        if (landed >= upper(nlines)) return nlines;
        const double r = ((landed - 64.0) / in_prf + t_in0 - 0.5 * cpi - shift - t_out0) * out_prf;
        int rows = (int) std::max(0.0, std::min(std::floor(r) + 1.0, (double) nlines));
        while (rows > 0 && upper(rows) > landed) --rows; // (rounding)
        return rows;
    }
};

static PulseBound conservative_pulse_bound(const HostScene& hs, const Shard& sh)
{
    PulseBound B;
    const I3B_BackprojectArgs& a = hs.a;
    const I3B_RadarGrid &og = a.out_geometry.grid, &ig = a.in_geometry.grid;
    if (sh.nlines <= 0 || og.width <= 0 || ig.length <= 0 || !(og.prf > 0)) return B;
    const I3B_Orbit& orb = a.in_geometry.orbit;
    double pmax = 0, vmin = 1e300;
    for (int i = 0; i < orb.n; ++i) {
        const double* p = orb.pos + 3 * (size_t) i;
        const double* v = orb.vel + 3 * (size_t) i;
        pmax = std::max(pmax, std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]));
        vmin = std::min(vmin, std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]));
    }
    if (!(pmax > 0) || !(vmin > 0) || !std::isfinite(pmax) || !std::isfinite(vmin)) return B;
    auto fmax = [](const I3B_LUT2d& l) {
        if (!l.have_data) return std::fabs(l.ref_value);
        double m = 0;
        const size_t n = (size_t) l.length * (size_t) l.width;
        for (size_t i = 0; i < n; ++i) m = std::max(m, std::fabs(l.data[i]));
        return m;
    };
    const double fd = fmax(a.in_geometry.doppler) + fmax(a.out_geometry.doppler);
    if (!std::isfinite(fd)) return B;
    const double wvl = kC / a.fc;
    const double r_max = 1.05 * (og.starting_range + (og.width - 1) * og.range_pixel_spacing);
    B.cpi = wvl * r_max * (pmax / 6.30e6) / (2.0 * a.ds * vmin);
    B.shift = 1.2 * fd * wvl * r_max / (2.0 * vmin * vmin);
    B.t_out0 = og.sensing_start + sh.line0 / og.prf;
    B.out_prf = og.prf;
    B.t_in0 = ig.sensing_start;
    B.in_prf = ig.prf;
    B.n_pulses = (int) ig.length;
    B.nlines = sh.nlines;
    if (!std::isfinite(B.cpi) || !std::isfinite(B.shift)) return B;
    B.ok = B.upper(sh.nlines) > B.lower(0);
    return B;
}

// accumulate pulses [k0, k1) (must be staged) into acc
static void shard_accumulate(Shard& sh, int k0, int k1, cudaStream_t s, int line_begin = 0,
                             int line_end = 0, int k_landed = 0)
{
    AccumParams A = sh.ap;
    A.rc_pitch = sh.rc_pitch;
    A.rc_k0 = sh.rc_k0;
    A.rc_rows = sh.rc_rows;
    A.k_begin = k0;
    A.k_end = k1;
    A.line_begin = line_begin;
    A.line_end = line_end;
    A.k_landed = k_landed;
    A.tile_mask = nullptr;
    bool done = false;
    if (sh.use_fast) {
        const PulseRec* pulse = sh.pulse_shared ? sh.pulse_shared : sh.pulse.p + kPulsePadLo;
        const int rc = launch_accumulate_fast(A, sh.host_kernel, sh.pix.p, pulse, sh.rc_dev,
                                              sh.acc.p, sh.tile_info.p, sh.status.p, s);
        if (rc > 0) CK((cudaError_t) rc);
        if (rc == 0) {
            done = true;
            sh.stats.accumulate_launches += 1;
            sh.stats.total_launches += 1;
            if (sh.status_code != 0) {
                // tiles holding failed pixels were skipped: generic kernel on just those
                A.tile_mask = sh.tile_info.p;
                launch_accumulate_generic(A, sh.pix.p, sh.pv_shared ? sh.pv_shared : sh.pv.p, sh.rc_dev, sh.acc.p, s);
                CK(cudaGetLastError());
                sh.stats.total_launches += 1;
            }
        } else {
            sh.use_fast = false;
        }
    }
    if (!done) {
        A.tile_mask = nullptr;
        launch_accumulate_generic(A, sh.pix.p, sh.pv_shared ? sh.pv_shared : sh.pv.p, sh.rc_dev, sh.acc.p, s);
        CK(cudaGetLastError());
        sh.stats.accumulate_launches += 1;
        sh.stats.total_launches += 1;
    }
}

// Stage the needed pulses and integrate them.
//
//  * exact pulse range known up front (resident plan, device input, or no host-side bound):
//    slabs are copied back to back on the copy stream and an accumulation launch is issued
//    for everything copied so far whenever the compute stream has run dry -- launches end on
//    absolute multiples of the staged pulse tile, so the image does not depend on the cuts;
//  * one-shot call from host memory: the upload of a conservative pulse range starts while the
//    target solve is still running; once the solve is in, whole APERTURES are integrated row
//    block by row block, each as soon as the pulses its rows can need have landed ("row
//    wavefront": every tile runs exactly once, over its full pulse range, as in a resident plan).
static void shard_run(const HostScene& hs, Shard& sh, bool resident_only)
{
    const I3B_BackprojectArgs& a = hs.a;
    CK(cudaSetDevice(sh.device));
    cudaStream_t s = sh.compute;
    const int nr = sh.ap.nr;
    const int tk = fast_pulse_tile();
    const bool devptr = (a.flags & (I3B_FLAG_DEVICE_POINTERS | I3B_FLAG_DEVICE_INPUT)) != 0;
    CK(cudaMemsetAsync(sh.acc.p, 0, sh.acc.n * sizeof(double2), s));
    static const bool no_early = [] {
        const char* e = std::getenv("I3B_NO_EARLY_UPLOAD"); // test / tuning knob
        return e && std::atoi(e) != 0;
    }();
    // I3B_LAUNCH_PER_SLAB=1 (test knob): one launch per slab, like the reference
    static const bool per_slab = [] {
        const char* e = std::getenv("I3B_LAUNCH_PER_SLAB");
        return e && std::atoi(e) != 0;
    }();
    PulseBound bound;
    if (sh.solve_pending && !resident_only && !sh.rc_resident && !devptr && !no_early && !per_slab &&
        sh.ap.npix > 0 && hs.pulse_times.empty())
        bound = conservative_pulse_bound(hs, sh);
    bool early = bound.ok;
    if (!early) shard_solve_finish(sh);
    Event ea0, ea1;
    bool ea0_recorded = false;
    auto mark_start = [&]() {
        if (!ea0_recorded) {
            ea0.record(s);
            ea0_recorded = true;
        }
    };
    double ms_h2d = 0.0;
    const float2* in = reinterpret_cast<const float2*>(a.in);
    const int slab = std::max(a.batch, 1);
    std::vector<std::unique_ptr<Event>> landed;
    std::vector<int> slab_end; // landed[i] fires when pulses < slab_end[i] are on the device
    auto stage_input = [&](int b0, int b1) {
        sh.rc_pitch = (nr + 1) & ~1; // 16-byte line pitch (TMA global stride rule)
        sh.rc_k0 = b0;
        sh.rc_rows = b1 - b0;
        sh.rc.alloc((size_t) sh.rc_rows * sh.rc_pitch, s);
        sh.rc_dev = sh.rc.p;
        // pool memory is recycled: clear it, so that rows a pulse tile stages ahead of
        // the landed slab (and the pad column) never hold stale bit patterns
        CK(cudaMemsetAsync(sh.rc.p, 0, sh.rc.n * sizeof(float2), sh.copy));
    };
    auto upload_slab = [&](int k, int rows) {
        CK(cudaMemcpy2DAsync(sh.rc.p + (size_t) (k - sh.rc_k0) * sh.rc_pitch,
                             (size_t) sh.rc_pitch * sizeof(float2), in + (size_t) k * nr,
                             (size_t) nr * sizeof(float2), (size_t) nr * sizeof(float2), rows,
                             devptr ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, sh.copy));
        landed.emplace_back(new Event());
        landed.back()->record(sh.copy);
        slab_end.push_back(k + rows);
        sh.stats.h2d_bytes += (int64_t) rows * nr * (int64_t) sizeof(float2);
    };
    const auto t0 = std::chrono::steady_clock::now();
    static const bool debug_timing = std::getenv("I3B_DEBUG_TIMING") != nullptr;
    auto since = [&]() {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    double t_solved = 0, t_first = -1, t_queued = 0;

    if (early) {
        // ---- one-shot call: conservative upload while the solve runs, then row wavefront ----
        const int b0 = bound.lower(0), b1 = bound.upper(sh.nlines);
        stage_input(b0, b1);
        int uploaded = b0, kfirst = 0, klast = 0;
        bool solved = false, fits = true;
        size_t n_landed = 0; // slabs known to have landed
        int rows_done = 0;   // rows [0, rows_done) have been handed their whole apertures
        int k_split = 0;     // rows >= rows_done have been integrated over pulses < k_split
        bool split_set = false;
        const int tile_az = sh.ap.tile_az;
        int ramp = 8 * tk; // lines of the next slab while ramping up to `batch`
        // Launch policy once the solve is in.  Rows whose every pulse has landed get their whole
        // remaining aperture in one launch (row wavefront); while no row block is ready and the
        // GPU would idle -- the first rows need a full aperture on the device, most of the upload
        // when the block is short -- the landed pulses are integrated for ALL remaining rows
        // (pulse split, on a pulse-tile boundary).  Either way each pixel sees its pulses in
        // order, whole pulse tiles at a time: the image does not depend on the cuts.
        auto try_launch_rows = [&](bool all_queued) {
            if (!solved || !fits || klast <= kfirst || rows_done >= sh.nlines) return;
            if (!split_set) {
                k_split = kfirst;
                split_set = true;
            }
            while (n_landed < landed.size() && cudaEventQuery(landed[n_landed]->e) == cudaSuccess) ++n_landed;
            const bool all_landed = all_queued && n_landed == landed.size();
            // everything needed is queued and either landed or a launch is already running: the
            // last launch takes the rest and waits (in stream order) for the last copy
            if (all_queued && (all_landed || rows_done > 0)) {
                mark_start();
                if (!landed.empty()) CK(cudaStreamWaitEvent(s, landed.back()->e, 0));
                if (t_first < 0) t_first = since();
                // The last launch in three parts (3/4, 3/16 and 1/16 of its rows): the image rows
                // of a part are finalised and copied to the host while the next part is still
                // being summed, which leaves a sixteenth of the image for after the last kernel.
                struct PartCopy {
                    long long first, n;
                    std::unique_ptr<Event> done;
                } parts[2];
                int n_parts = 0;
                if (sh.early_out) {
                    const int total = sh.nlines - rows_done;
                    for (int part = 0; part < 2; ++part) {
                        const int upto = part == 0 ? total * 3 / 4 : total * 15 / 16;
                        const int r_mid = sh.nlines - total + (upto / tile_az) * tile_az;
                        if (r_mid - rows_done < 8 * tile_az || sh.nlines - r_mid < 4 * tile_az) continue;
                        shard_accumulate(sh, k_split, klast, s, rows_done, r_mid, 0);
                        const long long first = (long long) sh.out_rows_sent * sh.ap.out_width;
                        const long long n = (long long) r_mid * sh.ap.out_width - first;
                        launch_finalize(n, sh.ap.out_width, sh.pix.p + first, sh.acc.p + first, sh.out.p + first,
                                        sh.range_cor_view, hs.a.mantissa_nbits, s);
                        CK(cudaGetLastError());
                        sh.stats.total_launches += 1;
                        PartCopy& pc = parts[n_parts++];
                        pc.first = first;
                        pc.n = n;
                        pc.done.reset(new Event());
                        pc.done->record(s);
                        sh.out_rows_sent = r_mid;
                        rows_done = r_mid;
                    }
                }
                shard_accumulate(sh, k_split, klast, s, rows_done, sh.nlines, 0);
                rows_done = sh.nlines;
                // every kernel is queued: now the copies (into pageable arrays each one returns
                // when it is done -- by then the GPU is busy with the parts after it)
                for (int i = 0; i < n_parts; ++i) {
                    CK(cudaStreamWaitEvent(sh.down, parts[i].done->e, 0));
                    CK(cudaMemcpyAsync(sh.early_out + parts[i].first, sh.out.p + parts[i].first,
                                       (size_t) parts[i].n * sizeof(float2), cudaMemcpyDeviceToHost, sh.down));
                    sh.stats.d2h_bytes += (int64_t) parts[i].n * (int64_t) sizeof(float2);
                }
                if (sh.early_height && !sh.height_sent && sh.ap.npix > 0 && n_parts > 0) {
                    CK(cudaMemcpyAsync(sh.early_height, sh.height.p, sh.height.n * sizeof(float), cudaMemcpyDeviceToHost, sh.down));
                    sh.stats.d2h_bytes += (int64_t) (sh.height.n * sizeof(float));
                    sh.height_sent = true;
                }
                return;
            }
            const int have = n_landed ? slab_end[n_landed - 1] : b0;
            int r = bound.rows_ready(have);
            if (r < sh.nlines) r = (r / tile_az) * tile_az;
            // a launch of its own only for a good part of the block (with a slow host link the
            // launches queue up behind each other and the GPU never idles)
            if (r - rows_done >= std::max(tile_az, sh.nlines / 32)) {
                mark_start();
                if (t_first < 0) t_first = since();
                shard_accumulate(sh, k_split, klast, s, rows_done, r, have);
                rows_done = r;
                return;
            }
            // nothing row-ready: keep an idle GPU busy with the pulses that have landed
            const int k1 = std::min((have / tk) * tk, klast);
            if (k1 - k_split >= 8 * tk && (t_first < 0 || cudaStreamQuery(s) == cudaSuccess)) {
                mark_start();
                if (t_first < 0) t_first = since();
                shard_accumulate(sh, k_split, k1, s, rows_done, sh.nlines, 0);
                k_split = k1;
            }
        };
        while (true) {
            if (!solved && (uploaded >= b1 || shard_solve_ready(sh))) {
                shard_solve_finish(sh); // (blocks only when every slab is already queued)
                solved = true;
                t_solved = since();
                if (sh.early_height && sh.early_pinned && sh.ap.npix > 0) {
                    // the height layer is the solve's own output: final already
                    CK(cudaMemcpyAsync(sh.early_height, sh.height.p, sh.height.n * sizeof(float), cudaMemcpyDeviceToHost, sh.down));
                    sh.stats.d2h_bytes += (int64_t) (sh.height.n * sizeof(float));
                    sh.height_sent = true;
                }
                kfirst = sh.stats.pulse_first;
                klast = sh.stats.pulse_last;
                fits = !(klast > kfirst && (kfirst < b0 || klast > b1));
                if (!fits) break; // the bound did not hold (exotic geometry)
            }
            const int stop = solved ? std::max(std::min(klast, b1), b0) : b1;
            if (uploaded >= stop) break;
            // (the first slabs are short, so that the first launch does not wait for `batch`
            // lines to cross a host link that several devices may be sharing)
            const int rows = std::min(std::min(slab, ramp), stop - uploaded);
            ramp = std::min(2 * ramp, std::max(slab, ramp));
            upload_slab(uploaded, rows);
            uploaded += rows;
            try_launch_rows(false);
        }
        if (fits) {
            while (klast > kfirst && rows_done < sh.nlines) {
                try_launch_rows(true);
                if (rows_done < sh.nlines && n_landed < landed.size())
                    CK(cudaEventSynchronize(landed[n_landed]->e)); // wait for the next slab
            }
        } else {
            early = false; // fall through to the exact-range path below
            CK(cudaStreamSynchronize(sh.copy));
            landed.clear();
            slab_end.clear();
        }
    }
    if (!early && sh.stats.pulse_last > sh.stats.pulse_first && sh.ap.npix > 0) {
        // ---- exact pulse range known ----
        const int kfirst = sh.stats.pulse_first, klast = sh.stats.pulse_last;
        if (!sh.rc_resident && devptr && (nr % 2 == 0)) {
            sh.rc_dev = reinterpret_cast<const float2*>(a.in);
            sh.rc_pitch = nr;
            sh.rc_k0 = 0;
            sh.rc_rows = (int) a.in_geometry.grid.length;
            sh.rc_resident = true;
        }
        if (sh.rc_resident) {
            mark_start();
            if (!resident_only) shard_accumulate(sh, kfirst, klast, s);
        } else {
            stage_input(kfirst, klast);
            // Slabs are copied back to back; an accumulation launch is issued for everything
            // copied so far whenever the compute stream has run dry (and for the first and the
            // last slab), so a fast host link gives 2 launches per call and a slow one a few
            // more -- not one per slab, each of which would re-run every tile's prologue.
            int pending = kfirst; // first pulse not yet covered by an accumulation launch
            for (int k = kfirst; k < klast; k += slab) {
                const int rows = std::min(slab, klast - k);
                upload_slab(k, rows);
                if (resident_only) continue;
                const bool first = k == kfirst, last = k + rows >= klast;
                if (first || last || per_slab || cudaStreamQuery(s) == cudaSuccess) {
                    // launches end on absolute multiples of the staged pulse tile (the last one
                    // at klast): the image is then bit-identical for every batch size and
                    // whatever the host link's timing made of the launch boundaries
                    const int kend = last ? klast : ((k + rows) / tk) * tk;
                    if (kend <= pending) continue;
                    mark_start();
                    CK(cudaStreamWaitEvent(s, landed.back()->e, 0));
                    shard_accumulate(sh, pending, kend, s);
                    pending = kend;
                }
            }
            if (resident_only) sh.rc_resident = true;
        }
    }
    t_queued = since();
    if (!sh.rc_resident || resident_only) CK(cudaStreamSynchronize(sh.copy));
    ms_h2d = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    mark_start();
    ea1.record(s);
    if (resident_only) {
        CK(cudaStreamSynchronize(s));
        sh.stats.ms_h2d = ms_h2d;
        return;
    }
    const int kfirst = sh.stats.pulse_first, klast = sh.stats.pulse_last;
    if (sh.ap.npix > 0) {
        // (rows [0, out_rows_sent) were finalised, and are being copied out, already)
        const long long done = (long long) sh.out_rows_sent * sh.ap.out_width;
        launch_finalize(sh.ap.npix - done, sh.ap.out_width, sh.pix.p + done, sh.acc.p + done, sh.out.p + done,
                        sh.range_cor_view, hs.a.mantissa_nbits, s);
        CK(cudaGetLastError());
        sh.stats.total_launches += 1;
    }
    DevStatus st;
    CK(cudaMemcpyAsync(&st, sh.status.p, sizeof st, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    sh.stats.ms_accumulate = elapsed(ea0, ea1);
    sh.stats.ms_h2d = sh.rc_resident ? 0.0 : ms_h2d;
    if (debug_timing && early)
        fprintf(stderr, "[i3b]   run: solve in at %.1f, first launch %.1f, all queued %.1f, copies done %.1f, "
                        "synced %.1f ms; launches %d; solve kernel %.1f, accumulate %.1f ms\n",
                t_solved, t_first, t_queued, ms_h2d, since(), sh.stats.accumulate_launches,
                sh.stats.ms_target_solve, sh.stats.ms_accumulate);
    sh.stats.used_fast_kernel = sh.use_fast ? 1 : 0;
    sh.stats.fast_variant = sh.use_fast ? sh.fast_variant : -1;
    const bool redo_generic = st.window_overflow && sh.use_fast;
    if (redo_generic || st.premature) {
        // window_overflow: the staged range window was too small for some gather -> generic
        // kernel.  premature: a row-wavefront launch met a tile whose pulses had not all landed
        // (the host-side aperture bound failed) -> everything is on the device by now, again.
        if (redo_generic) sh.use_fast = false;
        if (sh.out_rows_sent > 0) {
            CK(cudaStreamSynchronize(sh.down)); // the rows copied out early are redone as well
            sh.out_rows_sent = 0;
        }
        DevStatus clr = st;
        clr.window_overflow = 0;
        clr.premature = 0;
        CK(cudaMemcpyAsync(sh.status.p, &clr, sizeof clr, cudaMemcpyHostToDevice, s));
        CK(cudaMemsetAsync(sh.acc.p, 0, sh.acc.n * sizeof(double2), s));
        Event eb0, eb1;
        eb0.record(s);
        shard_accumulate(sh, kfirst, klast, s);
        eb1.record(s);
        launch_finalize(sh.ap.npix, sh.ap.out_width, sh.pix.p, sh.acc.p, sh.out.p, sh.range_cor_view, hs.a.mantissa_nbits, s);
        CK(cudaStreamSynchronize(s));
        sh.stats.ms_accumulate += elapsed(eb0, eb1);
        if (redo_generic) sh.stats.used_fast_kernel = 0;
    }
}

// Page-locked host memory (cudaHostAlloc / cudaHostRegister): an asynchronous copy into it
// does not block the calling thread.
static bool host_is_page_locked(const void* p)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

static void shard_download(const HostScene& hs, Shard& sh, float* out, float* height)
{
    CK(cudaSetDevice(sh.device));
    cudaStream_t s = sh.compute;
    const size_t width = (size_t) hs.a.out_geometry.grid.width;
    const size_t off = (size_t) sh.line0 * width;
    const bool devptr = (hs.a.flags & I3B_FLAG_DEVICE_POINTERS) != 0;
    const auto kind = devptr ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    Event e0, e1;
    e0.record(s);
    if (out && sh.ap.npix) {
        // (rows the one-shot call sent ahead are not copied again)
        const size_t sent = sh.early_out ? (size_t) sh.out_rows_sent * width : 0;
        CK(cudaMemcpyAsync(reinterpret_cast<float2*>(out) + off + sent, sh.out.p + sent,
                           (sh.out.n - sent) * sizeof(float2), kind, s));
        sh.stats.d2h_bytes += (int64_t) ((sh.out.n - sent) * sizeof(float2));
    }
    if (height && sh.ap.npix && !(sh.early_height && sh.height_sent)) {
        CK(cudaMemcpyAsync(height + off, sh.height.p, sh.height.n * sizeof(float), kind, s));
        sh.stats.d2h_bytes += (int64_t) (sh.height.n * sizeof(float));
    }
    e1.record(s);
    CK(cudaStreamSynchronize(s));
    if (sh.down) CK(cudaStreamSynchronize(sh.down));
    sh.stats.ms_d2h = elapsed(e0, e1);
}

// `deep_copy`: a resident plan outlives the caller's descriptor arrays (orbit, LUTs, DEM raster,
// kernel table) and keeps its own copies; the blocking one-shot call reads them in place (a
// raster DEM is hundreds of MB).
static std::unique_ptr<I3B_Plan> make_plan(const I3B_BackprojectArgs* args, bool deep_copy)
{
    if (!args) throw ApiError(I3B_EXC_INVALID_ARGUMENT, "null argument block");
    validate(*args);
    std::unique_ptr<I3B_Plan> plan(new I3B_Plan());
    HostScene& hs = plan->hs;
    hs.a = *args;
    if (deep_copy) {
        copy_geometry(args->out_geometry, hs.out_pos, hs.out_vel, hs.out_dop, hs.a.out_geometry);
        copy_geometry(args->in_geometry, hs.in_pos, hs.in_vel, hs.in_dop, hs.a.in_geometry);
        if (args->dem.have_raster) {
            hs.dem.assign(args->dem.data, args->dem.data + (size_t) args->dem.length * args->dem.width);
            hs.a.dem.data = hs.dem.data();
        }
        if (args->kernel.data && args->kernel.n > 0) {
            hs.kdata.assign(args->kernel.data, args->kernel.data + args->kernel.n);
            hs.a.kernel.data = hs.kdata.data();
        }
        if (args->range_cor) {
            hs.range_cor.assign(args->range_cor, args->range_cor + 2 * (size_t) args->out_geometry.grid.width);
            hs.a.range_cor = hs.range_cor.data();
        }
    }
    if (args->pulse_times && args->in_geometry.grid.length > 0) {
        // explicit pulse times that ARE the uniform grid take the uniform path (bit-identical to
        // a call without them); otherwise keep a copy extended past both ends with the first /
        // last pulse interval (the pulse table is padded, see kPulsePadLo / kPulsePadHi)
        const I3B_RadarGrid& g = args->in_geometry.grid;
        const int64_t n = g.length;
        const double dt = 1.0 / g.prf;
        bool uniform = true;
        // (to a few ulps: t0 + k / prf and t0 + k * (1 / prf) are the same grid)
        for (int64_t k = 0; k < n && uniform; ++k) {
            const double t = g.sensing_start + (double) k * dt;
            uniform = std::fabs(args->pulse_times[k] - t) <= 8.0 * 2.220446049250313e-16 * std::max(std::fabs(t), dt * (double) n);
        }
        if (!uniform) {
            const double* T = args->pulse_times;
            const double d0 = n > 1 ? T[1] - T[0] : dt, d1 = n > 1 ? T[n - 1] - T[n - 2] : dt;
            hs.pulse_times.resize((size_t) n + kPulsePadLo + kPulsePadHi);
            for (int64_t k = -kPulsePadLo; k < n + kPulsePadHi; ++k)
                hs.pulse_times[(size_t) (k + kPulsePadLo)] =
                        k < 0 ? T[0] + (double) k * d0 : (k >= n ? T[n - 1] + (double) (k - n + 1) * d1 : T[k]);
            // the fast kernel's time axis: nominal pulse intervals, relative to each pulse's
            // geometry-segment base (absolute multiples of the segment length)
            const int seg = scene_segment(*args);
            hs.pulse_tn.resize(hs.pulse_times.size());
            hs.pulse_xi.resize(hs.pulse_times.size());
            for (size_t i = 0; i < hs.pulse_times.size(); ++i) hs.pulse_tn[i] = hs.pulse_times[i] * g.prf;
            for (int64_t k = -kPulsePadLo; k < n + kPulsePadHi; ++k) {
                int64_t b = (k >= 0 ? k / seg : -((-k + seg - 1) / seg)) * seg; // floor to a multiple of seg
                b = std::max<int64_t>(b, -kPulsePadLo);
                hs.pulse_xi[(size_t) (k + kPulsePadLo)] =
                        (float) (hs.pulse_tn[(size_t) (k + kPulsePadLo)] - hs.pulse_tn[(size_t) (b + kPulsePadLo)]);
            }
        }
    }
    hs.a.pulse_times = nullptr; // (the scene's own copy is the one used from here on)
    int ndev_avail = 0;
    {
        cudaError_t e = cudaGetDeviceCount(&ndev_avail);
        if (e != cudaSuccess || ndev_avail < 1)
            throw ApiError(I3B_EXC_NO_DEVICE,
                           std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                   "); isce3_b200 has no CPU fallback");
    }
    if (args->n_devices > 0) hs.devices.assign(args->devices, args->devices + args->n_devices);
    else {
        int cur = 0;
        CK(cudaGetDevice(&cur));
        hs.devices.assign(1, cur);
    }
    for (int d : hs.devices)
        if (d < 0 || d >= ndev_avail) throw ApiError(I3B_EXC_INVALID_ARGUMENT, "device ordinal out of range");
    hs.a.devices = hs.devices.data();
    hs.a.n_devices = (int) hs.devices.size();
    // contiguous azimuth blocks, one per device
    // Blocks are whole rows of fast-kernel tiles (except the last), so a pixel shares its CTA
    // with the same neighbours however many devices there are: which staged tiles count as
    // aperture edges -- and with it the arithmetic path of every pixel -- does not depend on
    // the device list.
    const int lines = (int) args->out_geometry.grid.length;
    int tile_az = 1, tile_rg = 1;
    fast_tile_shape(&tile_az, &tile_rg);
    const int units = (lines + tile_az - 1) / tile_az;
    const int nsh = (int) std::min<size_t>(hs.devices.size(), (size_t) std::max(units, 1));
    int line = 0, unit = 0;
    for (int i = 0; i < nsh; ++i) {
        const int nu = units / nsh + (i < units % nsh ? 1 : 0);
        unit += nu;
        const int n = std::min(unit * tile_az, lines) - line;
        std::unique_ptr<Shard> sh(new Shard());
        sh->device = hs.devices[i];
        sh->line0 = line;
        sh->nlines = n;
        line += n;
        plan->shards.push_back(std::move(sh));
    }
    return plan;
}

template<class F>
static void for_each_shard(I3B_Plan& plan, F f)
{
    if (plan.shards.size() == 1) {
        f(*plan.shards[0]);
        return;
    }
    std::vector<std::thread> th;
    for (auto& shp : plan.shards) {
        Shard* sh = shp.get();
        th.emplace_back([sh, &f]() {
            try {
                f(*sh);
            } catch (const ApiError& e) {
                sh->status_code = e.code;
                sh->error = e.what();
            } catch (const std::exception& e) {
                sh->status_code = I3B_EXC_RUNTIME_ERROR;
                sh->error = e.what();
            }
        });
    }
    for (auto& t : th) t.join();
    for (auto& shp : plan.shards)
        if (shp->status_code < 0) throw ApiError(shp->status_code, shp->error);
}

static int merge_status(I3B_Plan& plan)
{
    int code = 0;
    for (auto& sh : plan.shards)
        if (sh->status_code > 0) code = sh->status_code;
    return code;
}

static void merge_stats(I3B_Plan& plan, double ms_total)
{
    I3B_Stats t;
    std::memset(&t, 0, sizeof t);
    t.pulse_first = INT_MAX;
    t.pulse_last = INT_MIN;
    t.used_fast_kernel = 1;
    for (auto& shp : plan.shards) {
        const I3B_Stats& s = shp->stats;
        t.pixel_pulses += s.pixel_pulses;
        t.ms_h2d = std::max(t.ms_h2d, s.ms_h2d);
        t.ms_target_solve = std::max(t.ms_target_solve, s.ms_target_solve);
        t.ms_accumulate = std::max(t.ms_accumulate, s.ms_accumulate);
        t.ms_d2h = std::max(t.ms_d2h, s.ms_d2h);
        t.accumulate_launches += s.accumulate_launches;
        t.total_launches += s.total_launches;
        t.used_fast_kernel &= s.used_fast_kernel;
        t.taps = s.taps;
        t.fast_variant = s.fast_variant;
        t.h2d_bytes += s.h2d_bytes;
        t.d2h_bytes += s.d2h_bytes;
        if (s.pulse_last > s.pulse_first) {
            t.pulse_first = std::min(t.pulse_first, s.pulse_first);
            t.pulse_last = std::max(t.pulse_last, s.pulse_last);
        }
    }
    if (t.pulse_first == INT_MAX) t.pulse_first = t.pulse_last = 0;
    t.ms_total = ms_total;
    t.n_devices = (int) plan.shards.size();
    g_last_stats = t;
}

// The library selects devices on the calling thread (shards, plan destruction, the buffer
// cache); the caller's current device is put back on every way out of an entry point -- the
// reference never touches it (focus.py:1592-1593 sets it once for all isce3.cuda modules).
struct DeviceGuard {
    int prev = -1;
    DeviceGuard()
    {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            prev = -1;
            cudaGetLastError();
        }
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template<class F>
static int guarded(F&& f)
{
    DeviceGuard device_guard;
    try {
        return f();
    } catch (const ApiError& e) {
        g_last_error = e.what();
        return e.code;
    } catch (const std::bad_alloc&) {
        g_last_error = "out of host memory";
        return I3B_EXC_RUNTIME_ERROR;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return I3B_EXC_RUNTIME_ERROR;
    }
}

} // namespace i3b

namespace {

struct GeomBatch {
    cudaStream_t s = nullptr;
    DevBuf<double> pos, vel, lut, sinc;
    DevBuf<float> dem;
    DevBuf<DevStatus> status;
    ~GeomBatch()
    {
        if (s) {
            cudaStreamSynchronize(s);
            cudaStreamDestroy(s);
        }
    }
    void check_device()
    {
        int dev = 0, cc_major = 0;
        CK(cudaGetDevice(&dev));
        CK(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
        if (cc_major < 10)
            throw ApiError(I3B_EXC_NO_DEVICE, "current device is not sm_100-class; isce3_b200 has no fallback path");
        CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        status.alloc(1);
        CK(cudaMemsetAsync(status.p, 0, sizeof(DevStatus), s));
    }
    DevOrbit orbit(const I3B_Orbit& o)
    {
        if (o.n < 2 || !o.pos || !o.vel) throw ApiError(I3B_EXC_INVALID_ARGUMENT, "orbit needs state vectors");
        if (o.method != I3B_ORBIT_HERMITE && o.method != I3B_ORBIT_LEGENDRE)
            throw ApiError(I3B_EXC_INVALID_ARGUMENT, "unknown orbit interpolation method");
        pos.upload(o.pos, 3 * (size_t) o.n, s);
        vel.upload(o.vel, 3 * (size_t) o.n, s);
        return DevOrbit {o.t0, o.dt, o.n, o.method, pos.p, vel.p};
    }
    int finish()
    {
        DevStatus st;
        CK(cudaMemcpyAsync(&st, status.p, sizeof st, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        return st.soft_error;
    }
};

// host array -> device (or pass a device pointer through)
template<typename T>
const T* stage_in(DevBuf<T>& buf, const T* p, size_t n, bool devptr, cudaStream_t s)
{
    if (devptr || !p) return p;
    buf.upload(p, n, s);
    return buf.p;
}

} // namespace

// ---- blocks API: one swath, many output blocks -----------------------------------------
// The workflow focuses an image block by block, handing the WHOLE range-compressed swath to
// every call (nisar/workflows/focus.py:726-783 plan_processing_blocks, :1988-2007).  Here the
// swath, the per-pulse tables, DEM, LUTs and kernel table are uploaded once per device (the
// "scene" shard); each block only allocates its per-pixel buffers, and free devices pull the
// next block from a shared counter.
struct I3B_Blocks {
    HostScene hs;
    std::vector<std::unique_ptr<Shard>> scenes; // one per device
    std::mutex mtx;
};

namespace i3b {

static void scene_setup(const HostScene& hs, Shard& sc)
{
    sc.line0 = 0;
    sc.nlines = 0;
    shard_setup(hs, sc);
    CK(cudaSetDevice(sc.device));
    cudaStream_t s = sc.compute;
    const I3B_BackprojectArgs& a = hs.a;
    const int nr = sc.ap.nr, n_pulses = sc.ap.n_pulses;
    // per-pulse tables, once
    DevStatus init;
    std::memset(&init, 0, sizeof init);
    init.kmin = INT_MAX;
    init.kmax = INT_MIN;
    CK(cudaMemcpyAsync(sc.status.p, &init, sizeof init, cudaMemcpyHostToDevice, s));
    launch_pulse_table(sc.sp.in_orbit, sc.sp.in_time, sc.sp.in_times, a.fc, sc.pulse.p + kPulsePadLo, sc.pv.p, sc.status.p, s);
    CK(cudaGetLastError());
    DevStatus st;
    CK(cudaMemcpyAsync(&st, sc.status.p, sizeof st, cudaMemcpyDeviceToHost, s));
    // the whole swath, resident
    const bool devptr = (a.flags & (I3B_FLAG_DEVICE_POINTERS | I3B_FLAG_DEVICE_INPUT)) != 0;
    if (devptr && nr % 2 == 0) {
        sc.rc_dev = reinterpret_cast<const float2*>(a.in);
        sc.rc_pitch = nr;
    } else {
        sc.rc_pitch = (nr + 1) & ~1;
        sc.rc.alloc((size_t) n_pulses * sc.rc_pitch, s);
        sc.rc_dev = sc.rc.p;
        if (sc.rc_pitch != nr) CK(cudaMemsetAsync(sc.rc.p, 0, sc.rc.n * sizeof(float2), sc.copy));
        CK(cudaMemcpy2DAsync(sc.rc.p, (size_t) sc.rc_pitch * sizeof(float2), a.in, (size_t) nr * sizeof(float2),
                             (size_t) nr * sizeof(float2), (size_t) n_pulses,
                             devptr ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, sc.copy));
        sc.stats.h2d_bytes += (int64_t) n_pulses * nr * (int64_t) sizeof(float2);
    }
    sc.rc_k0 = 0;
    sc.rc_rows = n_pulses;
    sc.rc_resident = true;
    CK(cudaStreamSynchronize(s));
    CK(cudaStreamSynchronize(sc.copy));
    if (st.hard_error) throw ApiError(st.hard_error, "orbit interpolation outside of orbit domain");
}

// one output block on the device of `scene`
static void block_focus(const HostScene& hs, const Shard& scene, const I3B_RadarGrid& g, float* out,
                        float* height, Shard& b)
{
    const I3B_BackprojectArgs& a = hs.a;
    if (g.length < 0 || g.width < 0 || g.length > INT_MAX || g.width > INT_MAX || !(g.prf > 0))
        throw ApiError(I3B_EXC_INVALID_ARGUMENT, "bad block grid");
    CK(cudaSetDevice(scene.device));
    b.device = scene.device;
    b.line0 = 0;
    b.nlines = (int) g.length;
    CK(cudaStreamCreateWithFlags(&b.compute, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&b.copy, cudaStreamNonBlocking));
    cudaStream_t s = b.compute;
    b.sp = scene.sp;
    b.ap = scene.ap;
    b.host_kernel = scene.host_kernel;
    b.use_fast = scene.use_fast;
    b.fast_variant = scene.fast_variant;
    SolveParams& P = b.sp;
    P.out_time = Linspace {g.sensing_start, 1.0 / g.prf, (int) g.length};
    P.out_range = Linspace {g.starting_range, g.range_pixel_spacing, (int) g.width};
    P.out_side = g.look_side;
    P.line0 = 0;
    P.out_lines = b.nlines;
    P.out_width = (int) g.width;
    const size_t npix = (size_t) b.nlines * (size_t) g.width;
    AccumParams& A = b.ap;
    A.npix = (long long) npix;
    A.out_lines = b.nlines;
    A.out_width = (int) g.width;
    A.spacing_ratio = g.range_pixel_spacing / a.in_geometry.grid.range_pixel_spacing;
    A.tiles_rg = (A.out_width + A.tile_rg - 1) / A.tile_rg;
    P.tiles_rg = A.tiles_rg;
    b.status.alloc(1, s);
    b.pix.alloc(npix, s);
    b.acc.alloc(npix, s);
    b.out.alloc(npix, s);
    b.height.alloc(npix, s);
    b.tile_info.alloc((size_t) std::max(fast_tiles(b.nlines, (int) g.width), 1), s);
    b.pulse_shared = scene.pulse.p + kPulsePadLo;
    b.pv_shared = scene.pv.p;
    b.rc_dev = scene.rc_dev;
    b.rc_pitch = scene.rc_pitch;
    b.rc_k0 = scene.rc_k0;
    b.rc_rows = scene.rc_rows;
    b.rc_resident = true;
    b.range_cor_view = nullptr;
    if (scene.range_cor.p) {
        // the image's per-column phasors, at this block's columns
        const I3B_RadarGrid& full = a.out_geometry.grid;
        const double c = (g.starting_range - full.starting_range) / full.range_pixel_spacing;
        const long long col0 = (long long) std::llround(c);
        if (std::fabs(c - (double) col0) > 1e-6 || g.range_pixel_spacing != full.range_pixel_spacing || col0 < 0 ||
            col0 + g.width > full.width)
            throw ApiError(I3B_EXC_INVALID_ARGUMENT, "range_cor needs blocks on the columns of the output grid");
        b.range_cor_view = scene.range_cor.p + col0;
    }
    std::memset(&b.stats, 0, sizeof b.stats);
    b.stats.taps = A.kernel.taps;
    shard_solve(hs, b);
    shard_run(hs, b, false);
    // shard_download() addresses a shard inside a full image; a block IS its own image
    Event e0, e1;
    e0.record(s);
    if (out && npix) {
        CK(cudaMemcpyAsync(out, b.out.p, npix * sizeof(float2), cudaMemcpyDeviceToHost, s));
        b.stats.d2h_bytes += (int64_t) (npix * sizeof(float2));
    }
    if (height && npix) {
        CK(cudaMemcpyAsync(height, b.height.p, npix * sizeof(float), cudaMemcpyDeviceToHost, s));
        b.stats.d2h_bytes += (int64_t) (npix * sizeof(float));
    }
    e1.record(s);
    CK(cudaStreamSynchronize(s));
    b.stats.ms_d2h = elapsed(e0, e1);
}

} // namespace i3b

extern "C" {

int i3b_blocks_create(const I3B_BackprojectArgs* args, I3B_Blocks** out_blocks)
{
    return guarded([&]() {
        if (!out_blocks) throw ApiError(I3B_EXC_INVALID_ARGUMENT, "null handle pointer");
        *out_blocks = nullptr;
        auto plan = make_plan(args, true); // (validation, deep copy of the descriptors, device list)
        std::unique_ptr<I3B_Blocks> B(new I3B_Blocks());
        B->hs = std::move(plan->hs);
        // re-point the descriptor copies (vectors moved with the scene keep their storage)
        HostScene& hs = B->hs;
        hs.a.devices = hs.devices.data();
        std::vector<std::thread> th;
        std::vector<std::string> errors(hs.devices.size());
        std::vector<int> codes(hs.devices.size(), 0);
        for (size_t i = 0; i < hs.devices.size(); ++i) {
            B->scenes.emplace_back(new Shard());
            B->scenes.back()->device = hs.devices[i];
        }
        for (size_t i = 0; i < hs.devices.size(); ++i) {
            Shard* sc = B->scenes[i].get();
            th.emplace_back([&, i, sc]() {
                try {
                    scene_setup(hs, *sc);
                } catch (const ApiError& e) {
                    codes[i] = e.code;
                    errors[i] = e.what();
                } catch (const std::exception& e) {
                    codes[i] = I3B_EXC_RUNTIME_ERROR;
                    errors[i] = e.what();
                }
            });
        }
        for (auto& t : th) t.join();
        for (size_t i = 0; i < codes.size(); ++i)
            if (codes[i] < 0) throw ApiError(codes[i], errors[i]);
        hs.a.in = hs.a.flags & (I3B_FLAG_DEVICE_POINTERS | I3B_FLAG_DEVICE_INPUT) ? hs.a.in : nullptr;
        hs.a.out = nullptr;
        hs.a.height = nullptr;
        *out_blocks = B.release();
        return 0;
    });
}

int i3b_blocks_run(I3B_Blocks* B, int32_t n, const I3B_RadarGrid* grids, float* const* out, float* const* height)
{
    return guarded([&]() {
        if (!B || n < 0 || (n > 0 && (!grids || !out))) throw ApiError(I3B_EXC_INVALID_ARGUMENT, "null argument");
        const auto t0 = std::chrono::steady_clock::now();
        std::atomic<int> next {0};
        const size_t ndev = B->scenes.size();
        std::vector<int> codes(ndev, 0), soft(ndev, 0);
        std::vector<std::string> errors(ndev);
        std::vector<I3B_Stats> totals(ndev);
        auto worker = [&](size_t d) {
            I3B_Stats& t = totals[d];
            std::memset(&t, 0, sizeof t);
            t.used_fast_kernel = 1;
            t.fast_variant = -1;
            try {
                for (;;) {
                    const int i = next.fetch_add(1);
                    if (i >= n) break;
                    Shard b;
                    block_focus(B->hs, *B->scenes[d], grids[i], out[i], height ? height[i] : nullptr, b);
                    if (b.status_code > 0) soft[d] = b.status_code;
                    const I3B_Stats& s = b.stats;
                    t.pixel_pulses += s.pixel_pulses;
                    t.ms_target_solve += s.ms_target_solve;
                    t.ms_accumulate += s.ms_accumulate;
                    t.ms_d2h += s.ms_d2h;
                    t.accumulate_launches += s.accumulate_launches;
                    t.total_launches += s.total_launches;
                    t.used_fast_kernel &= s.used_fast_kernel;
                    t.fast_variant = s.fast_variant;
                    t.taps = s.taps;
                    t.d2h_bytes += s.d2h_bytes;
                }
            } catch (const ApiError& e) {
                codes[d] = e.code;
                errors[d] = e.what();
            } catch (const std::exception& e) {
                codes[d] = I3B_EXC_RUNTIME_ERROR;
                errors[d] = e.what();
            }
        };
        if (ndev == 1) {
            worker(0);
        } else {
            std::vector<std::thread> th;
            for (size_t d = 0; d < ndev; ++d) th.emplace_back(worker, d);
            for (auto& t : th) t.join();
        }
        for (size_t d = 0; d < ndev; ++d)
            if (codes[d] < 0) throw ApiError(codes[d], errors[d]);
        I3B_Stats t;
        std::memset(&t, 0, sizeof t);
        t.used_fast_kernel = 1;
        int code = 0;
        for (size_t d = 0; d < ndev; ++d) {
            const I3B_Stats& s = totals[d];
            t.pixel_pulses += s.pixel_pulses;
            t.ms_target_solve = std::max(t.ms_target_solve, s.ms_target_solve);
            t.ms_accumulate = std::max(t.ms_accumulate, s.ms_accumulate);
            t.ms_d2h = std::max(t.ms_d2h, s.ms_d2h);
            t.accumulate_launches += s.accumulate_launches;
            t.total_launches += s.total_launches;
            t.used_fast_kernel &= s.used_fast_kernel;
            t.fast_variant = s.fast_variant;
            t.taps = s.taps;
            t.d2h_bytes += s.d2h_bytes;
            t.h2d_bytes += B->scenes[d]->stats.h2d_bytes;
            if (soft[d] > 0) code = soft[d];
        }
        t.n_devices = (int) ndev;
        t.ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        g_last_stats = t;
        return code;
    });
}

int i3b_blocks_destroy(I3B_Blocks* B)
{
    return guarded([&]() {
        delete B;
        return 0;
    });
}

int i3b_backproject(const I3B_BackprojectArgs* args)
{
    return guarded([&]() {
        const auto t0 = std::chrono::steady_clock::now();
        auto lap = [&]() {
            return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        };
        if (args && !args->out) throw ApiError(I3B_EXC_INVALID_ARGUMENT, "output array is null");
        auto plan = make_plan(args, false); // descriptors are used in place: the call blocks
        float* out = args->out;
        float* height = args->height;
        double t_setup = 0, t_solve = 0, t_run = 0, t_down = 0;
        const bool single = plan->shards.size() == 1; // phase laps: single-shard calls only
        for_each_shard(*plan, [&](Shard& sh) {
            shard_setup(plan->hs, sh);
            if (single) t_setup = lap();
            // the one-shot call streams from the caller's buffer (no deep copy of `in`), and
            // starts doing so while the target solve is still running
            shard_solve_launch(plan->hs, sh);
            if (single) t_solve = lap();
            if (!(args->flags & I3B_FLAG_DEVICE_POINTERS)) {
                const size_t off = (size_t) sh.line0 * (size_t) args->out_geometry.grid.width;
                sh.early_out = reinterpret_cast<float2*>(out) + off;
                sh.early_height = height ? height + off : nullptr;
                sh.early_pinned = host_is_page_locked(out) && (!height || host_is_page_locked(height));
            }
            shard_run(plan->hs, sh, false);
            if (single) t_run = lap();
            shard_download(plan->hs, sh, out, height);
            if (single) t_down = lap();
        });
        I3B_Plan* raw = plan.get();
        const int status = merge_status(*raw);
        const double ms_work = lap();
        merge_stats(*raw, ms_work);
        plan.reset(); // device memory is released inside the timed call
        g_last_stats.ms_total = lap();
        if (single && std::getenv("I3B_DEBUG_TIMING"))
            fprintf(stderr, "[i3b] setup %.1f solve %.1f run %.1f download %.1f teardown %.1f total %.1f ms\n",
                    t_setup, t_solve - t_setup, t_run - t_solve, t_down - t_run,
                    g_last_stats.ms_total - ms_work, g_last_stats.ms_total);
        return status;
    });
}

int i3b_plan_create(const I3B_BackprojectArgs* args, I3B_Plan** out_plan)
{
    return guarded([&]() {
        if (!out_plan) throw ApiError(I3B_EXC_INVALID_ARGUMENT, "null plan pointer");
        *out_plan = nullptr;
        auto plan = make_plan(args, true);
        for_each_shard(*plan, [&](Shard& sh) {
            shard_setup(plan->hs, sh);
            shard_solve(plan->hs, sh);   // needed to know which pulses to keep resident
            shard_run(plan->hs, sh, true); // upload only
        });
        plan->hs.a.in = nullptr; // the caller's buffer is no longer referenced
        plan->hs.a.out = nullptr;
        plan->hs.a.height = nullptr;
        *out_plan = plan.release();
        return 0;
    });
}

int i3b_plan_execute(I3B_Plan* plan)
{
    return guarded([&]() {
        if (!plan) throw ApiError(I3B_EXC_INVALID_ARGUMENT, "null plan");
        const auto t0 = std::chrono::steady_clock::now();
        for_each_shard(*plan, [&](Shard& sh) {
            const int64_t h2d = sh.stats.h2d_bytes;
            std::memset(&sh.stats, 0, sizeof sh.stats);
            sh.stats.taps = sh.ap.kernel.taps;
            sh.stats.h2d_bytes = h2d;
            shard_solve(plan->hs, sh);
            shard_run(plan->hs, sh, false);
        });
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        merge_stats(*plan, ms);
        plan->executed = true;
        return merge_status(*plan);
    });
}

int i3b_plan_download(I3B_Plan* plan, float* out, float* height)
{
    return guarded([&]() {
        if (!plan) throw ApiError(I3B_EXC_INVALID_ARGUMENT, "null plan");
        if (!plan->executed)
            throw ApiError(I3B_EXC_RUNTIME_ERROR, "i3b_plan_download before any i3b_plan_execute: no image yet");
        for_each_shard(*plan, [&](Shard& sh) { shard_download(plan->hs, sh, out, height); });
        return 0;
    });
}

int i3b_plan_destroy(I3B_Plan* plan)
{
    return guarded([&]() {
        delete plan;
        return 0;
    });
}

int i3b_last_stats(I3B_Stats* stats)
{
    if (!stats) return I3B_EXC_INVALID_ARGUMENT;
    *stats = g_last_stats;
    return 0;
}

const char* i3b_last_error(void) { return g_last_error.c_str(); }

const char* i3b_version(void) { return "isce3_b200 0.1.0 (sm_100a)"; }

int i3b_current_device(void)
{
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess) return -1;
    return d;
}

int i3b_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int i3b_measure_peaks(int device, I3B_Peaks* peaks)
{
    return guarded([&]() {
        if (!peaks) throw ApiError(I3B_EXC_INVALID_ARGUMENT, "null peaks pointer");
        const int rc = measure_peaks(device, peaks);
        if (rc != 0) CK((cudaError_t) rc);
        return 0;
    });
}

int i3b_set_device_memory_pool(int64_t keep_mb)
{
    DeviceCache::limit_setting().store(keep_mb < 0 ? -1LL : (long long) keep_mb << 20);
    if (keep_mb == 0) device_cache().release_all();
    return 0;
}

int i3b_release_device_memory(void)
{
    return guarded([&]() {
        device_cache().release_all();
        return 0;
    });
}

int i3b_rdr2geo_bracket_batch(const I3B_Orbit* orbit, const I3B_DEM* dem, double wavelength,
                              int32_t look_side, const I3B_Rdr2GeoBracketParams* params, int64_t n,
                              const double* aztime, const double* slant_range, const double* doppler,
                              double* xyz, int32_t* status, uint32_t flags)
{
    return guarded([&]() {
        if (!orbit || !dem || !params || n < 0 || (n > 0 && (!aztime || !slant_range || !xyz)))
            throw ApiError(I3B_EXC_INVALID_ARGUMENT, "null argument");
        if (look_side != I3B_LOOK_LEFT && look_side != I3B_LOOK_RIGHT)
            throw ApiError(I3B_EXC_INVALID_ARGUMENT, "invalid look side");
        if (n == 0) return 0;
        const bool devptr = (flags & I3B_FLAG_DEVICE_POINTERS) != 0;
        GeomBatch g;
        g.check_device();
        const DevOrbit o = g.orbit(*orbit);
        DevDEM d;
        std::memset(&d, 0, sizeof d);
        d.have_raster = dem->have_raster; d.epsg = dem->epsg; d.method = dem->method;
        d.length = (int) dem->length; d.width = (int) dem->width;
        d.ref_height = dem->ref_height; d.xstart = dem->xstart; d.ystart = dem->ystart;
        d.dx = dem->dx; d.dy = dem->dy;
        if (dem->have_raster) {
            if (!dem->data || dem->length < 4 || dem->width < 4)
                throw ApiError(I3B_EXC_INVALID_ARGUMENT, "DEM raster too small");
            if (dem->method == I3B_INTERP_SINC) {
                static const std::vector<double> table = make_sinc_table();
                g.sinc.upload(table.data(), table.size(), g.s);
                d.sinc = g.sinc.p;
            }
            if (!proj_setup(dem->epsg, kA, kE2, &d.proj))
                throw ApiError(I3B_EXC_INVALID_ARGUMENT, "unknown EPSG code for a raster DEM");
            g.dem.upload(dem->data, (size_t) dem->length * dem->width, g.s);
            d.data = g.dem.p;
        } else {
            proj_setup(4326, kA, kE2, &d.proj);
        }
        DevBuf<double> b_t, b_r, b_f, b_x;
        DevBuf<int> b_st;
        const double* dt = stage_in(b_t, aztime, (size_t) n, devptr, g.s);
        const double* dr = stage_in(b_r, slant_range, (size_t) n, devptr, g.s);
        const double* df = stage_in(b_f, doppler, (size_t) n, devptr, g.s);
        double* dx = xyz;
        int* dst = status;
        if (!devptr) {
            b_x.alloc(3 * (size_t) n);
            dx = b_x.p;
            if (status) {
                b_st.alloc((size_t) n);
                dst = b_st.p;
            }
        }
        launch_rdr2geo_batch(o, d, wavelength, look_side, *params, n, dt, dr, df, dx, dst, g.status.p, g.s);
        CK(cudaGetLastError());
        if (!devptr) {
            CK(cudaMemcpyAsync(xyz, dx, 3 * (size_t) n * sizeof(double), cudaMemcpyDeviceToHost, g.s));
            if (status) CK(cudaMemcpyAsync(status, dst, (size_t) n * sizeof(int), cudaMemcpyDeviceToHost, g.s));
        }
        return g.finish();
    });
}

int i3b_geo2rdr_bracket_batch(const I3B_Orbit* orbit, const I3B_LUT2d* doppler, double wavelength,
                              int32_t look_side, const I3B_Geo2RdrBracketParams* params, int64_t n,
                              const double* xyz, double* aztime, double* slant_range, int32_t* status,
                              uint32_t flags)
{
    return guarded([&]() {
        if (!orbit || !doppler || !params || n < 0 || (n > 0 && (!aztime || !slant_range || !xyz)))
            throw ApiError(I3B_EXC_INVALID_ARGUMENT, "null argument");
        if (look_side != I3B_LOOK_LEFT && look_side != I3B_LOOK_RIGHT)
            throw ApiError(I3B_EXC_INVALID_ARGUMENT, "invalid look side");
        if (n == 0) return 0;
        const bool devptr = (flags & I3B_FLAG_DEVICE_POINTERS) != 0;
        GeomBatch g;
        g.check_device();
        const DevOrbit o = g.orbit(*orbit);
        DevLUT2d l;
        std::memset(&l, 0, sizeof l);
        l.have_data = doppler->have_data; l.bounds_error = doppler->bounds_error; l.method = doppler->method;
        l.length = (int) doppler->length; l.width = (int) doppler->width;
        l.ref_value = doppler->ref_value; l.xstart = doppler->xstart; l.ystart = doppler->ystart;
        l.dx = doppler->dx; l.dy = doppler->dy;
        if (doppler->have_data) {
            if (!doppler->data || doppler->length < 1 || doppler->width < 1)
                throw ApiError(I3B_EXC_INVALID_ARGUMENT, "Doppler LUT has no data");
            if (doppler->method == I3B_INTERP_SINC) {
                static const std::vector<double> table = make_sinc_table();
                g.sinc.upload(table.data(), table.size(), g.s);
                l.sinc = g.sinc.p;
            }
            g.lut.upload(doppler->data, (size_t) doppler->length * doppler->width, g.s);
            l.data = g.lut.p;
        }
        DevBuf<double> b_x, b_t, b_r;
        DevBuf<int> b_st;
        const double* dx = stage_in(b_x, xyz, 3 * (size_t) n, devptr, g.s);
        double *dt = aztime, *dr = slant_range;
        int* dst = status;
        if (!devptr) {
            b_t.alloc((size_t) n);
            b_r.alloc((size_t) n);
            dt = b_t.p;
            dr = b_r.p;
            if (status) {
                b_st.alloc((size_t) n);
                dst = b_st.p;
            }
        }
        launch_geo2rdr_batch(o, l, wavelength, look_side, *params, n, dx, dt, dr, dst, g.status.p, g.s);
        CK(cudaGetLastError());
        if (!devptr) {
            CK(cudaMemcpyAsync(aztime, dt, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost, g.s));
            CK(cudaMemcpyAsync(slant_range, dr, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost, g.s));
            if (status) CK(cudaMemcpyAsync(status, dst, (size_t) n * sizeof(int), cudaMemcpyDeviceToHost, g.s));
        }
        return g.finish();
    });
}

int i3b_fit_tap_polynomials(const I3B_Kernel* kernel, I3B_TapPolyFit* fit)
{
    return guarded([&]() {
        if (!kernel || !fit) throw ApiError(I3B_EXC_INVALID_ARGUMENT, "null argument");
        const DevKernel hk = make_kernel(*kernel, kernel->data);
        char why[160] = "";
        fast_fit(hk, fit, why, sizeof why);
        g_last_error = why;
        return 0;
    });
}

} // extern "C"
