// Fast accumulation kernel of the B200 TDBP backend (sm_100a): the O(pixels x pulses)
// hot loop, sumCoherent of cxx/isce3/focus/Backproject.cpp:30-63, re-designed for B200.
//
//  * CTA = 8 warps.  A CTA owns a tile of TILE_AZ x TILE_RG output pixels; every thread
//    owns PX = 2 range-adjacent pixels; the TMA producer role rotates over the warps.
//    Tiles are numbered azimuth-major inside groups of GROUP_RG range columns so that
//    co-resident CTAs share pulse lines in L2.
//  * The producer streams pulse tiles (TK range-compressed lines clipped to the range
//    window the CTA's pixels can touch) into a multi-stage shared-memory ring with TMA
//    (cp.async.bulk.tensor.2d) completing on mbarriers; out-of-swath samples arrive as
//    zeros (TMA OOB fill), which IS the CPU reference's zero-padded edge window
//    (core/detail/Interp1d.h:54-80).
//  * Geometry in FP64 only at segment boundaries (64 or 128 pulses, chosen per scene by
//    fast_segment(); exact carrier phase from an 80-byte per-pulse record); inside a segment
//    the phase is a cubic in FP32 and the sample coordinate an affine function of it, both
//    pixels of a thread packed in FFMA2.  Segments and pulse tiles are anchored at ABSOLUTE
//    pulse indices and the FP64 sums continue from what earlier launches left, so the image
//    does not depend on how the pulses were split over launches / devices.
//  * A pulse tile is worked through in RUNS of SUB = 8 pulses.  Each run re-centres the cubic
//    on its first pulse with whole turns removed (RunPoly: the phase the loop evaluates stays
//    below ~40 rad) and takes the cheapest path it qualifies for: "steady" (the window position
//    provably does not move: straight-line code, compile-time window parity, no per-pulse
//    rounding or index arithmetic), per-pulse rounding (tile_body), or the out-of-line
//    aperture-edge body.
//  * Interpolation weights: per-tap polynomials in the fractional sample offset, fitted on
//    the host to the caller's kernel (table-lerp, Chebyshev or Knab), each tap pair at the
//    lowest degree that keeps the residual.  Kernels with a build-time coefficient table
//    (tap_poly_imm.h) take the coefficients as FFMA2 immediates; any other kernel reads them
//    from shared memory (broadcast LDS.128).  Taps m and K-1-m share even/odd parts (the
//    kernels are even functions).
//  * The two pixels of a thread share one register window of K+1 samples read with
//    LDS.128 (stride-16B across lanes: conflict-free); 16- and 32-tap kernels process the
//    window in chunks of 4 tap pairs.
//
// Numerics vs the reference: weights differ from table-lerp by the fit residual (checked
// on the host, <= 3e-5 abs; 2.7e-6 for the workflow's kernel), the FP32 phase carries a few
// 1e-6 rad, partial sums are FP32 within a pulse tile and FP64 across tiles and launches.
// Measured against the reference CPU code: relative RMS 5e-7 (point targets) ... 1.6e-5
// (noise-like airborne scene); gate 1e-4.
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "geometry.cuh"
#include "kernels.cuh"
#include "launch.h"
#include "tap_poly_imm.h" // generated: baked coefficient tables of the specialised kernels

#include <climits>

namespace i3b {

constexpr int TILE_RG = 128;  // output range pixels per CTA tile
#ifndef I3B_TILE_AZ
#define I3B_TILE_AZ 4
#endif
// Ring of staged pulse tiles.  I3B_LAST_ARRIVER = 0 (default): 4 stages of 16 pulses, the
// producer role rotates over the warps and refills a stage released two tiles ago.
// I3B_LAST_ARRIVER = 1: two stages of I3B_TK = 32 pulses; the warp that is LAST to finish a tile
// refills its stage with the tile after next, so nobody waits for a release and per-tile work
// is paid every 32 pulses -- measured on B200: no gain (K = 9: 0.743 vs 0.749 of FP32 peak; the
// slowest warp now also carries the window computation), kept for experiments.
#ifndef I3B_LAST_ARRIVER
#define I3B_LAST_ARRIVER 0
#endif
#ifndef I3B_TK
#define I3B_TK (I3B_LAST_ARRIVER ? 32 : 16)
#endif
#ifndef I3B_NSTAGE
#define I3B_NSTAGE (I3B_LAST_ARRIVER ? 2 : 4)
#endif
constexpr int TILE_AZ = I3B_TILE_AZ;   // output azimuth lines per CTA tile
constexpr int PX = 2;         // pixels per thread (range-adjacent)
constexpr int NTHREADS = TILE_AZ * TILE_RG / PX; // 256 threads
constexpr int TK = I3B_TK;    // pulses per stage
constexpr int NSTAGE = I3B_NSTAGE;
constexpr int POLY_OFFSET = 256;   // per-tap polynomial rows (copied from the kernel parameter)
#ifndef I3B_ROTATE_PRODUCER
#define I3B_ROTATE_PRODUCER 1
#endif
// Pulse tiles are requested I3B_PREFETCH_DIST tiles ahead.  With NSTAGE - 2 the stage being
// refilled was consumed TWO tiles ago, so the producing warp practically never waits for
// the slower warps of the CTA (with NSTAGE - 1 it waited ~15 % of its time, ncu r01 v4).
#ifndef I3B_PREFETCH_DIST
#define I3B_PREFETCH_DIST (I3B_LAST_ARRIVER ? I3B_NSTAGE - 1 : I3B_NSTAGE - 2)
#endif
constexpr int PREFETCH = I3B_PREFETCH_DIST;
#ifndef I3B_GROUP_RG
#define I3B_GROUP_RG 2
#endif
constexpr int GROUP_RG = I3B_GROUP_RG; // range columns per azimuth-major tile group
#ifndef I3B_EDGE_SPLIT
#define I3B_EDGE_SPLIT 1
#endif
static_assert(PREFETCH >= 1 && PREFETCH < I3B_NSTAGE, "prefetch distance must be in [1, NSTAGE)");
constexpr int NWARPS_ROT = TILE_AZ * TILE_RG / 2 / 32;
constexpr int HEADER_BYTES = 1024; // barriers, window origins, corner pixels, polynomial rows
constexpr int MAX_TAPS = 32;
constexpr int MAX_COEF = 8;   // degree <= 7

// Polynomial coefficients of the per-tap weights, in f = frac - 0.5 in [-0.5, 0.5):
//   w_m(f) = E_m(f^2) + f * O_m(f^2),  w_{K-1-m}(f) = E_m(f^2) - f * O_m(f^2)
// c_even[m][i] multiplies f^(2i), c_odd[m][i] multiplies f^(2i+1); m < ceil(K/2).
// One 32-byte row per tap pair {e0..e3, o0..o3} so a row is two 128-bit constant loads.
struct __align__(16) TapPoly {
    float e[MAX_COEF / 2];
    float o[MAX_COEF / 2];
};
// The fitted rows travel as a kernel parameter (constant bank 0), not in a __constant__
// symbol: concurrent calls on one device with different kernels must not share them.
struct PolyTable {
    TapPoly rows[MAX_TAPS / 2 + 1];
};

// ---- PTX wrappers --------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t) __cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
                 "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes,
                                             uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
                 "[%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct FastParams {
    long long npix;
    int out_lines, out_width;
    int nr, rc_k0, rc_rows;
    int k_begin, k_end;
    int n_pulses;   // pulses in the input grid (size of the pulse table)
    int W;          // staged samples per pulse (even, <= 256)
    int tiles_rg, tiles_az;
    int tile_j0;    // first tile row of this launch (tiles_az counts the rows of the launch)
    int k_landed;   // pulses below this index are on the device (0: no such limit)
    int seg;        // pulses per geometry segment (a power of two, multiple of TK)
    double G;       // samples per cycle: 1 / (fc * dtau)
    float Gf, Grf;  // (float) G and (float) (G / 2 pi): samples per turn / per radian of carrier phase
    double U0;      // swst / dtau
    double fc;
    int zero;       // always 0 (opaque to the compiler, see Weights::load_top)
    // non-uniform pulse times (null: uniform): tn[k] = t_k * nominal PRF (FP64), xi[k] = tn[k]
    // minus tn at the base of k's geometry segment (FP32); both indexed like the pulse table
    const double* tn;
    const float* xi;
};


// Shared-memory carve-up (dynamic): per stage [TK][W] float2 then [TK] PulseRec.
struct SmemHeader {
    uint64_t full[NSTAGE];
    uint64_t empty[NSTAGE];
    int winlo[NSTAGE];
    int done[NSTAGE]; // warps that have finished the tile held by the stage (last-arriver scheme)
    int kb, ke;     // CTA pulse range
    int bad;        // tile holds a failed pixel -> generic kernel
    int ks_max, ke_min; // every pixel of the CTA integrates pulses [ks_max, ke_min)
    int pad[3];
    double corner[4][4]; // x, y, z, fc*tau_atm of the 4 corner pixels
};

static_assert(sizeof(SmemHeader) <= POLY_OFFSET, "shared-memory header overflows its slot");

__host__ __device__ inline size_t stage_bytes(int W)
{
    size_t b = (size_t) TK * W * sizeof(float2);
    return (b + 127) & ~(size_t) 127;
}

// exact two-way sample coordinate u for one pixel/pulse (producer's window bounds)
__device__ inline double sample_coord(const double* c, const PulseRec& r, double G, double U0)
{
    const double xx = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    const double r2 = fma(c[0], r.m2px, fma(c[1], r.m2py, fma(c[2], r.m2pz, xx + r.pp)));
    const double s = sqrt(r2);
    const double tcyc = fma(r.Cs, s, fma(c[0], r.vBx, fma(c[1], r.vBy, fma(c[2], r.vBz, c[3] + r.E))));
    return fma(tcyc, G, -U0);
}

// ---- packed FP32x2 arithmetic (Blackwell FFMA2: one issue slot, two FMAs) -----------------
// A 64-bit register pair holds (lo, hi).  ptxas folds pack2(x, x) into the scalar-broadcast
// operand form of FFMA2 (R.F32), so "real weight x complex sample" is ONE instruction.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ f32x2 bcast2(float x) { return pack2(x, x); }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr));
    return v;
}
__device__ __forceinline__ f32x2 lds64(uint32_t addr)
{
    f32x2 v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double2 lds_d2(uint32_t addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

// Coefficient sources for the weight polynomials.  CoefBank reads the fitted rows from the
// constant bank (any kernel; costs ~22 LDC per pulse at K = 9 because FFMA2 has no
// constant-bank operand form).  CoefImm<V> reads a build-time table (tap_poly_imm.h): after
// unrolling every coefficient is an FFMA2 immediate -- no loads at all.  The launcher picks
// CoefImm<V> only when the run-time fit of the caller's kernel equals table V.
// Degrees are per tap PAIR (the host fit gives outer pairs, whose weights are small, the
// lowest degree that meets its tolerance): a baked table knows each pair's degree at compile
// time and evaluates no more terms than that; the general kernel runs every pair at the
// maximum degree with zeros for the missing leading coefficients -- fma(0, h, c) == c, so
// both produce bit-identical weights.
struct CoefBank {
    static constexpr bool kImm = false;
    __device__ static __forceinline__ float e(int, int) { return 0.f; } // (rows come through Top::bank)
    __device__ static __forceinline__ float o(int, int) { return 0.f; }
    __host__ __device__ static constexpr int deg(int, int dmax) { return dmax; }
};
template<int V>
struct CoefImm {
    static constexpr bool kImm = true;
    __device__ static __forceinline__ constexpr float e(int m, int i) { return imm::Table<V>::even(m, i); }
    __device__ static __forceinline__ constexpr float o(int m, int i) { return imm::Table<V>::odd(m, i); }
    __host__ __device__ static constexpr int deg(int m, int) { return imm::Table<V>::deg(m); }
};

#ifndef I3B_HOIST_TOP
#define I3B_HOIST_TOP 1
#endif
template<int K, int D, class Coef>
struct Weights {
    static constexpr int NE = D / 2 + 1;       // even coefficients  f^0, f^2, ...   (maximum degree)
    static constexpr int NO = (D + 1) / 2;     // odd coefficients   f^1, f^3, ...
    // per tap pair
    __host__ __device__ static constexpr int ne(int m) { return Coef::deg(m, D) / 2 + 1; }
    __host__ __device__ static constexpr int no(int m) { return (Coef::deg(m, D) + 1) / 2; }
    static constexpr bool kHoist = Coef::kImm && I3B_HOIST_TOP;
    // Leading coefficients.  An FFMA2 takes ONE immediate, so the first Horner step
    // (c_top * h + c_next) needs c_top in a register; left to itself ptxas re-creates these
    // registers every pulse with FMA-pipe moves (IMAD.MOV / HFMA2, 2 issue cycles each next
    // to FFMA2).  Loaded once per pulse tile and made opaque instead.
    struct Top {
        float e[(K + 1) / 2];
        float o[K / 2 > 0 ? K / 2 : 1];
        uint32_t bank; // general kernel: shared-memory address of the fitted rows
    };
    // `zero` is a kernel parameter that is always 0: adding it to the bit pattern keeps ptxas
    // from constant-propagating the value back into per-pulse immediate moves.
    __device__ static __forceinline__ void load_top(Top& t, int zero, uint32_t bank)
    {
        t.bank = bank;
        if (kHoist) {
#pragma unroll
            for (int m = 0; m < (K + 1) / 2; ++m)
                t.e[m] = __int_as_float(__float_as_int(Coef::e(m, ne(m) - 1)) + zero);
#pragma unroll
            for (int m = 0; m < K / 2; ++m)
                t.o[m] = __int_as_float(__float_as_int(Coef::o(m, no(m) - 1)) + zero);
        }
    }
    // even / odd parts of tap pair m at h = f * f (both pixels)
    template<int m>
    __device__ static __forceinline__ void parts(f32x2 h, f32x2& e, f32x2& o, const Top& top)
    {
        constexpr int ne_ = ne(m), no_ = no(m);
        float cev[4], cov[4];
        if (Coef::kImm) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                cev[i] = Coef::e(m, i);
                cov[i] = Coef::o(m, i);
            }
            if (kHoist) {
                cev[ne_ - 1] = top.e[m];
                if (2 * m != K - 1) cov[no_ - 1] = top.o[m < K / 2 ? m : 0];
            }
        } else {
            // broadcast LDS.128 (all lanes the same address): they issue in the FFMA2 shadow
            const float4 ce = lds128(top.bank + 32u * m);
            cev[0] = ce.x; cev[1] = ce.y; cev[2] = ce.z; cev[3] = ce.w;
            if (2 * m != K - 1) {
                const float4 co = lds128(top.bank + 32u * m + 16u);
                cov[0] = co.x; cov[1] = co.y; cov[2] = co.z; cov[3] = co.w;
            } else {
                cov[0] = cov[1] = cov[2] = cov[3] = 0.f;
            }
        }
        e = bcast2(cev[ne_ - 1]);
#pragma unroll
        for (int i = ne_ - 2; i >= 0; --i) e = fma2(e, h, bcast2(cev[i]));
        o = bcast2(cov[no_ - 1]);
        if (2 * m != K - 1) {
#pragma unroll
            for (int i = no_ - 2; i >= 0; --i) o = fma2(o, h, bcast2(cov[i]));
        }
    }
    // One tap pair: wlo = w_m(f), whi = w_{K-1-m}(f) of both pixels, from h = f*f, nf = -f.
    template<int m>
    __device__ static __forceinline__ void pair(f32x2 f, f32x2 h, f32x2 nf, f32x2& wlo, f32x2& whi,
                                                const Top& top)
    {
        f32x2 e, o;
        parts<m>(h, e, o, top);
        wlo = fma2(f, o, e);
        whi = fma2(nf, o, e);
    }
    // centre tap of an odd-length kernel (even polynomial only)
    __device__ static __forceinline__ f32x2 center(f32x2 h, const Top& top)
    {
        f32x2 e, o;
        parts<K / 2>(h, e, o, top);
        return e;
    }

    // Tap weights of TWO pixels at once: f = (f_pixel0, f_pixel1) in [-0.5, 0.5);
    // w[m] = (w_m(f0), w_m(f1)).  Coefficients enter as scalar-broadcast operands.
    template<int m = 0>
    __device__ static __forceinline__ void eval_from(f32x2 f, f32x2 h, f32x2 nf, f32x2 (&w)[K], const Top& top)
    {
        if constexpr (m < K / 2) {
            pair<m>(f, h, nf, w[m], w[K - 1 - m], top);
            eval_from<m + 1>(f, h, nf, w, top);
        } else if constexpr (K & 1) {
            w[K / 2] = center(h, top);
        }
    }
    __device__ static __forceinline__ void eval(f32x2 f, f32x2 (&w)[K], const Top& top)
    {
        const f32x2 h = mul2(f, f);
        const f32x2 nf = mul2(f, bcast2(-1.0f));
        eval_from<0>(f, h, nf, w, top);
    }
};

__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// Loop state of a thread's pixel PAIR (all FP32 / int: the FP64 geometry lives at segment
// boundaries only).  Packed fields hold (pixel 0, pixel 1) so the per-pulse phase / sample
// coordinate arithmetic of both pixels is one FFMA2 / FADD2 / FMUL2 each.
struct PairState {
    f32x2 ang0;        // 2*pi*(carrier phase in cycles at the segment base, reduced to [-1/2, 1/2])
    f32x2 f0m;         // frac(sample coordinate at the segment base) - 1/2 - Gr * ang0
    f32x2 c1, c2, c3;  // carrier phase increment over the segment: ((c3 j + c2) j + c1) j  [rad]
    f32x2 c1l;         // c1 = c1 + c1l to twice the precision (the term that reaches hundreds of rad)
    int i0rel[PX];     // window start at the base minus the magic bias (window origin added per tile)
    f32x2 accp[PX], accq[PX]; // FP32 partial sums of the pulse tile: sum cos*(re,im), sum sin*(re,im)
    float flim;        // steady sub-tiles: 1/2 - bound on what the cubic's curvature adds between checks
};

// Pulses per geometry segment (one exact FP64 phase evaluation per pixel per segment, a cubic
// in between): a run-time parameter, 128 or 64 -- the host takes 128 where the cubic's own
// error (0.0234 h^4 |d4 phase / dt4|, h = segment duration) stays below ~2e-6 rad (orbital
// geometries) and 64 otherwise (airborne: 1.4e-5 rad at 128).  FP32 evaluation error does not
// grow with the segment length any more (see RunPoly).  I3B_SEG forces a value (experiments).
#ifndef I3B_SEG
#define I3B_SEG 0
#endif
#ifndef I3B_SUB
#define I3B_SUB 8
#endif
constexpr int SUB = I3B_SUB;      // pulses per run (steady / per-pulse / aperture-edge path chosen per run)
constexpr int EDGE_RUN = SUB;
static_assert(TK % SUB == 0, "a staged pulse tile is a whole number of runs");
#ifndef I3B_KK_UNROLL
#define I3B_KK_UNROLL 4
#endif
constexpr int KK_UNROLL = I3B_KK_UNROLL;
static_assert(64 % TK == 0, "a geometry segment is a whole number of staged pulse tiles");
constexpr float MAGIC32 = 12582912.0f;         // 1.5 * 2^23: float -> nearest integer by addition
constexpr int MAGIC32_BITS = 0x4B400000;

// The per-pixel record is re-read at every geometry segment (64 pulses); with the L1
// evict-last hint the CTA's 20 KB of records stay in the small L1 left beside the shared
// memory carve-out instead of coming from L2 each time.
#ifndef I3B_PIX_L1
#define I3B_PIX_L1 1
#endif
__device__ __forceinline__ PixelRec load_pixel(const PixelRec* p)
{
#if I3B_PIX_L1
    PixelRec q;
    asm volatile("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(q.x) : "l"(&p->x));
    asm volatile("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(q.y) : "l"(&p->y));
    asm volatile("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(q.z) : "l"(&p->z));
    asm volatile("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(q.tau_atm) : "l"(&p->tau_atm));
    q.kstart = q.kstop = 0;
    return q;
#else
    return *p;
#endif
}

// exact carrier phase (cycles, incl. troposphere) of one pixel at one pulse, FP64
__device__ __forceinline__ double exact_cycles(const PixelRec& q, double xx, double t0cyc,
                                               const PulseRec& r)
{
    const double r2 = fma(q.x, r.m2px, fma(q.y, r.m2py, fma(q.z, r.m2pz, xx + r.pp)));
    const double sr = sqrt(r2);
    return fma(r.Cs, sr, fma(q.x, r.vBx, fma(q.y, r.vBy, fma(q.z, r.vBz, t0cyc + r.E))));
}

#ifndef I3B_MUFU_EARLY
#define I3B_MUFU_EARLY 0
#endif
// sin/cos on the SFU (MUFU after a range-reduction multiply).  The volatile form pins the
// issue point relative to the (volatile) shared-memory loads of the sample window.
__device__ __forceinline__ void sincos_fast(float x, float& sn, float& cs)
{
#if I3B_MUFU_EARLY
    asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(x));
    asm volatile("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(x));
#else
    sn = __sinf(x);
    cs = __cosf(x);
#endif
}

// Carrier rotation of the interpolated samples into the tile sums.  EDGE tiles skip pulses
// outside a pixel's aperture altogether (never 0 * sample: staged rows outside the aperture
// may hold anything, including NaN / Inf or rows whose upload is still in flight).
template<bool EDGE>
__device__ __forceinline__ void rotate_accumulate(PairState& S, f32x2 a0, f32x2 a1, float cs0, float sn0,
                                                  float cs1, float sn1, bool in0, bool in1)
{
    if (!EDGE || in0) {
        S.accp[0] = fma2(bcast2(cs0), a0, S.accp[0]);
        S.accq[0] = fma2(bcast2(sn0), a0, S.accq[0]);
    }
    if (!EDGE || in1) {
        S.accp[1] = fma2(bcast2(cs1), a1, S.accp[1]);
        S.accq[1] = fma2(bcast2(sn1), a1, S.accq[1]);
    }
}

// MAC of both pixels over the shared register window + carrier rotation into the tile sums.
// OFF = 0: window starts on an even sample, 1: odd.
template<int K, int OFF, bool EDGE, int NS>
__device__ __forceinline__ void mac_rotate(PairState& S, const f32x2 (&w)[K], const f32x2 (&sm)[NS],
                                           float cs0, float sn0, float cs1, float sn1, bool in0, bool in1)
{
    f32x2 a0 = 0ull, a1 = 0ull; // (re, im) of the interpolated sample, pixel 0 / 1
#pragma unroll
    for (int i = 0; i < K; ++i) {
        float w0, w1;
        unpack2(w[i], w0, w1);
        a0 = fma2(bcast2(w0), sm[i + OFF], a0);
        a1 = fma2(bcast2(w1), sm[i + OFF + 1], a1);
    }
    rotate_accumulate<EDGE>(S, a0, a1, cs0, sn0, cs1, sn1, in0, in1);
}

// Wide kernels (K = 16, 32): weights and window samples of all K taps do not fit the
// register file next to each other, so the taps are processed in chunks of 4 tap PAIRS
// (m, K-1-m share their even/odd polynomial parts): 8 weights, a 6-sample piece of the
// window at each end, 16 MACs -- then the next chunk, moving inwards from both ends.  The
// window pieces of consecutive chunks overlap by one LDS.128, which is carried in registers.
template<int K, int D, class Coef, int OFF, int C = 0>
struct ChunkedMac {
    typedef Weights<K, D, Coef> WT;
    __device__ static __forceinline__ void run(f32x2 f, f32x2 h, f32x2 nf, uint32_t src,
                                               f32x2 (&lo)[6], f32x2 (&hi)[6], f32x2& a0, f32x2& a1,
                                               const typename WT::Top& top)
    {
        constexpr int LO0 = 4 * C;         // first sample of the low piece  (even)
        constexpr int HI0 = K - 4 * C - 4; // first sample of the high piece (even)
        // low piece: samples [LO0, LO0 + 6); the first two came with the previous chunk
        if (C == 0) {
            const float4 v = lds128(src + 8u * LO0);
            lo[0] = pack2(v.x, v.y);
            lo[1] = pack2(v.z, v.w);
        }
        {
            const float4 v = lds128(src + 8u * (LO0 + 2)), u = lds128(src + 8u * (LO0 + 4));
            lo[2] = pack2(v.x, v.y);
            lo[3] = pack2(v.z, v.w);
            lo[4] = pack2(u.x, u.y);
            lo[5] = pack2(u.z, u.w);
        }
        // high piece: samples [HI0, HI0 + 6); the last two came with the previous chunk
        if (C == 0) {
            const float4 v = lds128(src + 8u * (HI0 + 4));
            hi[4] = pack2(v.x, v.y);
            hi[5] = pack2(v.z, v.w);
        }
        {
            const float4 v = lds128(src + 8u * HI0), u = lds128(src + 8u * (HI0 + 2));
            hi[0] = pack2(v.x, v.y);
            hi[1] = pack2(v.z, v.w);
            hi[2] = pack2(u.x, u.y);
            hi[3] = pack2(u.z, u.w);
        }
        f32x2 wl[4], wh[4];
        WT::template pair<4 * C + 0>(f, h, nf, wl[0], wh[0], top);
        WT::template pair<4 * C + 1>(f, h, nf, wl[1], wh[1], top);
        WT::template pair<4 * C + 2>(f, h, nf, wl[2], wh[2], top);
        WT::template pair<4 * C + 3>(f, h, nf, wl[3], wh[3], top);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float w0, w1;
            // tap 4C+i sits at sample OFF + 4C + i of the window, i.e. lo[OFF + i]
            unpack2(wl[i], w0, w1);
            a0 = fma2(bcast2(w0), lo[OFF + i], a0);
            a1 = fma2(bcast2(w1), lo[OFF + i + 1], a1);
            // tap K-1-(4C+i) sits at sample OFF + K-1-4C-i, i.e. hi[OFF + 3 - i]
            unpack2(wh[i], w0, w1);
            a0 = fma2(bcast2(w0), hi[OFF + 3 - i], a0);
            a1 = fma2(bcast2(w1), hi[OFF + 3 - i + 1], a1);
        }
        if constexpr (4 * (C + 1) < K / 2) {
            // next chunk: low piece moves up by 4 samples, high piece down by 4
            lo[0] = lo[4];
            lo[1] = lo[5];
            hi[4] = hi[0];
            hi[5] = hi[1];
            ChunkedMac<K, D, Coef, OFF, C + 1>::run(f, h, nf, src, lo, hi, a0, a1, top);
        }
    }
};

// Carrier phase of the pixel pair over ONE run of pulses, re-centred on the run's first pulse:
//   ang(x) = A0 + A1 x + A2 x^2 + A3 x^3,   x = position on the segment axis - that of the run start
// with A0 reduced modulo 2 pi.  Over a 128-pulse segment the unreduced phase reaches hundreds of
// radians, where an FP32 value resolves 3e-5 rad and the SFU's own range reduction another
// 2e-5: per-pulse errors that do not average out on noise-like scenes (relative RMS 6e-5 ...
// 8e-5 measured on the 80 MHz / airborne frames).  Here the one large term, c1 * j, is formed
// exactly (FMA two-product, c1 carried as a float pair), whole turns are removed (Cody-Waite
// split of 2 pi) and everything the pulse loop evaluates stays below ~40 rad.  `f0m` is the
// sample-coordinate offset adjusted for the removed turns (coordinate = f0m + Gr * ang).
struct RunPoly {
    f32x2 A0, A1, A2, A3, f0m;
};
constexpr float INV_TWO_PI_F = 0.15915494309189535f;
constexpr float TWO_PI_HI_F = 6.28125f;                  // 201 / 32: n * hi is exact
constexpr float TWO_PI_LO_F = 0.0019353071795864769f;    // 2 pi - hi

__device__ __forceinline__ RunPoly make_run_poly(const PairState& S, float js, float G)
{
    RunPoly R;
    const f32x2 js2 = bcast2(js);
    const f32x2 nq = mul2(S.c1, bcast2(-js));             // -(c1 js), rounded
    f32x2 e = fma2(S.c1, js2, nq);                        // c1 js + nq: the rounding error, exact
    e = fma2(S.c1l, js2, e);
    const f32x2 t = fma2(nq, bcast2(-INV_TWO_PI_F), bcast2(MAGIC32));
    const f32x2 n = add2(t, bcast2(-MAGIC32));            // whole turns in c1 js
    const f32x2 rneg = fma2(n, bcast2(TWO_PI_HI_F), nq);  // -(c1 js - n 2pi_hi)
    f32x2 rest = fma2(fma2(S.c3, js2, S.c2), mul2(js2, js2), add2(S.ang0, e));
    rest = fma2(n, bcast2(-TWO_PI_LO_F), rest);
    R.A0 = sub2(rest, rneg);
    const f32x2 c3x3 = mul2(S.c3, bcast2(3.0f)), c2x2 = add2(S.c2, S.c2);
    R.A2 = fma2(c3x3, js2, S.c2);
    R.A1 = fma2(fma2(c3x3, js2, c2x2), js2, S.c1);
    R.A3 = S.c3;
    R.f0m = fma2(n, bcast2(G), S.f0m);                    // Gr * 2 pi n = G n samples
    return R;
}

// One staged pulse tile (TK pulses) for the thread's pixel pair.  EDGE = false: every pixel
// of the CTA integrates every pulse of the tile (no aperture test in the loop).
//
// Issue-cost model measured on B200 (scripts/ubench/pipes2.cu, 4 warps per scheduler): FFMA2 /
// FADD2 / FMUL2 2.1 cycles, scalar FFMA / IMAD / FMUL 2.0 (no cheaper than the packed form),
// FADD 1.4, LOP3 / IADD3 / VIMNMX 0.6-0.9, MOV / LDS ~0.2 (issue in the FFMA2 shadow), MUFU 8
// XU cycles (asynchronous).  So the loop is written to be FFMA2-only on the FMA pipe.
template<int K, int D, class Coef, bool EDGE, int NP>
__device__ __forceinline__ void tile_body(PairState& S, const RunPoly& R, float js, unsigned& jjmax,
                                          uint32_t lines_addr, uint32_t row_bytes, int wlo,
                                          unsigned jmax, float Gr, unsigned krel0, unsigned krel1,
                                          int zero, uint32_t bank, const float* __restrict__ xi = nullptr,
                                          unsigned kspan0 = 0u, unsigned kspan1 = 0u)
{
    // krel / kspan (EDGE only): pulse index of the run's first pulse relative to each pixel's
    // aperture start, and the aperture lengths -- pulses outside are skipped.
    // xi (non-uniform pulse times only): position of each pulse of the run on the segment's
    // time axis, in nominal pulse intervals (js = that of the run's first pulse) -- replaces
    // the pulse index in the phase polynomial
    float xf = 0.f;
    typedef Weights<K, D, Coef> WT;
    typename WT::Top top;
    WT::load_top(top, zero, bank);
    const int iw0 = S.i0rel[0] - wlo, iw1 = S.i0rel[1] - wlo;
    uint32_t line_addr = lines_addr;
    // (wide kernels: the body is already hundreds of instructions per pulse)
    constexpr int kUnroll = (K >= 16) ? 1 : KK_UNROLL;
#pragma unroll kUnroll
    for (int kk = 0; kk < NP; ++kk) {
        // keep the staged line address a loop-carried register (ptxas otherwise rebuilds it
        // from the shared-memory base every pulse: ~10 instructions)
        asm volatile("" : "+r"(line_addr));
        // carrier phase [rad] (reduced at the segment base) of both pixels at once
        if (xi) xf = __ldg(xi + kk) - js;
        const f32x2 x2 = bcast2(xf);
        const f32x2 ang = fma2(fma2(fma2(R.A3, x2, R.A2), x2, R.A1), x2, R.A0);
        const f32x2 g = fma2(ang, bcast2(Gr), R.f0m); // sample coordinate - floor(base) - 1/2
        const f32x2 m = add2(g, bcast2(MAGIC32));     // nearest integer of g == floor(coordinate)
        const f32x2 f = sub2(g, add2(m, bcast2(-MAGIC32))); // centred fraction in [-1/2, 1/2]
        xf += 1.0f;
        float m0, m1, ang0, ang1;
        unpack2(m, m0, m1);
        unpack2(ang, ang0, ang1);
        const unsigned jj0 = (unsigned) (iw0 + __float_as_int(m0));
        const unsigned jj1 = (unsigned) (iw1 + __float_as_int(m1));
        jjmax = max(jjmax, max(jj0, jj1));
        const unsigned j0 = min(jj0, jmax), j1 = min(jj1, jmax);
        float cs0, sn0, cs1, sn1;
        // pulses outside a pixel's aperture (EDGE tiles only) are skipped at the rotation
        const bool in0 = !EDGE || krel0 + (unsigned) kk < kspan0;
        const bool in1 = !EDGE || krel1 + (unsigned) kk < kspan1;
        sincos_fast(ang0, sn0, cs0);
        sincos_fast(ang1, sn1, cs1);
        if constexpr (K >= 16 && K % 8 == 0) {
            if (j1 == j0 + 1) {
                const uint32_t src = line_addr + ((j0 >> 1) << 4);
                const f32x2 h = mul2(f, f), nf = mul2(f, bcast2(-1.0f));
                f32x2 lo[6], hi[6], a0 = 0ull, a1 = 0ull;
                if (j0 & 1u) ChunkedMac<K, D, Coef, 1>::run(f, h, nf, src, lo, hi, a0, a1, top);
                else ChunkedMac<K, D, Coef, 0>::run(f, h, nf, src, lo, hi, a0, a1, top);
                rotate_accumulate<EDGE>(S, a0, a1, cs0, sn0, cs1, sn1, in0, in1);
                line_addr += row_bytes;
                continue;
            }
        }
        f32x2 w[K];
        if (j1 == j0 + 1) {
            // shared register window: K+1 samples (+1 when the start is odd); the loads are
            // issued BEFORE the weight polynomials so their latency hides behind them
            constexpr int NV = (K + 3) / 2; // 16-byte loads covering K+2 samples
            f32x2 sm[2 * NV];
            const uint32_t src = line_addr + ((j0 >> 1) << 4);
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const float4 v4 = lds128(src + 16u * i);
                sm[2 * i] = pack2(v4.x, v4.y);
                sm[2 * i + 1] = pack2(v4.z, v4.w);
            }
            WT::eval(f, w, top);
            if (j0 & 1u) mac_rotate<K, 1, EDGE>(S, w, sm, cs0, sn0, cs1, sn1, in0, in1);
            else mac_rotate<K, 0, EDGE>(S, w, sm, cs0, sn0, cs1, sn1, in0, in1);
        } else {
            // general spacing: independent windows
            WT::eval(f, w, top);
            const uint32_t s0 = line_addr + j0 * (uint32_t) sizeof(float2);
            const uint32_t s1 = line_addr + j1 * (uint32_t) sizeof(float2);
            f32x2 a0 = 0ull, a1 = 0ull;
#pragma unroll
            for (int i = 0; i < K; ++i) {
                float w0, w1;
                unpack2(w[i], w0, w1);
                a0 = fma2(bcast2(w0), lds64(s0 + 8u * i), a0);
                a1 = fma2(bcast2(w1), lds64(s1 + 8u * i), a1);
            }
            rotate_accumulate<EDGE>(S, a0, a1, cs0, sn0, cs1, sn1, in0, in1);
        }
        line_addr += row_bytes;
    }
}

// Aperture-edge tiles (a few per pixel) go through a real call: kept out of line, their
// aperture tests do not take registers away from the interior loop, which is >95 % of the work.
#ifndef I3B_EDGE_NOINLINE
#define I3B_EDGE_NOINLINE 1
#endif
template<int K, int D, class Coef>
#if I3B_EDGE_NOINLINE
__device__ __noinline__
#else
__device__ __forceinline__
#endif
unsigned tile_body_edge(PairState& S, RunPoly R, float js, unsigned jjmax, uint32_t lines_addr,
                        uint32_t row_bytes, int wlo, unsigned jmax, float Gr, unsigned krel0,
                        unsigned krel1, int zero, uint32_t bank, const float* __restrict__ xi,
                        unsigned kspan0, unsigned kspan1)
{
    // (the run polynomial and jjmax by value: only the pair state has to live in memory around the call)
    tile_body<K, D, Coef, true, EDGE_RUN>(S, R, js, jjmax, lines_addr, row_bytes, wlo, jmax, Gr, krel0, krel1, zero, bank, xi,
                                          kspan0, kspan1);
    return jjmax;
}

// Tap pair by tap pair: the two weights of a pair are consumed by their four MACs right
// away, so only 4 weight registers are live at a time instead of 2 K (source order for ptxas:
// leaves room to hoist the next pulse's window loads).
#ifndef I3B_PAIRWISE_MAC
#define I3B_PAIRWISE_MAC 0
#endif
template<int K, int D, class Coef, int OFF, int M = 0>
struct PairwiseMac {
    typedef Weights<K, D, Coef> WT;
    template<int NS>
    __device__ static __forceinline__ void run(f32x2 f, f32x2 h, f32x2 nf, const f32x2 (&sm)[NS],
                                               f32x2& a0, f32x2& a1, const typename WT::Top& top)
    {
        if constexpr (M < K / 2) {
            f32x2 wl, wh;
            WT::template pair<M>(f, h, nf, wl, wh, top);
            float w0, w1;
            unpack2(wl, w0, w1);
            a0 = fma2(bcast2(w0), sm[M + OFF], a0);
            a1 = fma2(bcast2(w1), sm[M + OFF + 1], a1);
            unpack2(wh, w0, w1);
            a0 = fma2(bcast2(w0), sm[K - 1 - M + OFF], a0);
            a1 = fma2(bcast2(w1), sm[K - M + OFF], a1);
            PairwiseMac<K, D, Coef, OFF, M + 1>::run(f, h, nf, sm, a0, a1, top);
        } else if constexpr (K & 1) {
            const f32x2 wc = WT::center(h, top);
            float w0, w1;
            unpack2(wc, w0, w1);
            a0 = fma2(bcast2(w0), sm[K / 2 + OFF], a0);
            a1 = fma2(bcast2(w1), sm[K / 2 + OFF + 1], a1);
        }
    }
};

// ---- steady sub-tiles --------------------------------------------------------------------
// Over a handful of pulses the sample coordinate of a pixel moves by a small fraction of a
// sample (range migration is <= ~1e-2 samples per pulse), so for most runs of I3B_SUB pulses
// the INTEGER part of the coordinate -- the window position -- is the same for every pulse
// of the run and the two pixels of a thread keep adjacent windows.  Such a run ("steady
// sub-tile", proven before it starts from the phase cubic itself: values at both ends plus
// a curvature bound) needs no per-pulse rounding, index arithmetic, clamping or branching:
// the window address advances by one staged row per pulse, the parity of the window start
// is a compile-time constant of the loop, and the fraction is ONE FFMA2 from the phase.
// The loop is straight-line code (fully unrolled, pulse offsets are immediates), which also
// lets ptxas overlap the phase / sincos / loads of the next pulses with the FMA work of the
// current one.  Runs that are not steady take tile_body (per-pulse rounding), bit-for-bit
// the same arithmetic as before.
#ifndef I3B_STEADY
#define I3B_STEADY 1
#endif
#ifndef I3B_STEADY_MAX_TAPS
#define I3B_STEADY_MAX_TAPS 32
#endif
// Inside a steady run the phase is the QUADRATIC through the cubic's values at the first,
// middle and last pulse of the run (one FFMA2 less per pulse); the cubic term it drops is
// |c3| * 0.048 (SUB-1)^3 rad at most -- ~1e-8 rad for orbital and airborne geometries -- and
// segments where that bound exceeds 2e-6 rad take the per-pulse path with the full cubic.
#ifndef I3B_QUAD_RUN
#define I3B_QUAD_RUN 1
#endif

template<int K, int D, class Coef, int OFF, int NP>
__device__ __forceinline__ void subtile_steady(PairState& S, f32x2 A0, f32x2 A1, f32x2 A2, f32x2 A3,
                                               f32x2 fbase, float Gr, uint32_t src, uint32_t row_bytes,
                                               int zero, uint32_t bank)
{
    typedef Weights<K, D, Coef> WT;
    typename WT::Top top;
    WT::load_top(top, zero, bank);
    // narrow kernels: fully unrolled (pulse offsets are immediates); wide ones (hundreds of
    // instructions per pulse already) keep a rolled loop with a running offset
    constexpr int kUnroll = (K >= 16) ? 1 : NP;
    float xf = 0.f;
#pragma unroll kUnroll
    for (int x = 0; x < NP; ++x) {
        // carrier phase of both pixels x pulses into the run (cubic re-centred on the run)
        f32x2 ang = A0;
        if constexpr (K >= 16) {
            const f32x2 X = bcast2(xf);
            ang = I3B_QUAD_RUN ? fma2(fma2(A2, X, A1), X, A0) : fma2(fma2(fma2(A3, X, A2), X, A1), X, A0);
            xf += 1.0f;
        } else if (x > 0) {
            const f32x2 X = bcast2((float) x);
            ang = I3B_QUAD_RUN ? fma2(fma2(A2, X, A1), X, A0) : fma2(fma2(fma2(A3, X, A2), X, A1), X, A0);
        }
        const f32x2 f = fma2(ang, bcast2(Gr), fbase); // centred fraction (integer part is fixed)
        float ang0, ang1, cs0, sn0, cs1, sn1;
        unpack2(ang, ang0, ang1);
        sincos_fast(ang0, sn0, cs0);
        sincos_fast(ang1, sn1, cs1);
        if constexpr (K >= 16 && K % 8 == 0) {
            const f32x2 h = mul2(f, f), nf = mul2(f, bcast2(-1.0f));
            f32x2 lo[6], hi[6], a0 = 0ull, a1 = 0ull;
            ChunkedMac<K, D, Coef, OFF>::run(f, h, nf, src, lo, hi, a0, a1, top);
            rotate_accumulate<false>(S, a0, a1, cs0, sn0, cs1, sn1, true, true);
        } else {
            constexpr int NV = (K + 3) / 2;
            f32x2 sm[2 * NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const float4 v4 = lds128(src + 16u * i);
                sm[2 * i] = pack2(v4.x, v4.y);
                sm[2 * i + 1] = pack2(v4.z, v4.w);
            }
#if I3B_PAIRWISE_MAC
            const f32x2 h = mul2(f, f), nf = mul2(f, bcast2(-1.0f));
            f32x2 a0 = 0ull, a1 = 0ull;
            PairwiseMac<K, D, Coef, OFF>::run(f, h, nf, sm, a0, a1, top);
            rotate_accumulate<false>(S, a0, a1, cs0, sn0, cs1, sn1, true, true);
#else
            f32x2 w[K];
            WT::eval(f, w, top);
            mac_rotate<K, OFF, false>(S, w, sm, cs0, sn0, cs1, sn1, true, true);
#endif
        }
        src += row_bytes;
    }
}

template<int V>
struct NonUniformTag {
    static constexpr int value = V;
};
// run loop: carry the run's position / staged-row address across runs; separate instance of the
// loop for non-uniform pulse trains
#ifndef I3B_RUN_CARRY
#define I3B_RUN_CARRY 1
#endif
#ifndef I3B_NU_SPLIT
#define I3B_NU_SPLIT 0 // (measured: the second instance of the loop costs 2 % -- instruction cache)
#endif
// I3B_PAIR_RUNS = 1: a steady run that stays steady over the whole 16-pulse tile is proven once
// and the tile's second run reuses the (re-centred) phase quadratic, window and fraction base of
// the first -- 60 instructions fewer per tile, and measured SLOWER (9 taps: 0.731 against 0.744
// of FP32 peak; 8 taps airborne 0.54 against 0.58): the run set-up issues in the shadow of other
// warps' FFMA2 work, what the chain adds is a longer dependent path into the second body.
#ifndef I3B_PAIR_RUNS
#define I3B_PAIR_RUNS 0
#endif
// interior runs that are not steady through the out-of-line per-pulse body: measured slower
// (0.734 against 0.744 at 9 taps, 0.52 against 0.58 on the airborne frame, where an eighth of the
// pulses sit in runs that cross a sample boundary) -- the call moves the pair state through memory
#ifndef I3B_NONSTEADY_CALL
#define I3B_NONSTEADY_CALL 0
#endif

template<int K, int D, class Coef>
__global__ void __launch_bounds__(NTHREADS, 2)
accumulate_fast_kernel(const __grid_constant__ CUtensorMap rc_map, const __grid_constant__ PolyTable poly,
                       FastParams P,
                       const PixelRec* __restrict__ pix, const PulseRec* __restrict__ pulse,
                       double2* __restrict__ acc, const TileInfo* __restrict__ tiles,
                       DevStatus* status)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmemHeader* hdr = reinterpret_cast<SmemHeader*>(smem_raw);
    unsigned char* stage0 = smem_raw + HEADER_BYTES;
    const size_t sbytes = stage_bytes(P.W);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    // Tile order: azimuth-major inside groups of GROUP_RG range columns.  CTAs are started
    // in blockIdx order, so the ~300 co-resident CTAs are ~150 consecutive azimuth tiles of
    // the same 2 range columns: they stream the SAME pulse lines a few pulses apart and share
    // them in L2.  (Range-major order made every wave of CTAs re-read its whole aperture
    // window from HBM: 598 GB of DRAM reads per C2 frame, ncu r01 v7.)
    int tile_i, tile_j;
    {
        const int per_group = P.tiles_az * GROUP_RG;
        const int grp = blockIdx.x / per_group;
        const int rem = blockIdx.x - grp * per_group;
        const int gcols = min(GROUP_RG, P.tiles_rg - grp * GROUP_RG);
        tile_j = rem / gcols;
        tile_i = grp * GROUP_RG + rem - tile_j * gcols;
        tile_j += P.tile_j0;
    }
    const int tile_id = tile_j * P.tiles_rg + tile_i; // row-major id (target-solve tile table)
    const int col0 = tile_i * TILE_RG, line0 = tile_j * TILE_AZ;
    {
        // nothing of this tile in the launch's pulse range (or a failed pixel: generic kernel's
        // job): leave before touching the pixel records
        const TileInfo ti = tiles[tile_id];
        if (ti.bad || ti.kmax <= P.k_begin || ti.kmin >= P.k_end || ti.kmin >= ti.kmax) return;
        if (P.k_landed > 0 && min(ti.kmax, P.k_end) > P.k_landed) {
            // the host's bound on this row's aperture did not hold: nothing of the tile is
            // integrated, the call is redone once every pulse is on the device
            if (threadIdx.x == 0) status->premature = 1;
            return;
        }
    }

    constexpr int LOWOFF = (K & 1) ? -(K / 2) : 1 - K / 2;
    constexpr double SHIFT = (K & 1) ? 0.5 : 0.0;

    // general kernel: the fitted polynomial rows, from the kernel parameter to shared memory
    const uint32_t poly_addr = smem_u32(smem_raw + POLY_OFFSET);
    if (!Coef::kImm) {
        constexpr int NF = (int) (sizeof(PolyTable) / sizeof(float));
        static_assert(POLY_OFFSET + sizeof(PolyTable) <= HEADER_BYTES, "polynomial rows overflow the header");
        const float* src = reinterpret_cast<const float*>(&poly);
        float* dst = reinterpret_cast<float*>(smem_raw + POLY_OFFSET);
        for (int i = tid; i < NF; i += NTHREADS) dst[i] = src[i];
    }
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(&hdr->full[s], 1);
            mbar_init(&hdr->empty[s], NTHREADS / 32);
            hdr->done[s] = 0;
        }
        hdr->kb = INT_MAX;
        hdr->ke = INT_MIN;
        hdr->ks_max = INT_MIN;
        hdr->ke_min = INT_MAX;
        hdr->bad = 0;
        fence_mbar_init();
    }
    __syncthreads();

    // ---- prologue: pixel records, CTA pulse range, corner positions ----------------
    // Threads past the grid edge shadow the nearest in-grid pixel (same aperture, nothing
    // written back), so they never force the aperture test on the rest of the CTA.
    PairState S;
    long long gidx[PX];
    const int lrow = tid / (TILE_RG / PX);
    const int lcol = (tid % (TILE_RG / PX)) * PX;
    {
        const int last_row = min(TILE_AZ, P.out_lines - line0) - 1;
        const int last_col = min(TILE_RG, P.out_width - col0) - 1;
        int kmin = INT_MAX, kmax = INT_MIN, ksmax = INT_MIN, kemin = INT_MAX;
        bool bad = false;
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            const int jj = line0 + lrow, ii = col0 + lcol + p;
            const int jc = min(jj, P.out_lines - 1), ic = min(ii, P.out_width - 1);
            gidx[p] = (long long) jc * P.out_width + ic;
            const PixelRec r = pix[gidx[p]];
            if (r.kstart < 0) bad = true;
            const int ks = max(r.kstart, P.k_begin);
            const int ke_p = min(r.kstop, P.k_end);
            S.accp[p] = S.accq[p] = 0ull;
            if (ke_p > ks) {
                kmin = min(kmin, ks);
                kmax = max(kmax, ke_p);
            }
            ksmax = max(ksmax, ks);
            kemin = min(kemin, ke_p);
            // the four corner pixels of the (grid-clipped) tile publish their position
            const int lc = lcol + p;
            const bool top = lrow == 0, bot = lrow == last_row;
            const bool lef = lc == 0, rig = lc == last_col;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const bool rsel = (c & 2) ? bot : top, csel = (c & 1) ? rig : lef;
                if (rsel && csel) {
                    hdr->corner[c][0] = r.x; hdr->corner[c][1] = r.y; hdr->corner[c][2] = r.z;
                    hdr->corner[c][3] = P.fc * r.tau_atm;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
            kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
            ksmax = max(ksmax, __shfl_xor_sync(0xffffffffu, ksmax, o));
            kemin = min(kemin, __shfl_xor_sync(0xffffffffu, kemin, o));
        }
        const bool anybad = __any_sync(0xffffffffu, bad);
        if (lane == 0) {
            if (kmin != INT_MAX) atomicMin(&hdr->kb, kmin);
            if (kmax != INT_MIN) atomicMax(&hdr->ke, kmax);
            atomicMax(&hdr->ks_max, ksmax);
            atomicMin(&hdr->ke_min, kemin);
            if (anybad) hdr->bad = 1;
        }
    }
    __syncthreads();
    const int kb = hdr->kb, ke = hdr->ke;
    const int ks_max = hdr->ks_max, ke_min = hdr->ke_min;
    if (hdr->bad) return; // (tile table says the same)
    if (kb >= ke) return; // nothing to integrate in this launch
    // Pulse tiles and geometry segments sit on ABSOLUTE pulse indices (multiples of TK / SEG),
    // not on the first pulse this CTA happens to integrate in this launch: every FP32 tile
    // sum, every phase cubic and the order of the FP64 additions are then properties of the
    // pixel alone, so the image does not depend on how the pulses were cut into launches
    // (`batch`, upload timing) nor on how the grid was cut into shards.  The host keeps
    // launch boundaries on multiples of TK (fast_pulse_tile()).
    const int SEG = P.seg;
    const int t0 = (kb / TK) * TK; // kb >= 0
    const int ntiles = (ke - t0 + TK - 1) / TK;

    // Producer role (one warp, all lanes converge here): stage pulse tile n.
    auto produce = [&](int n) {
        const int s = n % NSTAGE;
        if (n >= NSTAGE) mbar_wait(&hdr->empty[s], ((n / NSTAGE) - 1) & 1);
        const int kfirst = t0 + n * TK;
        const int klast = kfirst + TK - 1; // the pulse table is padded past the last pulse
        // range window from the 4 corner pixels at the first/last pulse of the tile
        double u = 0.;
        if (lane < 8) {
            const PulseRec r = pulse[(lane < 4) ? kfirst : klast];
            u = sample_coord(hdr->corner[lane & 3], r, P.G, P.U0);
        }
        double umin = (lane < 8) ? u : 1e300, umax = (lane < 8) ? u : -1e300;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            umin = fmin(umin, __shfl_xor_sync(0xffffffffu, umin, o));
            umax = fmax(umax, __shfl_xor_sync(0xffffffffu, umax, o));
        }
        if (lane == 0) {
            int wlo = (int) floor(umin) - K / 2 - 3;
            wlo &= ~1; // even: keeps LDS.128 parity == sample-index parity
            const int whi = (int) ceil(umax) + K / 2 + 6;
            if (!(umin == umin) || whi - wlo > P.W) status->window_overflow = 1;
            hdr->winlo[s] = wlo;
            unsigned char* sp = stage0 + (size_t) s * sbytes;
            mbar_arrive_expect_tx(&hdr->full[s], (uint32_t) ((size_t) TK * P.W * sizeof(float2)));
            // rows before / after the staged swath (kfirst < rc_k0, ...) arrive as zeros
            tma_load_2d(sp, &rc_map, wlo, kfirst - P.rc_k0, &hdr->full[s]);
        }
        __syncwarp();
    };
    if (warp == 0) {
        for (int n = 0; n < (I3B_LAST_ARRIVER ? NSTAGE : PREFETCH) && n < ntiles; ++n) produce(n);
    }

    // Per-thread shared-memory slots: FP64 running sums (2 per pixel) and the exact carrier
    // phase at the four segment boundaries around the current segment (4 per pixel).
    double* slot = reinterpret_cast<double*>(smem_raw + HEADER_BYTES + NSTAGE * sbytes) + tid * (6 * PX);
    double* accd = slot;          // [2 * PX]
    double* ybnd = slot + 2 * PX; // [PX][4]  T(b - SEG), T(b), T(b + SEG), T(b + 2 SEG)
    // The FP64 sums CONTINUE from what earlier launches left in `acc`: one sequential chain
    // of additions per pixel whatever the launch partition (bit-reproducible image).
#pragma unroll
    for (int p = 0; p < PX; ++p) {
        const double2 a = acc[gidx[p]];
        accd[2 * p] = a.x;
        accd[2 * p + 1] = a.y;
    }
    // segment boundaries are multiples of SEG; evaluate T at the three around the first segment
    // (the fourth comes with the segment)
    {
        const int b0 = t0 & ~(SEG - 1);
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            const PixelRec q = load_pixel(pix + gidx[p]);
            const double xx = q.x * q.x + q.y * q.y + q.z * q.z, tc = P.fc * q.tau_atm;
            ybnd[4 * p + 1] = exact_cycles(q, xx, tc, pulse[b0 - SEG]);
            ybnd[4 * p + 2] = exact_cycles(q, xx, tc, pulse[b0]);
            ybnd[4 * p + 3] = exact_cycles(q, xx, tc, pulse[b0 + SEG]);
        }
    }

    unsigned jjmax = 0; // sticky maximum of the (unsigned) window offsets: overflow detector
    const uint32_t stage_addr0 = smem_u32(stage0);
    const uint32_t row_bytes = (uint32_t) P.W * (uint32_t) sizeof(float2);
    const unsigned jmax = (unsigned) (P.W - (K + 3));
    const double TWO_PI_D = 6.283185307179586476925;
    const float Gr = P.Grf;   // samples per radian of carrier phase
    const float Gsamp = P.Gf; // samples per cycle (turn) of carrier phase
    int seg_b = 0;                              // first pulse of the current segment

    for (int n = 0; n < ntiles; ++n) {
        if (!I3B_LAST_ARRIVER && n + PREFETCH < ntiles &&
            warp == (I3B_ROTATE_PRODUCER ? (n + PREFETCH) % NWARPS_ROT : 0))
            produce(n + PREFETCH);

        const int kt = t0 + n * TK; // first pulse of the tile
        if (n == 0 || (kt & (SEG - 1)) == 0) {
            // ---- new geometry segment: one exact FP64 evaluation per pixel, cubic through
            // the four surrounding boundaries, everything inside the segment is FP32 ----
            const int b = kt & ~(SEG - 1);
            seg_b = b;
            float c1[PX], c1lo[PX], c2[PX], c3[PX], a0[PX], f0[PX];
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                const PixelRec q = load_pixel(pix + gidx[p]);
                const double xx = q.x * q.x + q.y * q.y + q.z * q.z, tc = P.fc * q.tau_atm;
                const double y0 = ybnd[4 * p + 1], y1 = ybnd[4 * p + 2], y2 = ybnd[4 * p + 3];
                const double y3 = exact_cycles(q, xx, tc, pulse[b + 2 * SEG]);
                ybnd[4 * p + 1] = y1;
                ybnd[4 * p + 2] = y2;
                ybnd[4 * p + 3] = y3;
                if (P.tn) {
                    // Non-uniform pulse times: the phase is smooth in TIME.  Cubic through the four
                    // boundary values at their actual positions x on the segment's time axis
                    // (nominal pulse intervals from the segment base; Newton form on the nodes
                    // 0, x2, x0, x3, expanded to monomials); the pulse loop evaluates it at xi[k].
                    const double x0 = P.tn[b - SEG] - P.tn[b], x2 = P.tn[b + SEG] - P.tn[b];
                    const double x3 = P.tn[b + 2 * SEG] - P.tn[b];
                    const double d12 = (y2 - y1) / x2;
                    const double d20 = (y0 - y2) / (x0 - x2);
                    const double d120 = (d20 - d12) / x0;
                    const double d03 = (y3 - y0) / (x3 - x0);
                    const double d203 = (d03 - d20) / (x3 - x2);
                    const double d1203 = (d203 - d120) / x3;
                    const double c1d = TWO_PI_D * (d12 - d120 * x2 + d1203 * x2 * x0);
                    c1[p] = (float) c1d;
                    c1lo[p] = (float) (c1d - (double) c1[p]);
                    c2[p] = (float) (TWO_PI_D * (d120 - d1203 * (x2 + x0)));
                    c3[p] = (float) (TWO_PI_D * d1203);
                } else {
                const double d1 = y2 - y1, d2 = (y2 - y1) - (y1 - y0);
                const double d3 = ((y3 - y2) - (y2 - y1)) - d2;
                // p(tau) - y1 = tau (d1 - d2/2 - d3/6) + tau^2 d2/2 + tau^3 d3/6, tau = j / SEG
                const double c1d = TWO_PI_D * (d1 - 0.5 * d2 - d3 * (1.0 / 6.0)) * (1.0 / SEG);
                c1[p] = (float) c1d;
                c1lo[p] = (float) (c1d - (double) c1[p]);
                c2[p] = (float) (TWO_PI_D * (0.5 * d2) * (1.0 / ((double) SEG * SEG)));
                c3[p] = (float) (TWO_PI_D * (d3 * (1.0 / 6.0)) * (1.0 / ((double) SEG * SEG * SEG)));
                }
                a0[p] = (float) (TWO_PI_D * (y1 - rint(y1)));
                const double uh = fma(y1, P.G, SHIFT - P.U0);
                const double ufl = floor(uh);
                f0[p] = (float) (uh - ufl) - 0.5f;
                S.i0rel[p] = (int) ufl + LOWOFF - MAGIC32_BITS; // window origin added per tile
            }
            // Guard on the phase rate: a run evaluates A0 + A1 x + ... with |A0| <= pi and x < SUB
            // in FP32, good to ~6e-8 |A1| SUB rad.  Beyond 1e3 rad per run (a Doppler centroid of
            // ~20 PRFs; nothing the workflow produces) the call is handed to the generic kernel
            // through the same flag as a window overflow.
            if (fmaxf(fabsf(c1[0]), fabsf(c1[1])) * (float) SUB > 1.0e3f) jjmax = 0xffffffffu;
            S.c1 = pack2(c1[0], c1[1]);
            S.c1l = pack2(c1lo[0], c1lo[1]);
            S.c2 = pack2(c2[0], c2[1]);
            S.c3 = pack2(c3[0], c3[1]);
            S.ang0 = pack2(a0[0], a0[1]);
            // the loop evaluates coordinate = f0m + Gr * (ang0 + increment)
            S.f0m = pack2(f0[0] - Gr * a0[0], f0[1] - Gr * a0[1]);
            // what the cubic's curvature can add to the coordinate between the two ends of a
            // steady run: |u''| (SUB-1)^2 / 8, u'' = Gr (2 c2 + 6 c3 j), j < SEG
            const float curv0 = 2.f * fabsf(c2[0]) + (6.f * SEG) * fabsf(c3[0]);
            const float curv1 = 2.f * fabsf(c2[1]) + (6.f * SEG) * fabsf(c3[1]);
            S.flim = 0.5f - 1e-5f - fabsf(Gr) * fmaxf(curv0, curv1) * ((SUB - 1) * (SUB - 1) / 8.0f);
            if (I3B_QUAD_RUN) {
                constexpr int kSpan = I3B_PAIR_RUNS ? TK - 1 : SUB - 1; // longest span a quadratic covers
                constexpr float kDev = 0.0481125f * kSpan * kSpan * kSpan;
                if (fmaxf(fabsf(c3[0]), fabsf(c3[1])) * kDev > 2e-6f) S.flim = -1.0f;
            }
            if (P.xi) S.flim = -1.0f; // steady runs step the phase by whole pulse indices
        }

        const int s = n % NSTAGE;
        mbar_wait(&hdr->full[s], (n / NSTAGE) & 1);
        const uint32_t lines_addr = stage_addr0 + (uint32_t) s * (uint32_t) sbytes;
        const int wlo = hdr->winlo[s];
        // The tile is worked through in runs of SUB pulses, each taking the cheapest path it
        // qualifies for: every pulse inside every pixel's aperture (almost all runs) -> steady
        // run if the window position provably does not move, else per-pulse rounding; runs that
        // straddle an aperture edge -> the out-of-line edge path (skips pulses outside).
        const int iw0 = S.i0rel[0] - wlo, iw1 = S.i0rel[1] - wlo;
        // position of each run's first pulse on the segment axis: the pulse index, or (non-uniform
        // pulse trains) the time in nominal pulse intervals.  Carried across the runs of the tile
        // together with the staged-row address, so a run starts from registers.
        // (NU: non-uniform pulse train -- positions come from the xi table and no run is steady;
        // the uniform instance of the loop never looks at the table)
        auto run_tile = [&](auto nu_tag) {
        constexpr int NUV = decltype(nu_tag)::value; // 1: non-uniform, 0: uniform, -1: decided per run
        constexpr bool NU = NUV == 1;
        // a steady run proven for the whole 16-pulse tile hands its quadratic to the second run
        bool chain = false;
        f32x2 cA0 = 0ull, cA1 = 0ull, cA2 = 0ull, cfb = 0ull;
        uint32_t csrc = 0u;
        unsigned cpar = 0u;
#if I3B_RUN_CARRY
        float js = (float) (kt - seg_b);
        uint32_t la = lines_addr;
        int kr = kt; // first pulse of the run
#pragma unroll 1
        for (int sub = 0; sub < TK / SUB; ++sub, kr += SUB, js += (float) SUB, la += (uint32_t) SUB * row_bytes) {
            asm volatile("" : "+r"(la), "+f"(js));
#else
#pragma unroll 1
        for (int sub = 0; sub < TK / SUB; ++sub) {
            const int kr = kt + sub * SUB;
            float js = (float) (kr - seg_b);
            const uint32_t la = lines_addr + (uint32_t) (sub * SUB) * row_bytes;
#endif
            const float* xi_run = nullptr;
            if constexpr (NUV == 1) {
                xi_run = P.xi + kr;
                js = __ldg(xi_run);
            } else if constexpr (NUV == -1) {
                if (P.xi) {
                    xi_run = P.xi + kr;
                    js = __ldg(xi_run);
                }
            }
            // ---- what this run is: second half of a proven 16-pulse steady run (chained), steady,
            // per-pulse, or aperture edge ----
            RunPoly R{};
            bool steady = false;
            f32x2 sA0 = 0ull, sA1 = 0ull, sA2 = 0ull, sfb = 0ull; // steady run: phase quadratic, fraction base
            uint32_t ssrc = 0u;                                   // ... first window address
            unsigned spar = 0u;                                   // ... window parity
            bool chain_next = false;
            const bool interior = I3B_EDGE_SPLIT && kr >= ks_max && kr + SUB <= ke_min;
            if (chain) {
                steady = true;
                sA0 = cA0, sA1 = cA1, sA2 = cA2, sfb = cfb, ssrc = csrc, spar = cpar;
            } else {
                R = make_run_poly(S, js, Gsamp);
                // (32 taps: the rolled steady loop gains nothing over the per-pulse path, measured)
                if constexpr (!NU && I3B_STEADY && K < I3B_STEADY_MAX_TAPS) {
                    if (interior) {
                        const f32x2 Gr2 = bcast2(Gr);
                        // coordinate (minus floor(base) + 1/2) at the first and the last pulse of the run
                        const f32x2 XE = bcast2((float) (SUB - 1));
                        const f32x2 ange = fma2(fma2(fma2(R.A3, XE, R.A2), XE, R.A1), XE, R.A0);
                        const f32x2 g0 = fma2(R.A0, Gr2, R.f0m), ge = fma2(ange, Gr2, R.f0m);
                        const f32x2 mm = add2(g0, bcast2(MAGIC32));
                        const f32x2 tt = add2(mm, bcast2(-MAGIC32)); // integer part at the first pulse
                        const f32x2 fa = sub2(g0, tt), fe = sub2(ge, tt);
                        float m0, m1, fa0, fa1, fe0, fe1;
                        unpack2(mm, m0, m1);
                        unpack2(fa, fa0, fa1);
                        unpack2(fe, fe0, fe1);
                        const unsigned jj0 = (unsigned) (iw0 + __float_as_int(m0));
                        const unsigned jj1 = (unsigned) (iw1 + __float_as_int(m1));
                        const float worst0 = fmaxf(fabsf(fa0), fabsf(fa1));
                        const float worst = fmaxf(worst0, fmaxf(fabsf(fe0), fabsf(fe1)));
                        // steady: same integer part over the whole run (both pixels), adjacent windows
                        // inside the staged rows
                        const bool placed = jj1 == jj0 + 1u && jj0 < jmax;
                        steady = worst <= S.flim && placed;
                        if (steady) {
                            ssrc = la + ((jj0 >> 1) << 4);
                            spar = jj0 & 1u;
                            sfb = sub2(R.f0m, tt);
                            sA0 = R.A0, sA1 = R.A1, sA2 = R.A2;
                            int span = SUB - 1; // the quadratic replaces the cubic over this many pulse steps
                            if (I3B_PAIR_RUNS && TK == 2 * SUB && sub == 0 && kr + TK <= ke_min) {
                                // Can the NEXT run ride along?  Same test over all 16 pulses of the
                                // tile (the curvature allowance grows with the span squared): if it
                                // holds, the second run needs no polynomial and no proof of its own.
                                const f32x2 XT = bcast2((float) (TK - 1));
                                const f32x2 angt = fma2(fma2(fma2(R.A3, XT, R.A2), XT, R.A1), XT, R.A0);
                                const f32x2 ft = sub2(fma2(angt, Gr2, R.f0m), tt);
                                float ft0, ft1;
                                unpack2(ft, ft0, ft1);
                                constexpr float kHalf = 0.5f - 1e-5f;
                                constexpr float kGrow = (float) ((TK - 1) * (TK - 1)) / (float) ((SUB - 1) * (SUB - 1));
                                const float flim_t = fmaf(kGrow, S.flim - kHalf, kHalf);
                                if (fmaxf(worst0, fmaxf(fabsf(ft0), fabsf(ft1))) <= flim_t) {
                                    chain_next = true;
                                    span = TK - 1;
                                }
                            }
                            if (I3B_QUAD_RUN) {
                                // quadratic through the cubic at x = 0, span / 2, span
                                const float n = (float) span;
                                sA1 = fma2(R.A3, bcast2(-0.5f * n * n), sA1);
                                sA2 = fma2(R.A3, bcast2(1.5f * n), sA2);
                            }
                        }
                    }
                }
            }
            if (steady) {
                if (spar)
                    subtile_steady<K, D, Coef, 1, SUB>(S, sA0, sA1, sA2, R.A3, sfb, Gr, ssrc, row_bytes, P.zero, poly_addr);
                else
                    subtile_steady<K, D, Coef, 0, SUB>(S, sA0, sA1, sA2, R.A3, sfb, Gr, ssrc, row_bytes, P.zero, poly_addr);
                if (chain_next) {
                    // re-centre the quadratic on the next run's first pulse; window, parity and
                    // fraction base carry over
                    const f32x2 X = bcast2((float) SUB);
                    cA0 = fma2(fma2(sA2, X, sA1), X, sA0);
                    cA1 = fma2(sA2, bcast2(2.0f * SUB), sA1);
                    cA2 = sA2;
                    cfb = sfb;
                    csrc = ssrc + (uint32_t) SUB * row_bytes;
                    cpar = spar;
                }
                chain = chain_next;
            } else if (interior) {
                chain = false;
                {
                    if constexpr (I3B_NONSTEADY_CALL && !NU && I3B_STEADY && K < I3B_STEADY_MAX_TAPS) {
                        // a few percent of the runs: through the out-of-line per-pulse body (every
                        // pulse inside both apertures), which keeps ~900 instructions out of the
                        // address range the steady runs loop over
                        PairState T = S;
                        jjmax = tile_body_edge<K, D, Coef>(T, R, js, jjmax, la, row_bytes, wlo, jmax, Gr, 0u, 0u, P.zero,
                                                           poly_addr, xi_run, 0x7fffffffu, 0x7fffffffu);
                        S.accp[0] = T.accp[0]; S.accp[1] = T.accp[1];
                        S.accq[0] = T.accq[0]; S.accq[1] = T.accq[1];
                    } else {
                        tile_body<K, D, Coef, false, SUB>(S, R, js, jjmax, la, row_bytes, wlo, jmax, Gr, 0u, 0u, P.zero, poly_addr,
                                                          xi_run);
                    }
                }
            } else {
                chain = false;
                // aperture of each pixel within this launch (re-read: edge runs are a few per
                // pixel, their bounds do not deserve registers in the interior loops) and
                // k - kstart for the first pulse of the run
                unsigned krel[PX], kspan[PX];
#pragma unroll
                for (int p = 0; p < PX; ++p) {
                    const PixelRec* q = pix + gidx[p];
                    const int ks = max(q->kstart, P.k_begin), ke_p = min(q->kstop, P.k_end);
                    krel[p] = (unsigned) (kr - ks);
                    kspan[p] = (unsigned) max(ke_p - ks, 0);
                }
                const unsigned krel0 = krel[0], krel1 = krel[1];
                // (through a copy: the out-of-line call wants its argument in memory, and the
                // interior paths should not find their pair state there)
                PairState T = S;
                jjmax = tile_body_edge<K, D, Coef>(T, R, js, jjmax, la, row_bytes, wlo, jmax, Gr, krel0, krel1, P.zero, poly_addr,
                                                   xi_run, kspan[0], kspan[1]);
                S.accp[0] = T.accp[0]; S.accp[1] = T.accp[1];
                S.accq[0] = T.accq[0]; S.accq[1] = T.accq[1];
            }
        }
        };
#if I3B_NU_SPLIT
        if (P.xi) run_tile(NonUniformTag<1>{});
        else run_tile(NonUniformTag<0>{});
#else
        run_tile(NonUniformTag<-1>{});
#endif

        // pulse tile done: fold FP32 partials into FP64, release the stage.
        // sum s*e^{j phi} = (P.x - Q.y) + j (P.y + Q.x)
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            float px_, py_, qx_, qy_;
            unpack2(S.accp[p], px_, py_);
            unpack2(S.accq[p], qx_, qy_);
            accd[2 * p] += (double) (px_ - qy_);
            accd[2 * p + 1] += (double) (py_ + qx_);
            S.accp[p] = S.accq[p] = 0ull;
        }
        __syncwarp();
        if (I3B_LAST_ARRIVER) {
            // the last warp to leave tile n refills its stage with tile n + NSTAGE: every other
            // warp has released the stage already, nobody waits
            int last = 0;
            if (lane == 0) {
                mbar_arrive(&hdr->empty[s]);
                last = atomicAdd(&hdr->done[s], 1) == NTHREADS / 32 - 1;
                if (last) hdr->done[s] = 0;
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last && n + NSTAGE < ntiles) produce(n + NSTAGE);
        } else {
            if (lane == 0) mbar_arrive(&hdr->empty[s]);
        }
    }
#pragma unroll
    for (int p = 0; p < PX; ++p) {
        const int jj = line0 + lrow, ii = col0 + lcol + p;
        if (jj < P.out_lines && ii < P.out_width) acc[gidx[p]] = make_double2(accd[2 * p], accd[2 * p + 1]);
    }
    if (jjmax > jmax) status->window_overflow = 1;
}

// ---- host side: polynomial fit of the tap weights ---------------------------------------

struct FitResult {
    bool ok;
    int K, D;
    double max_err;
    int pair_deg[MAX_TAPS / 2 + 1]; // degree of each tap pair's polynomial (<= D)
    float even[MAX_TAPS / 2 + 1][MAX_COEF / 2];
    float odd[MAX_TAPS / 2 + 1][MAX_COEF / 2];
    TapPoly rows[MAX_TAPS / 2 + 1];
};

// least squares fit of y(g), g in [-1,1], by Chebyshev-node sampling + normal equations in a
// Chebyshev basis (well conditioned), converted to monomials in f = g/2.
static void fit_tap(const DevKernel& k, double cm, int D, double* mono /*[D+1] in f*/)
{
    const int N = 64;
    std::vector<double> g(N), y(N);
    for (int j = 0; j < N; ++j) {
        g[j] = std::cos(M_PI * (j + 0.5) / N);
        y[j] = (double) kernel_eval(k, cm - 0.5 * g[j]);
    }
    // discrete Chebyshev transform (nodes are Chebyshev zeros -> orthogonal)
    std::vector<double> a(D + 1);
    for (int i = 0; i <= D; ++i) {
        double sum = 0;
        for (int j = 0; j < N; ++j) sum += y[j] * std::cos(i * M_PI * (j + 0.5) / N);
        a[i] = (i == 0 ? 1.0 : 2.0) * sum / N;
    }
    // Chebyshev -> monomial in g
    std::vector<std::vector<double>> T(D + 1, std::vector<double>(D + 1, 0.0));
    T[0][0] = 1;
    if (D >= 1) T[1][1] = 1;
    for (int i = 2; i <= D; ++i)
        for (int p = 0; p <= i; ++p)
            T[i][p] = (p > 0 ? 2 * T[i - 1][p - 1] : 0.0) - T[i - 2][p];
    for (int p = 0; p <= D; ++p) {
        double c = 0;
        for (int i = p; i <= D; ++i) c += a[i] * T[i][p];
        mono[p] = c * std::pow(2.0, p); // g = 2 f
    }
}

// One tap pair (taps m and K-1-m, mirror images of each other) at degree D: symmetrised
// monomial coefficients as floats and the residual of the float Horner evaluation against the
// caller's kernel.
static double fit_pair(const DevKernel& k, int K, int m, int D, float* even, float* odd)
{
    const double cm = m - 0.5 * (K - 1);
    double mono[MAX_COEF + 1] = {};
    fit_tap(k, cm, D, mono);
    // taps m and K-1-m are mirror images: symmetrise (w_m(f) = E + f O, w_{K-1-m} = E - f O)
    if (2 * m != K - 1) {
        double mono2[MAX_COEF + 1] = {};
        fit_tap(k, -cm, D, mono2);
        for (int p = 0; p <= D; ++p) {
            const double mir = (p & 1) ? -mono2[p] : mono2[p];
            mono[p] = 0.5 * (mono[p] + mir);
        }
    } else {
        for (int p = 1; p <= D; p += 2) mono[p] = 0.0;
    }
    for (int i = 0; i < MAX_COEF / 2; ++i) even[i] = odd[i] = 0.f;
    for (int p = 0; p <= D; ++p) {
        if (p & 1) odd[p / 2] = (float) mono[p];
        else even[p / 2] = (float) mono[p];
    }
    const int NE = D / 2 + 1, NO = (D + 1) / 2;
    double worst = 0;
    for (int j = 0; j <= 400; ++j) {
        const float f = (float) (-0.5 + j / 400.0);
        const float h = f * f;
        float e = even[NE - 1];
        for (int i = NE - 2; i >= 0; --i) e = fmaf(e, h, even[i]);
        float o = odd[NO - 1];
        for (int i = NO - 2; i >= 0; --i) o = fmaf(o, h, odd[i]);
        const double wp = fmaf(f, o, e), wm = fmaf(-f, o, e);
        worst = std::max(worst, std::fabs(wp - (double) kernel_eval(k, cm - (double) f)));
        if (2 * m != K - 1)
            worst = std::max(worst, std::fabs(wm - (double) kernel_eval(k, -cm - (double) f)));
    }
    return worst;
}

// Outer tap pairs carry small weights and are fitted by low degrees to the same absolute
// accuracy; every pair gets the LOWEST degree whose residual is within PAIR_TOL, or within 1.5x
// of what the maximum degree achieves for it (pairs limited by the table's own interpolation
// error gain nothing from more terms).
constexpr double PAIR_TOL = 8e-6;
constexpr int MIN_DEGREE = 2;

static FitResult fit_kernel(const DevKernel& k, double tol)
{
    FitResult R;
    std::memset(&R, 0, sizeof(R));
    R.K = k.taps;
    R.D = (k.taps & 1) ? 6 : 7;
    const int K = R.K, D = R.D;
    const int nh = (K + 1) / 2;
    double worst = 0;
    for (int m = 0; m < nh; ++m) {
        float ev[MAX_COEF / 2], od[MAX_COEF / 2];
        const double r_full = fit_pair(k, K, m, D, R.even[m], R.odd[m]);
        double r_used = r_full;
        int d_used = D;
        for (int d = MIN_DEGREE; d < D; ++d) {
            const double r = fit_pair(k, K, m, d, ev, od);
            if (r <= std::max(PAIR_TOL, 1.5 * r_full)) {
                std::memcpy(R.even[m], ev, sizeof ev);
                std::memcpy(R.odd[m], od, sizeof od);
                r_used = r;
                d_used = d;
                break;
            }
        }
        R.pair_deg[m] = d_used;
        worst = std::max(worst, r_used);
    }
    for (int m = 0; m < nh; ++m)
        for (int i = 0; i < MAX_COEF / 2; ++i) {
            R.rows[m].e[i] = R.even[m][i];
            R.rows[m].o[i] = R.odd[m][i];
        }
    R.max_err = worst;
    R.ok = worst <= tol;
    return R;
}

constexpr double FIT_TOL = 3e-5;

// The fit costs a fraction of a millisecond and a call issues several launches: remember the
// last one per thread, keyed by the kernel's descriptor and a hash of its table.
static const FitResult& cached_fit(const DevKernel& k)
{
    struct Key {
        int kind, n, taps;
        double halfwidth, bandwidth;
        unsigned long long hash;
        bool operator==(const Key& o) const
        {
            return kind == o.kind && n == o.n && taps == o.taps && halfwidth == o.halfwidth &&
                   bandwidth == o.bandwidth && hash == o.hash;
        }
    };
    thread_local Key last_key {-1, 0, 0, 0.0, 0.0, 0ull};
    thread_local FitResult last_fit;
    unsigned long long h = 1469598103934665603ull;
    if (k.data && k.n > 0) {
        const unsigned char* b = reinterpret_cast<const unsigned char*>(k.data);
        for (size_t i = 0; i < (size_t) k.n * sizeof(float); ++i) h = (h ^ b[i]) * 1099511628211ull;
    }
    const Key key {k.kind, k.n, k.taps, k.halfwidth, k.bandwidth, h};
    if (!(key == last_key)) {
        last_fit = fit_kernel(k, FIT_TOL);
        last_key = key;
    }
    return last_fit;
}

// instantiated tap counts: every width up to 13 (a thread keeps all K weights and the K+1
// sample window in registers), then 16 and 32 (chunked MAC)
static bool taps_supported(int K) { return (K >= 3 && K <= 13) || K == 16 || K == 32; }

// Index of the baked table (tap_poly_imm.h) equal to this fit, or -1.  "Equal" = every
// coefficient within 2e-7 absolute: the two fits then give weights that differ by less than
// FP32 rounding of the Horner evaluation itself, whichever machine produced the table.
static int match_imm_table(const FitResult& R)
{
    for (int v = 0; v < imm::kNumTables; ++v) {
        const imm::Desc& d = imm::kDesc[v];
        if (d.taps != R.K || d.degree != R.D) continue;
        bool same_degrees = true;
        for (int m = 0; m < (R.K + 1) / 2; ++m) same_degrees = same_degrees && d.pair_deg[m] == R.pair_deg[m];
        if (!same_degrees) continue;
        double worst = 0;
        for (int m = 0; m < (R.K + 1) / 2; ++m)
            for (int i = 0; i < MAX_COEF / 2; ++i) {
                worst = std::max(worst, (double) std::fabs(d.even[m * 4 + i] - R.even[m][i]));
                worst = std::max(worst, (double) std::fabs(d.odd[m * 4 + i] - R.odd[m][i]));
            }
        if (worst <= 2e-7) return v;
    }
    return -1;
}

// I3B_FAST_NO_IMM=1 (test / tuning knob, read per call): always use the general kernel
static bool imm_enabled()
{
    const char* e = std::getenv("I3B_FAST_NO_IMM");
    return !(e && std::atoi(e) != 0);
}

int fast_fit(const DevKernel& hk, I3B_TapPolyFit* fit, char* why, size_t why_len)
{
    std::memset(fit, 0, sizeof *fit);
    fit->imm_variant = -1;
    fit->taps = hk.taps;
    if (why_len) why[0] = 0;
    if (hk.taps < 1 || hk.taps > MAX_TAPS) {
        snprintf(why, why_len, "tap count %d outside [1, %d]", hk.taps, MAX_TAPS);
        return 0;
    }
    if (hk.kind == I3B_KERNEL_BARTLETT || hk.kind == I3B_KERNEL_LINEAR) {
        snprintf(why, why_len, "piecewise-linear kernel is not polynomial per tap");
        return 0;
    }
    const FitResult R = fit_kernel(hk, FIT_TOL);
    fit->degree = R.D;
    fit->max_err = R.max_err;
    for (int m = 0; m < (R.K + 1) / 2; ++m) {
        fit->pair_degree[m] = R.pair_deg[m];
        for (int i = 0; i < MAX_COEF / 2; ++i) {
            fit->even[m][i] = R.even[m][i];
            fit->odd[m][i] = R.odd[m][i];
        }
    }
    fit->imm_variant = imm_enabled() ? match_imm_table(R) : -1;
    if (!taps_supported(hk.taps)) {
        snprintf(why, why_len, "tap count %d has no fast instantiation (3..13, 16, 32)", hk.taps);
        return 0;
    }
    if (!R.ok) {
        snprintf(why, why_len, "per-tap polynomial fit residual %.2e > %.1e", R.max_err, FIT_TOL);
        return 0;
    }
    fit->supported = 1;
    return 1;
}

bool fast_supported(const DevKernel& hk, char* why, size_t why_len)
{
    I3B_TapPolyFit fit;
    return fast_fit(hk, &fit, why, why_len) != 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
                    cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn) p;
    }
    return fn;
}

struct LaunchArgs {
    const CUtensorMap* map;
    const PolyTable* poly;
    const FastParams* FP;
    const PixelRec* pix;
    const PulseRec* pulse;
    double2* acc;
    const TileInfo* tiles;
    DevStatus* status;
    size_t smem;
    cudaStream_t s;
};

template<int K, int D, class Coef>
static int launch_inst(const LaunchArgs& L)
{
    auto kern = accumulate_fast_kernel<K, D, Coef>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) L.smem);
    if (e != cudaSuccess) return (int) e;
    const unsigned grid = (unsigned) (L.FP->tiles_rg * L.FP->tiles_az);
    kern<<<grid, NTHREADS, L.smem, L.s>>>(*L.map, *L.poly, *L.FP, L.pix, L.pulse, L.acc, L.tiles, L.status);
    return (int) cudaGetLastError();
}

// build-time specialised instantiations, one per table of tap_poly_imm.h
template<int V>
static int launch_imm(int v, const LaunchArgs& L)
{
    if (v == V) return launch_inst<imm::Table<V>::taps, imm::Table<V>::degree, CoefImm<V>>(L);
    if constexpr (V + 1 < imm::kNumTables) return launch_imm<V + 1>(v, L);
    return -1;
}

// Returns 0 on success, >0 a cudaError_t, -1 if the configuration is unsupported
// (caller falls back to the generic kernel).  Tiles flagged bad in `tiles` (a failed pixel)
// are skipped; the caller runs the generic kernel on those tiles.
int launch_accumulate_fast(const AccumParams& P, const DevKernel& hk, const PixelRec* pix,
                           const PulseRec* pulse, const float2* rc, double2* acc,
                           const TileInfo* tiles, DevStatus* status, cudaStream_t s)
{
    const int rc_rows = P.rc_rows, n_pulses = P.n_pulses;
    const double out_in_spacing_ratio = P.spacing_ratio;
    // the magic-number splits need |fc*tau| and |u| below 2^28
    if (P.fc * (P.swst + (P.nr + 64) * P.dtau) > 2.6e8 || P.nr > (1 << 27)) return -1;
    if (!taps_supported(hk.taps)) return -1;
    const FitResult& R = cached_fit(hk);
    if (!R.ok) return -1;
    const int K = hk.taps;
    int W = (int) std::ceil(TILE_RG * std::fabs(out_in_spacing_ratio) * 1.002) + K + 16;
#ifdef I3B_EXTRA_W
    W += I3B_EXTRA_W; // experiment: stage more samples than needed (TMA sensitivity)
#endif
    W = (W + 1) & ~1;
    if (W > 256) return -1;
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return -1;
    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t) P.nr, (cuuint64_t) rc_rows};
    const cuuint64_t gstride[1] = {(cuuint64_t) P.rc_pitch * sizeof(float2)};
    const cuuint32_t box[2] = {(cuuint32_t) W, (cuuint32_t) TK};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*) rc, gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return -1;

    FastParams FP;
    FP.npix = P.npix;
    FP.out_lines = P.out_lines;
    FP.out_width = P.out_width;
    FP.nr = P.nr;
    FP.rc_k0 = P.rc_k0;
    FP.rc_rows = rc_rows;
    FP.k_begin = P.k_begin;
    FP.k_end = P.k_end;
    FP.n_pulses = n_pulses;
    FP.W = W;
    FP.tiles_rg = (P.out_width + TILE_RG - 1) / TILE_RG;
    {
        const int line_end = P.line_end > 0 ? std::min(P.line_end, P.out_lines) : P.out_lines;
        FP.tile_j0 = P.line_begin / TILE_AZ;
        FP.tiles_az = (line_end + TILE_AZ - 1) / TILE_AZ - FP.tile_j0;
        if (FP.tiles_az <= 0) return 0;
    }
    FP.k_landed = P.k_landed;
    FP.seg = P.seg == 128 ? 128 : 64;
    FP.G = 1.0 / (P.fc * P.dtau);
    FP.Gf = (float) FP.G;
    FP.Grf = (float) (FP.G / 6.283185307179586476925);
    FP.U0 = P.swst / P.dtau;
    FP.fc = P.fc;
    FP.zero = 0;
    FP.tn = P.tn;
    FP.xi = P.xi;
    const size_t smem = HEADER_BYTES + NSTAGE * stage_bytes(W) + (size_t) NTHREADS * 6 * PX * sizeof(double);
    PolyTable PT;
    std::memcpy(PT.rows, R.rows, sizeof PT.rows);
    const LaunchArgs L{&map, &PT, &FP, pix, pulse, acc, tiles, status, smem, s};

    const int v = imm_enabled() ? match_imm_table(R) : -1;
    if (v >= 0) {
        const int r = launch_imm<0>(v, L);
        if (r >= 0) return r;
    }
    switch (K) {
    case 3: return launch_inst<3, 6, CoefBank>(L);
    case 4: return launch_inst<4, 7, CoefBank>(L);
    case 5: return launch_inst<5, 6, CoefBank>(L);
    case 6: return launch_inst<6, 7, CoefBank>(L);
    case 7: return launch_inst<7, 6, CoefBank>(L);
    case 8: return launch_inst<8, 7, CoefBank>(L);
    case 9: return launch_inst<9, 6, CoefBank>(L);
    case 10: return launch_inst<10, 7, CoefBank>(L);
    case 11: return launch_inst<11, 6, CoefBank>(L);
    case 12: return launch_inst<12, 7, CoefBank>(L);
    case 13: return launch_inst<13, 6, CoefBank>(L);
    case 16: return launch_inst<16, 7, CoefBank>(L);
    case 32: return launch_inst<32, 7, CoefBank>(L);
    default: return -1;
    }
}

int fast_tiles(int out_lines, int out_width)
{
    return ((out_width + TILE_RG - 1) / TILE_RG) * ((out_lines + TILE_AZ - 1) / TILE_AZ);
}

int fast_pulse_tile() { return TK; }
// Segment length for a scene: the cubic through four exact phase values h = seg / prf apart
// is off by at most 0.0234 h^4 |d4 phase / dt4|; for a platform at speed v passing a target at
// range r that derivative is (4 pi / wavelength) 3 v^4 / r^3 at closest approach, where it is
// largest (x1.5 for margin).  128 pulses where that stays below 1.5e-6 rad (NISAR L-band:
// 1.0e-6), else 64 (the airborne configuration: 2e-5 rad at 128, 1.3e-6 at 64).
int fast_segment(double wavelength, double prf, double v_max, double r_min)
{
    if (I3B_SEG) return I3B_SEG;
    if (!(wavelength > 0) || !(prf > 0) || !(v_max > 0) || !(r_min > 0)) return 64;
    const double d4 = 1.5 * (4.0 * M_PI / wavelength) * 3.0 * std::pow(v_max, 4) / std::pow(r_min, 3);
    const double h = 128.0 / prf;
    return 0.0234 * std::pow(h, 4) * d4 <= 1.5e-6 ? 128 : 64;
}

void fast_tile_shape(int* tile_az, int* tile_rg)
{
    *tile_az = TILE_AZ;
    *tile_rg = TILE_RG;
}

} // namespace i3b
