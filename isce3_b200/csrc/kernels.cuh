// Interpolation-kernel evaluation shared by host (polynomial fitting for the fast
// path) and device (generic accumulation kernel).  Follows Kernel<float>::operator()
// of cxx/isce3/core/Kernels.icc: Bartlett/Linear :15-23, Knab :29-52 (evaluated in
// float, complex sqrt so it is defined slightly outside the support), Tabulated
// :139-154 (double index math, float table, _1_dx stored as float), Cheby :191-211
// (Clenshaw in float); sinc from cxx/isce3/math/Sinc.icc:69-91.
#pragma once
#include <cmath>

#include "common.cuh"

namespace i3b {

__host__ __device__ inline float sincf_ref(float t)
{
    const float eps1 = 3.4526698e-4f; // sqrt(FLT_EPSILON)
    const float eps2 = 1.8581361e-2f; // sqrt(sqrt(FLT_EPSILON))
    const float x = 3.14159265358979323846f * fabsf(t);
    if (x < eps2) {
        float out = 1.f;
        if (x > eps1) out -= x * x / 6.f;
        return out;
    }
    return sinf(x) / x;
}

__host__ __device__ inline float knabf_ref(double t, double halfwidth, double bandwidth)
{
    const float st = sincf_ref((float) t);
    const float hw = (float) halfwidth, bw = (float) bandwidth;
    const float c = (float) (M_PI * hw * (1.0 - bw));
    const float tf = (float) t / hw;
    const float a = (float) (1.0 - tf * tf);
    // real(cosh(c*sqrt(a)))/cosh(c): cosh for a >= 0, cos of the imaginary root otherwise
    const float num = (a >= 0.f) ? coshf(c * sqrtf(a)) : cosf(c * sqrtf(-a));
    return (num / coshf(c)) * st;
}

__host__ __device__ inline float kernel_eval(const DevKernel& k, double t)
{
    switch (k.kind) {
    case I3B_KERNEL_BARTLETT:
    case I3B_KERNEL_LINEAR: {
        const double t2 = fabs(t / k.halfwidth);
        if (t2 > 1.0) return 0.f;
        return (float) (1.0 - t2);
    }
    case I3B_KERNEL_KNAB: return knabf_ref(t, k.halfwidth, k.bandwidth);
    case I3B_KERNEL_TABULATED: {
        const double ax = fabs(t);
        if (ax > k.halfwidth) return 0.f;
        const double axn = ax * k.one_dx;
        int i = (int) floor(axn);
        i = i < k.imax ? i : k.imax;
        const float a = k.data[i], b = k.data[i + 1];
        return (float) (a + (axn - i) * (b - a));
    }
    case I3B_KERNEL_CHEBY: {
        const double ax = fabs(t);
        if (ax > k.halfwidth) return 0.f;
        const float q = (float) ((ax * k.cheb_scale) - 1.f);
        const float twoq = 2.f * q;
        float bk = 0.f, bk1 = 0.f, bk2 = 0.f;
        for (int i = k.n - 1; i > 0; --i) {
            bk = k.data[i] + twoq * bk1 - bk2;
            bk2 = bk1;
            bk1 = bk;
        }
        return k.data[0] + q * bk1 - bk2;
    }
    default: return nanf("");
    }
}

} // namespace i3b
