// Kernel table of the reference's 2-D sinc interpolator, as its constructor builds it
// (cxx/isce3/core/Sinc2dInterpolator.cpp:13-40 with _sinc_coef :116-133, called with
// beta = 1, pedestal = 0, weight = 1; SINC_LEN = 8 taps, SINC_SUB = 8192 sub-sample
// positions, core/Constants.h:32-35; rows normalised to unit sum).  Host only.
#pragma once
#include <cmath>
#include <vector>

namespace i3b {

constexpr int kSincLen = 8;    // SINC_LEN
constexpr int kSincSub = 8192; // SINC_SUB
constexpr int kSincHalf = kSincLen / 2;

// table[i * kSincLen + j]: weight of tap j at sub-sample position i
inline std::vector<double> make_sinc_table()
{
    const int n = kSincSub * kSincLen;
    std::vector<double> filter(n), table(n);
    const double pedestal = 0.0, beta = 1.0;
    const double wgthgt = (1.0 - pedestal) / 2.0;
    const double soff = (n - 1.) / 2.;
    for (int i = 0; i < n; ++i) {
        const double wgt = (1. - wgthgt) + (wgthgt * std::cos((M_PI * (i - soff)) / soff));
        const double s = (std::floor(i - soff) * beta) / (1. * kSincSub);
        const double fct = (s != 0.) ? (std::sin(M_PI * s) / (M_PI * s)) : 1.;
        filter[i] = fct * wgt;
    }
    for (int i = 0; i < kSincSub; ++i) {
        double ssum = 0.0;
        for (int j = 0; j < kSincLen; ++j) ssum += filter[i + kSincSub * j];
        for (int j = 0; j < kSincLen; ++j) table[(size_t) i * kSincLen + j] = filter[i + kSincSub * j] / ssum;
    }
    return table;
}

} // namespace i3b
