// TEST INFRASTRUCTURE (oracle/_ref build only).  Stand-in for pyre's journal header (pyre
// 1.12.5 is not installed): channels that swallow what is streamed into them.  The reference
// sources compiled into oracle/_ref only use them on error paths (core/DateTime.cpp:12,
// core/Spline2dInterpolator.cpp:18-24).
#pragma once
#include <string>
#define __HERE__ __FILE__, __LINE__, __func__
namespace pyre { namespace journal {
struct locator_t {};
inline locator_t at(const char*, int, const char*) { return {}; }
struct manip_t {};
constexpr manip_t newline {}, endl {};
class channel_t {
public:
    explicit channel_t(const std::string&) {}
    template<typename T> channel_t& operator<<(const T&) { return *this; }
};
using error_t = channel_t;
using warning_t = channel_t;
using info_t = channel_t;
using debug_t = channel_t;
}}
