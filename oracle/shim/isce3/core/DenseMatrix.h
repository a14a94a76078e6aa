// TEST INFRASTRUCTURE (oracle/_ref build only).  Shadows
// cxx/isce3/core/DenseMatrix.h: Mat3 is only named, never used, on the TDBP path.
#pragma once
#include <isce3/core/forward.h>
namespace isce3 { namespace core {
template<int N, typename T> class DenseMatrix {};
}}
