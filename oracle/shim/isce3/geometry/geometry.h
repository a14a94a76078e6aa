// TEST INFRASTRUCTURE (oracle/_ref build only).  Shadows
// cxx/isce3/geometry/geometry.h (pulls in GDAL-backed types); Backproject.cpp:12
// includes it but uses nothing declared there.
#pragma once
