// TEST INFRASTRUCTURE -- builds into oracle/_ref/libtdbp_refcuda.so.  Not product code: only
// bench.py's reference-CUDA comparator leg and tests/ may load it.
//
// "Reference kernels, reduced harness" (SURVEY.md 8c): the UNMODIFIED reference CUDA
// backprojection, compiled where it lies under /root/reference/cxx:
//
//   isce3/cuda/focus/Backproject.cu   all 8 kernels + the host driver (:79-754)
//   isce3/cuda/core/Orbit.cu, OrbitView.icc, Kernels.icc, Interp1d.icc
//   isce3/cuda/container/RadarGeometry.icc, isce3/cuda/except/Error.cpp
//   + the host sources of oracle/Makefile's `ref` target
//
// against oracle/shim (vector algebra standing in for Eigen, host LUT2d / DEMInterpolator /
// Projections) and oracle/cuda_ref/shim: gpuLUT2d and gpuDEMInterpolator REDUCED to the
// constant (no-data) case, gpuGeometry.h reduced to the two bracket wrappers.  So it runs the
// flat-DEM / zero-Doppler configurations (C1, C2, C3, C5), not C4.
#include "../ref_capi.cpp" // argument unflattening shared with the CPU reference wrapper

#include <isce3/cuda/focus/Backproject.h>

extern "C" {

const char* tdbp_refcuda_last_error() { return g_err.c_str(); }

// isce3::cuda::focus::backproject, cxx/isce3/cuda/focus/Backproject.cu:698-754
int tdbp_refcuda_backproject(const I3B_BackprojectArgs* a)
{
    return guarded([&]() {
        const auto out_geom = make_geometry(a->out_geometry);
        const auto in_geom = make_geometry(a->in_geometry);
        const isce3::geometry::DEMInterpolator dem(a->dem);
        const auto kernel = make_kernel(a->kernel);
        const auto atm = a->dry_tropo_model == I3B_TROPO_TSX ? isce3::focus::DryTroposphereModel::TSX
                                                           : isce3::focus::DryTroposphereModel::NoDelay;
        isce3::geometry::detail::Rdr2GeoBracketParams r2g;
        r2g.tol_height = a->rdr2geo.tol_height;
        r2g.look_min = a->rdr2geo.look_min;
        r2g.look_max = a->rdr2geo.look_max;
        isce3::geometry::detail::Geo2RdrBracketParams g2r;
        g2r.tol_aztime = a->geo2rdr.tol_aztime;
        if (a->geo2rdr.has_time_start) g2r.time_start = a->geo2rdr.time_start;
        if (a->geo2rdr.has_time_end) g2r.time_end = a->geo2rdr.time_end;
        const auto ec = isce3::cuda::focus::backproject(
                reinterpret_cast<std::complex<float>*>(a->out), out_geom,
                reinterpret_cast<const std::complex<float>*>(a->in), in_geom, dem, a->fc, a->ds,
                *kernel, atm, r2g, g2r, a->batch, a->height);
        return int(ec);
    });
}

} // extern "C"
