// TEST INFRASTRUCTURE -- builds into oracle/_ref/libtdbp_adapter.so.
//
// Exercises the isce3-side adapter (integration/isce3/cuda/focus/BackprojectB200.cpp, the file
// a maintainer adds to isce3) for real: this wrapper turns a flat argument block into the
// reference's own objects (RadarGeometry, Orbit, Kernel<float>, ... exactly as
// oracle/ref_capi.cpp does for the CPU reference), calls isce3::cuda::focus::backproject --
// which is the adapter -- and the adapter flattens them again and calls the product's
// i3b_backproject.  Links against isce3_b200/libisce3_b200_backproject.so.
#include "ref_capi.cpp"

#include <isce3/cuda/focus/Backproject.h>

extern "C" {

const char* tdbp_adapter_last_error() { return g_err.c_str(); }

int tdbp_adapter_backproject(const I3B_BackprojectArgs* a)
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
