// Standalone pybind11 module ``isce3_b200.ext._backproject``: the B200 backprojection behind a
// compiled extension with the call shape of the reference binding
//   python/extensions/pybind_isce3/cuda/focus/Backproject.cpp:25-117
// (same positional arguments, defaults, argument checks, GIL released around the C++ call, bool
// return) -- host C++ over the C-ABI of include/isce3_b200_backproject.h, no ctypes, no torch.
//
// isce3 itself is not buildable in this image (Eigen/GDAL/HDF5/pyre absent), so the value types
// (RadarGeometry, Orbit, LUT2d, DEMInterpolator, Kernel<float>) cannot be the reference's C++
// classes here: the module reads the same information through the attribute names the
// reference registers on those types (the stand-ins of isce3_b200/{core,
// product,container,geometry}.py and real isce3 objects both carry them) and flattens it into
// the C descriptors.  In an isce3 build the adapter of integration/isce3/cuda/focus/
// BackprojectB200.cpp does the same flattening from the C++ objects and this file shrinks to
// the reference binding verbatim.
#include <pybind11/complex.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cmath>
#include <complex>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/isce3_b200_backproject.h"

namespace py = pybind11;

namespace {

// exception classes with the reference's names; pybind11 translates the std bases the same
// way it translates isce3::except::{InvalidArgument, DomainError, OutOfRange, ...}
struct InvalidArgument : std::invalid_argument { using std::invalid_argument::invalid_argument; };
struct DomainError : std::domain_error { using std::domain_error::domain_error; };
struct OutOfRange : std::out_of_range { using std::out_of_range::out_of_range; };
struct OverflowError : std::overflow_error { using std::overflow_error::overflow_error; };
struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };

[[noreturn]] void raise_status(int status)
{
    const std::string msg = i3b_last_error() ? i3b_last_error() : "";
    switch (status) {
    case I3B_EXC_INVALID_ARGUMENT: throw InvalidArgument(msg);
    case I3B_EXC_DOMAIN_ERROR: throw DomainError(msg);
    case I3B_EXC_OVERFLOW_ERROR: throw OverflowError(msg);
    case I3B_EXC_OUT_OF_RANGE: throw OutOfRange(msg);
    case I3B_EXC_CUDA_ERROR:
    case I3B_EXC_NO_DEVICE: throw CudaError(msg);
    default: throw std::runtime_error(msg.empty() ? "isce3_b200 error " + std::to_string(status) : msg);
    }
}

// Keeps the numpy arrays behind the flattened pointers alive for the duration of the call.
struct Keep {
    std::vector<py::object> objs;
    template<typename T>
    const T* hold(const py::handle& h)
    {
        auto a = py::array_t<T, py::array::c_style | py::array::forcecast>::ensure(h);
        if (!a) throw InvalidArgument("expected a numeric array");
        objs.push_back(a);
        return a.data();
    }
};

I3B_Orbit flatten_orbit(const py::handle& o, Keep& keep)
{
    I3B_Orbit d {};
    const py::object time = o.attr("time"); // Linspace: first, spacing
    d.t0 = time.attr("first").cast<double>();
    d.dt = time.attr("spacing").cast<double>();
    d.n = o.attr("size").cast<int>();
    d.method = py::int_(o.attr("interp_method")).cast<int>();
    d.pos = keep.hold<double>(o.attr("position"));
    d.vel = keep.hold<double>(o.attr("velocity"));
    return d;
}

I3B_LUT2d flatten_lut2d(const py::handle& l, Keep& keep)
{
    I3B_LUT2d d {};
    d.have_data = l.attr("have_data").cast<bool>() ? 1 : 0;
    d.bounds_error = l.attr("bounds_error").cast<bool>() ? 1 : 0;
    d.method = py::int_(l.attr("interp_method")).cast<int>();
    d.ref_value = l.attr("ref_value").cast<double>();
    if (d.have_data) {
        auto a = py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(l.attr("data"));
        if (!a || a.ndim() != 2) throw InvalidArgument("LUT2d data must be 2-D");
        keep.objs.push_back(a);
        d.length = a.shape(0);
        d.width = a.shape(1);
        d.xstart = l.attr("x_start").cast<double>();
        d.ystart = l.attr("y_start").cast<double>();
        d.dx = l.attr("x_spacing").cast<double>();
        d.dy = l.attr("y_spacing").cast<double>();
        d.data = a.data();
    }
    return d;
}

I3B_RadarGrid flatten_grid(const py::handle& g)
{
    I3B_RadarGrid d {};
    d.sensing_start = g.attr("sensing_start").cast<double>();
    d.prf = g.attr("prf").cast<double>();
    d.starting_range = g.attr("starting_range").cast<double>();
    d.range_pixel_spacing = g.attr("range_pixel_spacing").cast<double>();
    d.wavelength = g.attr("wavelength").cast<double>();
    d.length = g.attr("length").cast<int64_t>();
    d.width = g.attr("width").cast<int64_t>();
    d.look_side = py::int_(g.attr("lookside")).cast<int>();
    return d;
}

I3B_RadarGeometry flatten_geometry(const py::handle& geom, Keep& keep)
{
    I3B_RadarGeometry d {};
    d.grid = flatten_grid(geom.attr("radar_grid"));
    d.orbit = flatten_orbit(geom.attr("orbit"), keep);
    d.doppler = flatten_lut2d(geom.attr("doppler"), keep);
    // (seconds since 1970, fractional second): only compared for equality in / out
    const py::tuple ep = geom.attr("reference_epoch").attr("epoch_pair")();
    d.ref_epoch_sec = ep[0].cast<int64_t>();
    d.ref_epoch_frac = ep[1].cast<double>();
    return d;
}

I3B_DEM flatten_dem(const py::handle& dem, Keep& keep)
{
    I3B_DEM d {};
    d.have_raster = dem.attr("have_raster").cast<bool>() ? 1 : 0;
    d.epsg = dem.attr("epsg_code").cast<int>();
    d.method = py::int_(dem.attr("interp_method")).cast<int>();
    d.ref_height = dem.attr("ref_height").cast<double>();
    if (d.have_raster) {
        auto a = py::array_t<float, py::array::c_style | py::array::forcecast>::ensure(dem.attr("data"));
        if (!a || a.ndim() != 2) throw InvalidArgument("DEM raster must be 2-D");
        keep.objs.push_back(a);
        d.length = a.shape(0);
        d.width = a.shape(1);
        d.xstart = dem.attr("x_start").cast<double>();
        d.ystart = dem.attr("y_start").cast<double>();
        d.dx = dem.attr("delta_x").cast<double>();
        d.dy = dem.attr("delta_y").cast<double>();
        d.data = a.data();
    }
    return d;
}

// Kernel<float>: the dynamic type decides the descriptor (cuda/focus/Backproject.cu:715-752
// dispatches on the same five types and throws "not implemented" for anything else)
I3B_Kernel flatten_kernel(const py::handle& k, Keep& keep)
{
    I3B_Kernel d {};
    const std::string name = py::str(py::type::of(k).attr("__name__"));
    d.width = k.attr("width").cast<double>();
    if (name == "BartlettKernelF32") {
        d.kind = I3B_KERNEL_BARTLETT;
    } else if (name == "LinearKernelF32") {
        d.kind = I3B_KERNEL_LINEAR;
    } else if (name == "KnabKernelF32") {
        d.kind = I3B_KERNEL_KNAB;
        d.bandwidth = k.attr("bandwidth").cast<double>();
    } else if (name == "TabulatedKernelF32") {
        d.kind = I3B_KERNEL_TABULATED;
        auto a = py::array_t<float, py::array::c_style | py::array::forcecast>::ensure(k.attr("table"));
        if (!a) throw InvalidArgument("TabulatedKernelF32 without a table");
        keep.objs.push_back(a);
        d.n = (int32_t) a.size();
        d.data = a.data();
    } else if (name == "ChebyKernelF32") {
        d.kind = I3B_KERNEL_CHEBY;
        auto a = py::array_t<float, py::array::c_style | py::array::forcecast>::ensure(k.attr("coeffs"));
        if (!a) throw InvalidArgument("ChebyKernelF32 without coefficients");
        keep.objs.push_back(a);
        d.n = (int32_t) a.size();
        d.data = a.data();
    } else {
        throw std::runtime_error("not implemented"); // Backproject.cu:750-752
    }
    return d;
}

// pybind_isce3/focus/Backproject.cpp:29-78
I3B_Rdr2GeoBracketParams parse_rdr2geo_params(const py::dict& params)
{
    I3B_Rdr2GeoBracketParams out {1e-5, 0.0, M_PI / 2}; // geometry/detail/Rdr2Geo.h:85-97
    for (auto item : params) {
        const std::string key = py::str(item.first);
        if (key == "tol_height") out.tol_height = item.second.cast<double>();
        else if (key == "look_min") out.look_min = item.second.cast<double>();
        else if (key == "look_max") out.look_max = item.second.cast<double>();
        else throw InvalidArgument("unexpected rdr2geo_bracket keyword: " + key);
    }
    return out;
}

I3B_Geo2RdrBracketParams parse_geo2rdr_params(const py::dict& params)
{
    I3B_Geo2RdrBracketParams out {1e-7, 0, 0, 0.0, 0.0}; // geometry/detail/Geo2Rdr.h:54-68
    for (auto item : params) {
        const std::string key = py::str(item.first);
        if (key == "tol_aztime") {
            out.tol_aztime = item.second.cast<double>();
        } else if (key == "time_start") {
            if (!item.second.is_none()) {
                out.has_time_start = 1;
                out.time_start = item.second.cast<double>();
            }
        } else if (key == "time_end") {
            if (!item.second.is_none()) {
                out.has_time_end = 1;
                out.time_end = item.second.cast<double>();
            }
        } else {
            throw InvalidArgument("unexpected geo2rdr_bracket keyword: " + key);
        }
    }
    return out;
}

bool backproject(py::array_t<std::complex<float>, py::array::c_style> out, const py::object& out_geometry,
                 py::array_t<std::complex<float>, py::array::c_style> in, const py::object& in_geometry,
                 const py::object& dem, double fc, double ds, const py::object& kernel,
                 const std::string& dry_tropo_model, py::dict rdr2geo_params, py::dict geo2rdr_params,
                 int batch, std::optional<py::array_t<float, py::array::c_style>> height,
                 std::optional<std::vector<int>> devices)
{
    Keep keep;
    I3B_BackprojectArgs a {};
    a.abi_version = I3B_ABI_VERSION;
    a.out_geometry = flatten_geometry(out_geometry, keep);
    a.in_geometry = flatten_geometry(in_geometry, keep);
    // argument checks of the reference binding (:42-89), same messages
    if (out.ndim() != 2) throw InvalidArgument("output array must be 2-D");
    if (out.shape(0) != a.out_geometry.grid.length || out.shape(1) != a.out_geometry.grid.width)
        throw InvalidArgument("output array shape must match output radar grid shape");
    if (in.ndim() != 2) throw InvalidArgument("input signal data must be 2-D");
    if (in.shape(0) != a.in_geometry.grid.length || in.shape(1) != a.in_geometry.grid.width)
        throw InvalidArgument("input signal data shape must match input radar grid shape");
    a.out = reinterpret_cast<float*>(out.mutable_data());
    a.in = reinterpret_cast<const float*>(in.data());
    if (height.has_value()) {
        auto& h = height.value();
        if (h.ndim() != 2 || h.shape(0) != a.out_geometry.grid.length || h.shape(1) != a.out_geometry.grid.width)
            throw InvalidArgument("height array shape must match output radar grid shape");
        a.height = h.mutable_data();
    }
    if (dry_tropo_model == "nodelay") a.dry_tropo_model = I3B_TROPO_NODELAY;
    else if (dry_tropo_model == "tsx") a.dry_tropo_model = I3B_TROPO_TSX;
    else throw InvalidArgument("unexpected dry troposphere model '" + dry_tropo_model + "'");
    a.rdr2geo = parse_rdr2geo_params(rdr2geo_params);
    a.geo2rdr = parse_geo2rdr_params(geo2rdr_params);
    if (batch < 1) throw DomainError("batch size must be > 0");
    a.batch = batch;
    a.dem = flatten_dem(dem, keep);
    a.fc = fc;
    a.ds = ds;
    a.kernel = flatten_kernel(kernel, keep);
    std::vector<int32_t> devs;
    if (devices.has_value()) {
        devs.assign(devices->begin(), devices->end());
        a.n_devices = (int32_t) devs.size();
        a.devices = devs.data();
    }
    int status;
    {
        py::gil_scoped_release release;
        status = i3b_backproject(&a);
    }
    if (status < 0) raise_status(status);
    // like the reference (":98-99 TODO bind ErrorCode class"): nonzero on failure
    return status != I3B_SUCCESS;
}

py::dict last_stats()
{
    I3B_Stats s {};
    i3b_last_stats(&s);
    py::dict d;
    d["pixel_pulses"] = s.pixel_pulses;
    d["ms_total"] = s.ms_total;
    d["ms_h2d"] = s.ms_h2d;
    d["ms_target_solve"] = s.ms_target_solve;
    d["ms_accumulate"] = s.ms_accumulate;
    d["ms_d2h"] = s.ms_d2h;
    d["accumulate_launches"] = s.accumulate_launches;
    d["total_launches"] = s.total_launches;
    d["used_fast_kernel"] = s.used_fast_kernel;
    d["taps"] = s.taps;
    d["h2d_bytes"] = s.h2d_bytes;
    d["d2h_bytes"] = s.d2h_bytes;
    d["n_devices"] = s.n_devices;
    d["fast_variant"] = s.fast_variant;
    return d;
}

} // namespace

PYBIND11_MODULE(_backproject, m)
{
    m.doc() = "B200 time-domain backprojection (compiled binding over the C-ABI of isce3_b200_backproject.h)";
    py::register_exception<InvalidArgument>(m, "InvalidArgument", PyExc_ValueError);
    py::register_exception<DomainError>(m, "DomainError", PyExc_ValueError);
    py::register_exception<OutOfRange>(m, "OutOfRange", PyExc_IndexError);
    py::register_exception<OverflowError>(m, "OverflowError", PyExc_OverflowError);
    py::register_exception<CudaError>(m, "CudaError", PyExc_RuntimeError);
    m.def("backproject", &backproject,
          R"(
                Focus in azimuth via time-domain backprojection.
            )",
          py::arg("out"), py::arg("out_geometry"), py::arg("in"), py::arg("in_geometry"), py::arg("dem"),
          py::arg("fc"), py::arg("ds"), py::arg("kernel"), py::arg("dry_tropo_model") = "tsx",
          py::arg("rdr2geo_params") = py::dict(), py::arg("geo2rdr_params") = py::dict(),
          py::arg("batch") = 1024, py::arg("height") = py::none(), py::kw_only(),
          py::arg("devices") = py::none());
    m.def("last_stats", &last_stats);
    m.def("version", []() { return std::string(i3b_version()); });
    m.def("device_count", []() { return i3b_device_count(); });
}
