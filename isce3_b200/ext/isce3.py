"""Mirror of the single pybind11 module ``isce3.ext.isce3``
(python/extensions/pybind_isce3/isce3.cpp:25-49), restricted to the submodules the
backproject path and its tests touch."""
import types as _types

from .. import container as _container
from .. import core as _core
from .. import cuda as _cuda
from .. import focus as _focus
from .. import geometry as _geometry
from .. import product as _product

core = _core
product = _product
container = _container
geometry = _geometry
cuda = _cuda

# isce3.ext.isce3.focus.backproject is the CPU path in the reference (no `batch`
# argument, pybind_isce3/focus/Backproject.cpp:83-95); here both names reach the
# same CUDA backend.
focus = _types.SimpleNamespace(
    backproject=lambda out, out_geometry, in_, in_geometry, dem, fc, ds, kernel,
    dry_tropo_model="tsx", rdr2geo_params=None, geo2rdr_params=None, height=None:
    _focus.backproject(out, out_geometry, in_, in_geometry, dem, fc, ds, kernel,
                       dry_tropo_model, rdr2geo_params, geo2rdr_params, 1024, height),
    # pybind_isce3/focus/focus.cpp: RangeComp, form_linear_chirp
    RangeComp=_focus.RangeComp,
    form_linear_chirp=_focus.form_linear_chirp)
