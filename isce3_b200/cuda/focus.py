"""``isce3.cuda.focus`` mirror (python/packages/isce3/cuda/focus/__init__.py:1).

``backproject`` here is the COMPILED binding (``isce3_b200/ext/_backproject``: pybind11 over
the C-ABI, built from csrc/pybind_module.cpp by ``__graft_entry__.build()``), with the call
shape of ``isce3.ext.isce3.cuda.focus.backproject``
(python/extensions/pybind_isce3/cuda/focus/Backproject.cpp:25-117).  ``isce3_b200.focus``
holds the ctypes route to the same entry point plus the extensions (resident plans, blocks,
pulse times)."""
try:
    from ..ext._backproject import backproject  # noqa: F401
except ImportError as exc:  # pragma: no cover - build() makes it
    _why = exc

    def backproject(*args, **kwargs):
        raise ImportError("isce3_b200.ext._backproject is not built (run __graft_entry__.build()): "
                          f"{_why}")
