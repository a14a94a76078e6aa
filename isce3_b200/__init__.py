"""isce3_b200 -- B200-native time-domain backprojection behind the
``isce3.focus.backproject`` / ``isce3.cuda.focus.backproject`` API.

Only the TDBP hot path of isce-framework/isce3 lives here:
CUDA kernels + C-ABI in ``csrc/``, the host-side mirror of the reference's
operator interface in ``focus.py`` and the value types it takes in ``core.py``,
``product.py``, ``container.py``, ``geometry.py``.
"""
from . import container, core, cuda, focus, geometry, product  # noqa: F401

__version__ = "0.1.0"
