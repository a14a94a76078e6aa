"""``isce3.cuda.focus`` mirror (python/packages/isce3/cuda/focus/__init__.py:1)."""
from ..focus import backproject  # noqa: F401
