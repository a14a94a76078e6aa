"""``isce3.cuda`` namespace mirror (python/packages/isce3/cuda/__init__.py)."""
from . import focus  # noqa: F401
