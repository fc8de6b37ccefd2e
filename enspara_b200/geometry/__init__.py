"""Distance kernels with the reference's module path (enspara.geometry.libdist)."""
from . import libdist  # noqa: F401
