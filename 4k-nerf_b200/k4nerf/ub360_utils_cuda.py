"""Drop-in for the reference's pybind module ``ub360_utils_cuda`` (lib/cuda/ub360_utils.cpp:20-26)."""
from .render_utils_cuda import cumdist_thres  # noqa: F401

__all__ = ['cumdist_thres']
