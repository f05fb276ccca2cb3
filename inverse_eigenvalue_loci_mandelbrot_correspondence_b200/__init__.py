"""B200-native (sm_100a) hot path of aortizt/inverse-eigenvalue-loci-mandelbrot-correspondence.

Python host code over a C-ABI CUDA library (liblm_b200.so, include/lm_b200.h).  Importing the
package does not need a GPU; calling any compute function without the built library or
without a Blackwell device raises -- there is no CPU fallback.
"""
from . import _shim  # noqa: F401

__all__ = ["_shim"]
__version__ = "0.1.0"
