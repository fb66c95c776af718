"""mplan2vdl_b200: a B200-native executor for the Voodoo dataflow graphs emitted by orm011/mplan2vdl.

The hot path is libvdl_cuda (hand-written sm_100a CUDA behind the C ABI in include/vdl_cuda.h); this
package is the host-side mirror of that interface plus the metadata loaders and the synthetic recipe.
"""
from .lib import VdlError  # noqa: F401

__all__ = ["VdlError"]
