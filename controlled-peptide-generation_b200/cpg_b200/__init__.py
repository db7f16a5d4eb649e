"""cpg_b200: ctypes binding + functional layer over libcpg_b200.so (sm_100a kernels for
the WAE-training / CLaSS-sampling hot path of IBM/controlled-peptide-generation)."""
from . import _lib  # noqa: F401
from ._lib import CpgLibraryError  # noqa: F401
