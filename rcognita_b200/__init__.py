"""rcognita_b200 -- B200 (sm_100a) batched agent-environment engine for rcognita's hot path.

The compute path is ``librcg_b200.so`` (hand-written CUDA kernels behind the C ABI of
``include/rcg.h``); this package binds it with ctypes and mirrors the reference's Python
interface for the path.  There is no CPU fallback: importing the package without the built
library raises, and every compute call without a CUDA device raises.
"""
from . import _C  # noqa: F401  (loads librcg_b200.so; raises ImportError if it is missing)
from ._C import LIB_PATH, last_error  # noqa: F401

__version__ = "0.1.0"


def version() -> int:
    return int(_C.lib.rcg_version())


def launch_count() -> int:
    """Kernels launched by librcg_b200 since load / last reset."""
    return int(_C.lib.rcg_launch_count())


def reset_launch_count() -> None:
    _C.lib.rcg_reset_launch_count()
