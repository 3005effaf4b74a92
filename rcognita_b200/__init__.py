"""rcognita_b200 -- B200 (sm_100a) batched agent-environment engine for rcognita's hot path.

The compute path is ``librcg_b200.so`` (hand-written CUDA kernels behind the C ABI of
``include/rcg.h``); this package binds it with ctypes (``_C``) and mirrors the reference's
Python interface for the path (``systems``, ``simulator``, ``controllers``).  There is no CPU
fallback: every module that computes imports ``_C``, which raises if the library has not been
built (``python -m rcognita_b200.build``), and every compute call raises without a CUDA device.
Only ``rcognita_b200.build`` is importable before the library exists.
"""
__version__ = "0.1.0"


def version() -> int:
    from . import _C
    return int(_C.lib.rcg_version())


def launch_count() -> int:
    """Kernels launched by librcg_b200 since load / last reset."""
    from . import _C
    return int(_C.lib.rcg_launch_count())


def reset_launch_count() -> None:
    from . import _C
    _C.lib.rcg_reset_launch_count()


def last_actor_kernel() -> str:
    """Kernel variant the last ``rcg_actor_cost`` call of this thread dispatched to (``rcg_last_actor_kernel``)."""
    from . import _C
    return (_C.lib.rcg_last_actor_kernel() or b"").decode()


def last_actor_opt_kernel() -> str:
    """Kernel variant the last ``rcg_actor_opt`` call of this thread dispatched to (``rcg_last_actor_opt_kernel``)."""
    from . import _C
    return (_C.lib.rcg_last_actor_opt_kernel() or b"").decode()


def actor_opt_lanes(lanes: int = 0) -> int:
    """Select the ``rcg_actor_opt`` kernel variant (``rcg_actor_opt_lanes``): 0 / 4 = four lanes per problem where an
    instantiation exists (default), 1 = the one-lane kernel.  Returns the previous setting."""
    from . import _C
    return int(_C.lib.rcg_actor_opt_lanes(int(lanes)))


def last_error() -> str:
    from . import _C
    return _C.last_error()
