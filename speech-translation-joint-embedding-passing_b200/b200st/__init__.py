"""b200st — Python binding of the hand-written sm_100a kernels for the joint speech-translation hot path."""
from . import runtime
from .runtime import set_compute_dtype, compute_dtype, manual_seed, join_deferred

__all__ = ['runtime', 'set_compute_dtype', 'compute_dtype', 'manual_seed', 'join_deferred']
