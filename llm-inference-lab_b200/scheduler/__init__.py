"""reference: src/scheduler/__init__.py -- same exports, plus the batch-sharding helpers."""
import pkgutil as _pkgutil

# Merge with a same-named package further down sys.path (the reference's src/ when this directory is put in front
# of it): submodules that exist here win, everything else resolves to the reference.
__path__ = _pkgutil.extend_path(__path__, __name__)

from .batch_shard import BatchShard, PeerExchange, StepResultExchange, advance_sequences  # noqa: E402,F401


def __getattr__(name):
    # imported lazily: the scheduler module pulls in the kernel library
    if name in ("SpeculativeScheduler", "create_speculative_scheduler"):
        from . import speculative_scheduler as _s

        return getattr(_s, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


__all__ = ["SpeculativeScheduler", "create_speculative_scheduler", "BatchShard", "StepResultExchange", "PeerExchange",
           "advance_sequences"]
