"""reference: src/specdec/cache/__init__.py -- same exports, arena-backed (plus ``KVArena``)."""
import pkgutil as _pkgutil

# Merge with a same-named package further down sys.path (the reference's src/ when this directory is put in front
# of it): submodules that exist here win, everything else resolves to the reference.
__path__ = _pkgutil.extend_path(__path__, __name__)

from .kv_arena import KVArena  # noqa: E402,F401
from .kv_cache_manager import SafeKVCacheManager  # noqa: E402,F401
from .kv_types import KVCache, validate_kv_compatibility  # noqa: E402,F401

__all__ = ["SafeKVCacheManager", "KVCache", "validate_kv_compatibility", "KVArena"]
