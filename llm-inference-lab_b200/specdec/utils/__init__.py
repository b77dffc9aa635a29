"""reference: src/specdec/utils/__init__.py -- ``graph_capture`` and ``token_validation`` live here, the other utilities
stay the reference's."""
import pkgutil as _pkgutil

# Merge with a same-named package further down sys.path (the reference's src/ when this directory is put in front
# of it): submodules that exist here win, everything else resolves to the reference.
__path__ = _pkgutil.extend_path(__path__, __name__)

from .token_validation import get_vocab_size, validate_and_clamp_tokens, validation_report  # noqa: E402,F401

_REFERENCE_EXPORTS = {"ensure_deterministic": ".deterministic", "LanguageModel": ".interfaces",
                      "SpeculativeDecoder": ".interfaces"}


def __getattr__(name):
    """Public names of the reference package that live in ITS submodules (resolved through the merged path)."""
    import importlib

    target = _REFERENCE_EXPORTS.get(name)
    if target is None:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
    try:
        return getattr(importlib.import_module(target, __name__), name)
    except ModuleNotFoundError as e:  # the reference tree is not on sys.path
        raise AttributeError(f"{__name__}.{name} needs the reference's {target} on sys.path") from e
