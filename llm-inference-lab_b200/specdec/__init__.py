"""B200 drop-ins for the parts of the reference's ``specdec`` package that sit on the verify path
(``cache``, ``policies``, ``utils.graph_capture``) plus ``verify_step``; merges with the reference's package."""
import pkgutil as _pkgutil

# Merge with a same-named package further down sys.path (the reference's src/ when this directory is put in front
# of it): submodules that exist here win, everything else resolves to the reference.
__path__ = _pkgutil.extend_path(__path__, __name__)

_REFERENCE_EXPORTS = {"SpeculativePipeline": ".core.pipeline", "DraftModel": ".models.draft_model",
                      "create_fake_lm": ".models.fake_lm", "Verifier": ".verifier"}


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
