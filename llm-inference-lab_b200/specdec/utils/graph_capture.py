"""CUDA-graph capture of the verify step -- drop-in for ``src/specdec/utils/graph_capture.py``.

Same API as the reference (graph_capture.py:14-146: ``GraphCapture(enabled)``, ``can_capture``,
``capture_once(fn, sample_inputs)``, ``replay``, ``is_ready``, ``reset`` and the module-level helpers on a
global instance), which the reference defines but never wires in.  Differences that make
it usable for this path:

* capture happens on a side stream after one eager warm-up call (module loading, lazy allocations and
  the workspace zeroing must not land inside the graph);
* the kernels of this library carry programmatic-dependent-launch edges inside the graph, so a captured
  verify -> kv_append pair overlaps the second launch with the tail of the first;
* ``replay`` returns the value ``fn`` returned at capture time (static output buffers), not the inputs;
* failures raise (there is nothing to fall back to) unless ``strict=False``.
"""

from __future__ import annotations

import logging
import os
from typing import Any, Callable, Optional

import torch

logger = logging.getLogger(__name__)


class GraphCapture:
    def __init__(self, enabled: Optional[bool] = None, strict: bool = True) -> None:
        if enabled is None:
            enabled = os.getenv("SPECDEC_CUDA_GRAPH", "0") == "1"
        self.strict = strict
        self.enabled = bool(enabled) and self.can_capture()
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.captured = False
        self.sample_inputs: Any = None
        self.outputs: Any = None

    def can_capture(self) -> bool:
        """reference: graph_capture.py:37-47 (the grad-mode check is dropped: this path has no autograd)."""
        return torch.cuda.is_available()

    def capture_once(self, fn: Callable, sample_inputs: Any = None) -> bool:
        if not self.enabled or self.captured:
            return False
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn(sample_inputs)  # eager warm-up
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.outputs = fn(sample_inputs)
            self.captured = True
            self.sample_inputs = sample_inputs
            return True
        except Exception as e:  # noqa: BLE001
            self.graph, self.captured = None, False
            if self.strict:
                raise
            logger.warning("CUDA graph capture failed: %s", e)
            return False

    def replay(self) -> Any:
        if not self.is_ready():
            raise RuntimeError("No graph captured or graph capture disabled")
        self.graph.replay()
        return self.outputs

    def is_ready(self) -> bool:
        return self.enabled and self.captured and self.graph is not None

    def reset(self) -> None:
        self.graph, self.captured, self.sample_inputs, self.outputs = None, False, None, None


_graph_capture: Optional[GraphCapture] = None


def get_graph_capture() -> GraphCapture:
    global _graph_capture
    if _graph_capture is None:
        _graph_capture = GraphCapture()
    return _graph_capture


def _on_global_instance(method: str):
    """Module-level helper forwarding to the process-wide instance (reference: graph_capture.py:124-146)."""

    def call(*args: Any, **kwargs: Any) -> Any:
        return getattr(get_graph_capture(), method)(*args, **kwargs)

    call.__name__ = method
    call.__doc__ = f"``get_graph_capture().{method}(...)``"
    return call


can_capture, capture_once, replay, is_ready, reset = (
    _on_global_instance(m) for m in ("can_capture", "capture_once", "replay", "is_ready", "reset")
)
