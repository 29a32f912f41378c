"""Noise plans: which Philox sub-stream each random draw of the hot path uses, and how a
draw is replaced by an injected tensor (north_star (d); draw order SURVEY.md App. A-6)."""
from __future__ import annotations

import ctypes as C
from typing import Any

from . import _lib as L

# draw ids inside one layer (low byte of the stream word; the layer id sits above it)
CALL_Z, CALL_OUT, KL_Z, KL_A, KL_B = 0, 1, 2, 3, 4
BERN_Q_CALL, BERN_Q_KL, BERN_R_KL = 16, 48, 80  # + flow step index
FLOW_BERN = 112  # stand-alone NormalizingFlow.forward calls


def stream_id(layer_id: int, draw: int) -> int:
    return ((layer_id & 0xFFFF) << 8) | (draw & 0xFF)


# While a step is being captured into a CUDA graph (graph.CapturedStep) every noise plan points at
# the graph's device-side replay counter, so each replay draws with call index + replay number.
_capture_delta: int | None = None


def make(seed: int, call_index: int, layer_id: int, draw: int, row_offset: int = 0, injected: Any = None) -> L.Noise:
    """seed: user seed (low 32 bits are the Philox key low word); call_index makes every
    call draw fresh noise, as tf.random.normal does."""
    n = L.Noise()
    n.seed = (int(seed) & 0xFFFFFFFF) | ((int(call_index) & 0xFFFFFFFF) << 32)
    n.row_offset = int(row_offset)
    n.stream = stream_id(layer_id, draw)
    n.reserved = 0
    n.injected = C.c_void_p(injected.ptr) if injected is not None else None
    n.call_delta = C.c_void_p(_capture_delta) if (_capture_delta and injected is None) else None
    return n
