"""Stub: tf_mnf/flows/maf.py imports tensorflow_probability at module scope for the
TFMAF/TFIAF wrappers, which are off the hot path."""
import types


class _Unavailable:
    def __init__(self, *a, **k):
        raise RuntimeError("tensorflow_probability is not available (stub)")


bijectors = types.SimpleNamespace(
    MaskedAutoregressiveFlow=_Unavailable, AutoregressiveNetwork=_Unavailable
)
