"""Drop-in facade: `import tf_mnf` resolves to the B200-native implementation (tf_mnf_b200),
keeping the reference's module paths (tf_mnf.layers, tf_mnf.flows, tf_mnf.models)."""
import os

from tf_mnf_b200 import set_seed  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tf_mnf/__init__.py:4
