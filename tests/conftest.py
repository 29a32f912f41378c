"""Shared test helpers.  `gpu` marks tests that need a B200; everything else runs on CPU."""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def unflatten(npz):
    """Inverse of tests/golden/make_golden.py:flatten."""
    tree: dict = {}
    for key in npz.files:
        parts = key.split("/")
        node = tree
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = npz[key]

    def fix(node):
        if not isinstance(node, dict):
            return node
        if "__none__" in node:
            return None
        if "__str__" in node:
            return str(node["__str__"])
        if "__len__" in node:
            return [fix(node[str(i)]) for i in range(int(node["__len__"]))]
        return {k: fix(v) for k, v in node.items()}

    return fix(tree)


def load_golden(name):
    return unflatten(np.load(os.path.join(GOLDEN, f"{name}.npz")))


def fix_steps(steps):
    """Fixture flow steps -> oracle layout (parity as int, kind as str)."""
    out = []
    for st in steps:
        st = dict(st)
        if "parity" in st:
            st["parity"] = int(st["parity"])
        out.append(st)
    return out


def fix_params(P):
    P = dict(P)
    P["flow_q"] = fix_steps(P["flow_q"])
    P["flow_r"] = fix_steps(P["flow_r"])
    return P


def split_draws(draws):
    normals = [d["value"] for d in draws if d["kind"] == "normal"]
    berns = [d["value"] for d in draws if d["kind"] == "bernoulli"]
    return normals, berns


@pytest.fixture(scope="session")
def golden():
    return load_golden
