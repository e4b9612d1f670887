"""Shared helpers for the test-suite (oracle access is allowed here: tests/ is checker territory)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
VECTOR_MAPS = ["basement_hallways_10cm", "basement_hallways_5cm", "small.map", "basement_fixed_rectangle"]
KINDS = {"bl": 0, "rm": 1, "cddt": 2, "pcddt": 3}


def golden(name):
    return np.load(os.path.join(GOLD, "vectors_%s.npz" % name))


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a.view(np.uint64) if a.dtype == np.float64 else a


def assert_bit_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    bad = np.nonzero(bits(a) != bits(b))[0]
    assert bad.size == 0, "%s: %d of %d differ, first at %d: %r vs %r" % (
        what, bad.size, a.size, bad[0], a.ravel()[bad[0]], b.ravel()[bad[0]])


def world_tuple(v):
    return tuple(float(x) for x in v)
