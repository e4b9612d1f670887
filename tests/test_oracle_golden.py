"""The CPU restatement (oracle/rangelib_oracle.c) against the committed golden vectors that were
produced by the unmodified reference (tests/golden/make_golden.py).  Bit-exact everywhere."""
import hashlib

import numpy as np
import pytest

from oracle import port
from range_libc_b200 import workloads as wl
from helpers import KINDS, VECTOR_MAPS, assert_bit_equal, golden, world_tuple


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", VECTOR_MAPS)
@pytest.mark.parametrize("kn", ["bl", "rm", "cddt", "pcddt"])
def test_oracle_matches_reference_vectors(name, kn):
    g = golden(name)
    occ = wl.load_map(name)
    o = port.Oracle(KINDS[kn], occ, 500.0, 108)
    assert_bit_equal(o.calc_range_many(g["queries"]), g[kn + "_grid"], kn + " grid")
    assert_bit_equal(o.numpy_calc_range_angles(g["particles"], g["angles"]), g[kn + "_angles"], kn + " angles")
    o.set_sensor_model(wl.sensor_table(501))
    assert_bit_equal(o.calc_range_repeat_angles_eval_sensor_model(g["particles"], g["angles"], g["obs"]),
                     g[kn + "_weights_fused"], kn + " fused")
    assert_bit_equal(o.eval_sensor_model(g["obs"], g[kn + "_angles"], len(g["angles"]), len(g["particles"])),
                     g[kn + "_weights_two_step"], kn + " two-step")
    o.set_world(*world_tuple(g["world"]))
    assert_bit_equal(o.numpy_calc_range(g["queries_world"]), g[kn + "_world"], kn + " world")
    o.set_world(*world_tuple(g["world_rot"]))
    assert_bit_equal(o.numpy_calc_range(g["queries_world_rot"]), g[kn + "_world_rot"], kn + " world_rot")
    if kn == "rm":
        dt = o.dt()
        assert sha(dt) == str(g["dt_sha256"])
        W, H = occ.shape
        assert_bit_equal(dt[:: max(1, W // 37), :: max(1, H // 41)], g["dt_sample"], "dt sample")
    if kn in ("cddt", "pcddt"):
        widths, trans, offsets, values = o.cddt_table()
        assert (widths == g[kn + "_widths"]).all()
        assert_bit_equal(trans, g[kn + "_trans"], "trans")
        assert len(values) == int(g[kn + "_nvalues"])
        assert sha(offsets) == str(g[kn + "_offsets_sha256"])
        assert sha(values) == str(g[kn + "_values_sha256"])


@pytest.mark.parametrize("name", ["quad.map", "single_pixel.map"])
def test_degenerate_maps_do_not_crash(name):
    occ = wl.load_map(name)
    q = wl.random_queries(occ.shape[0], occ.shape[1], 100, seed=3)
    for kind in range(4):
        r = port.Oracle(kind, occ, 500.0, 108).calc_range_many(q)
        assert np.isfinite(r).all()


def test_restated_glibc_trig_matches_libm_sample():
    """orc_sinf/orc_cosf (the algorithm the device trig restates) == libm sinf/cosf, bit for bit, on
    a strided sweep of all float bit patterns plus every float in [1, 8)."""
    assert port.trig_compare(0, 0xFFFFFFFF, 4099) == (0, 0)
    lo = int(np.float32(1.0).view(np.uint32))
    hi = int(np.float32(8.0).view(np.uint32))
    assert port.trig_compare(lo, hi, 7) == (0, 0)


@pytest.mark.parametrize("name", ["basement_hallways_10cm", "basement_hallways_5cm"])
@pytest.mark.parametrize("kn", ["bl", "rm", "cddt", "pcddt"])
def test_radial_optimized_matches_reference_vectors(name, kn):
    """calc_range_many_radial_optimized + CDDTCast::calc_range_pair (RangeLib.h:616-676, :1521-1649) against
    tests/golden/vectors_radial.npz (made by the unmodified reference, make_golden_radial.py)."""
    import os
    from helpers import GOLD
    g = np.load(os.path.join(GOLD, "vectors_radial.npz"))
    occ = wl.load_map(name)
    for wname in ("id", "rot"):
        o = port.Oracle(KINDS[kn], occ, 500.0, 108)
        if wname == "rot":
            o.set_world(*world_tuple(g["world_rot"]))
        ins = g[name + "/particles"] if wname == "id" else g[name + "/particles_rot"]
        for ci, (n, lo, hi) in enumerate(g["configs"]):
            n = int(n)
            outs = np.full(len(ins) * n, g["fill"], np.float32)
            o.calc_range_many_radial_optimized(n, float(lo), float(hi), ins, outs)
            assert_bit_equal(outs, g["%s/%s/%s/%d" % (name, kn, wname, ci)], "%s %s cfg %d" % (kn, wname, ci))
