"""Pins oracle/rangelib_oracle.c against the UNMODIFIED reference compiled into
oracle/_ref/libref_strict.so, on fresh seeded inputs (not the committed vectors).
Skipped when the reference library has not been built (it is built from /root/reference by
oracle/Makefile in the authoring container and travels to the GPU box as a binary)."""
import numpy as np
import pytest

from oracle import port, ref
from range_libc_b200 import workloads as wl
from helpers import assert_bit_equal

pytestmark = pytest.mark.skipif(not ref.available("strict"), reason="oracle/_ref not built")


@pytest.mark.parametrize("name", ["basement_hallways_10cm", "basement_fixed_rectangle", "small.map"])
@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_fresh_queries_bit_equal(name, kind):
    occ = wl.load_map(name)
    W, H = occ.shape
    q = wl.random_queries(W, H, 20000, seed=1000 + kind)
    rmap = ref.RefMap(occ=occ)
    r = ref.RefMethod(kind, rmap, 500.0, 108)
    o = port.Oracle(kind, occ, 500.0, 108)
    assert_bit_equal(o.calc_range_many(q), r.calc_range_many(q), "kind %d" % kind)
    if kind == 1:
        assert_bit_equal(o.dt(), r.dt(), "dt")
    if kind >= 2:
        for a, b in zip(o.cddt_table(), r.cddt_table(108)):
            assert_bit_equal(a, b, "cddt table")


def test_edge_map_equal():
    occ = wl.load_map("basement_hallways_10cm")
    assert (port.edge_map(occ) == ref.RefMap(occ=occ).edge()).all()


def test_other_theta_discretizations():
    occ = wl.load_map("basement_hallways_10cm")
    q = wl.random_queries(600, 600, 5000, seed=77)
    for td in (7, 16, 360):
        r = ref.RefMethod(ref.PCDDT, ref.RefMap(occ=occ), 300.0, td)
        o = port.Oracle(port.PCDDT, occ, 300.0, td)
        assert_bit_equal(o.calc_range_many(q), r.calc_range_many(q), "td %d" % td)
        for a, b in zip(o.cddt_table(), r.cddt_table(td)):
            assert_bit_equal(a, b, "table td %d" % td)


def test_multithreaded_slicing_is_identical():
    occ = wl.load_map("basement_hallways_10cm")
    q = wl.random_queries(600, 600, 5000, seed=78)
    a = port.Oracle(1, occ, 500.0, threads=1).numpy_calc_range(q)
    b = port.Oracle(1, occ, 500.0, threads=4).numpy_calc_range(q)
    assert_bit_equal(a, b)
    rmap = ref.RefMap(occ=occ)
    c = ref.RefMethod(1, rmap, 500.0, threads=4).numpy_calc_range(q)
    assert_bit_equal(a, c)


def test_giant_lut_cast():
    """GiantLUTCast (RangeLib.h:1772-1904): table and queries, incl. out-of-map and wrapped headings."""
    occ = wl.load_map("basement_hallways_10cm")[200:360, 250:400].copy()
    q = wl.random_queries(occ.shape[0], occ.shape[1], 20000, seed=3)
    q[:10, 0] = [-1, 0, 159.9, 160, 161, 5, 5, 5, 5, 5]
    q[:10, 2] = [0, 1, 2, 3, 4, -7, 7, 6.2831855, 6.28, 100]
    for td in (16, 108):
        r = ref.RefMethod(ref.GLT, ref.RefMap(occ=occ), 300.0, td)
        o = port.Oracle(port.GLT, occ, 300.0, td)
        assert np.array_equal(r.glt_table(td), o.glt_table())
        assert_bit_equal(o.calc_range_many(q), r.calc_range_many(q), "glt td %d" % td)


@pytest.mark.skipif(not ref.available("shipped"), reason="oracle/_ref shipped flavour not built")
def test_noise_floor_against_the_shipped_flag_build():
    """The bit-exactness target is the reference built STRICT (no FP contraction, no fast-math).  The same
    reference built with its own flags (-O3 -ffast-math + FMA) differs from that in the last bits: this records
    the reference's own noise floor and keeps it inside the tolerances north_star states
    (RM within 1e-4 relative for all but a vanishing fraction of rays; BL / CDDT likewise)."""
    occ = wl.load_map("basement_hallways_5cm")
    q = wl.random_queries(1200, 1200, 200000, seed=2025)
    shipped_map = ref.RefMap(occ=occ, flavor="shipped")
    report = {}
    for kind, name in ((ref.BL, "bl"), (ref.RM, "rm"), (ref.CDDT, "cddt")):
        a = port.Oracle(kind, occ, 500.0, 108, threads=8).calc_range_many(q)
        b = ref.RefMethod(kind, shipped_map, 500.0, 108, threads=8).calc_range_many(q)
        rel = np.abs(a - b) / np.maximum(np.abs(a), 1e-6)
        report[name] = (float((a.view(np.uint32) != b.view(np.uint32)).mean()), float((rel > 1e-4).mean()),
                        float((np.abs(a - b) > 1.0).mean()))
    print("STRICT vs SHIPPED (bit-mismatch, >1e-4 rel, >1 px):", report)
    assert report["rm"][1] < 1e-4 and report["bl"][1] < 1e-4 and report["cddt"][1] < 2e-3
    assert all(v[2] < 1e-4 for v in report.values())


def _random_small_map(rng):
    W, H = int(rng.integers(3, 70)), int(rng.integers(3, 70))
    occ = (rng.random((W, H)) < rng.choice([0.02, 0.1, 0.3])).astype(np.uint8)
    if rng.random() < 0.5:  # a wall or two
        occ[int(rng.integers(0, W)), :] = 1
        occ[:, int(rng.integers(0, H))] = 1
    return occ


def test_fuzz_small_random_maps_rm_cddt():
    """Random small non-square maps, odd theta discretizations and max ranges: the restatement against the
    unmodified reference for RM / CDDT / PCDDT (queries inside the map -- outside it the reference's CDDT indexes
    its grid out of bounds, RangeLib.h:1413), plus distance transform and tables."""
    rng = np.random.default_rng(20261017)
    for it in range(40):
        occ = _random_small_map(rng)
        W, H = occ.shape
        mr = float(rng.choice([3.5, 20.0, 50.0, 500.0]))
        td = int(rng.choice([4, 7, 16, 108, 361]))
        q = wl.random_queries(W, H, 400, seed=it)
        q[:, 0] = np.clip(q[:, 0], 0.01, W - 1.01)
        q[:, 1] = np.clip(q[:, 1], 0.01, H - 1.01)
        q[:8, 2] = [0, np.pi / 2, np.pi, -np.pi, 2 * np.pi, 7.5, -9.0, 1e-4]
        for kind in (ref.RM, ref.CDDT, ref.PCDDT):
            r = ref.RefMethod(kind, ref.RefMap(occ=occ), mr, td)
            o = port.Oracle(kind, occ, mr, td)
            what = "iter %d kind %d map %dx%d mr %g td %d" % (it, kind, W, H, mr, td)
            assert_bit_equal(o.calc_range_many(q), r.calc_range_many(q), what)
            if kind == ref.RM:
                assert_bit_equal(o.dt(), r.dt(), what + " dt")
            else:
                for a, b in zip(o.cddt_table(), r.cddt_table(td)):
                    assert_bit_equal(a, b, what + " table")


def test_fuzz_radial_optimized_and_giant_lut_small_maps():
    """calc_range_many_radial_optimized (pairs from CDDTCast::calc_range_pair) and GiantLUTCast on random small maps
    against the unmodified reference.  Poses stay inside the map (the reference indexes its tables without bounds
    checks outside it) and its output buffer is padded: it writes second beams past the row (RangeLib.h:665)."""
    rng = np.random.default_rng(77)
    for it in range(25):
        occ = _random_small_map(rng)
        W, H = occ.shape
        mr = float(rng.choice([20.0, 50.0, 500.0]))
        td = int(rng.choice([8, 16, 108, 120]))
        n = 40
        parts = np.empty((n, 3), np.float32)
        parts[:, 0] = rng.uniform(0.5, H - 1.5, n)  # world x runs along the map's height (x/y swap, RangeLib.h:475)
        parts[:, 1] = rng.uniform(0.5, W - 1.5, n)
        parts[:, 2] = rng.uniform(-7, 7, n)
        num_rays = int(rng.choice([7, 12, 37, 60, 100]))
        lo, hi = sorted(rng.uniform(-3.1, 3.1, 2))
        if hi - lo < 0.5:
            hi = lo + 0.5
        for kind in (ref.CDDT, ref.PCDDT, ref.RM):
            r = ref.RefMethod(kind, ref.RefMap(occ=occ), mr, td)
            o = port.Oracle(kind, occ, mr, td)
            pad = 8 * num_rays + 4096
            a = np.full(n * num_rays + pad, -7.0, np.float32)
            b = a.copy()
            r.calc_range_many_radial_optimized(num_rays, float(lo), float(hi), parts, a)
            o.calc_range_many_radial_optimized(num_rays, float(lo), float(hi), parts, b)
            # rows the reference overran into are rewritten by their own particle, except past the last row
            assert_bit_equal(b[: n * num_rays], a[: n * num_rays],
                             "iter %d kind %d %dx%d rays %d fan [%.2f, %.2f]" % (it, kind, W, H, num_rays, lo, hi))
        if W * H * td <= 400000:
            q = wl.random_queries(W, H, 300, seed=it)
            r = ref.RefMethod(ref.GLT, ref.RefMap(occ=occ), mr, td)
            o = port.Oracle(port.GLT, occ, mr, td)
            assert np.array_equal(r.glt_table(td), o.glt_table()), "glt table iter %d" % it
            assert_bit_equal(o.calc_range_many(q), r.calc_range_many(q), "glt iter %d" % it)
