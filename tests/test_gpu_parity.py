"""GPU parity: the CUDA path (through the C ABI) against the committed golden vectors of the
unmodified reference and against the CPU oracle on fresh seeded inputs.

Bars (north_star): BL / CDDT / PCDDT ranges bit-exact; RM bit-exact against the STRICT oracle
(stated tolerance: 0 differing rays -- device trig restates glibc's sinf/cosf in double);
distance transform and CDDT/PCDDT tables bit-exact; sensor-model weights bit-exact (the product
is formed in the reference's order), which is inside the required 1e-5 relative."""
import numpy as np
import pytest

import range_libc_b200 as rl
from range_libc_b200 import api, workloads as wl
from oracle import port
from helpers import KINDS, VECTOR_MAPS, assert_bit_equal, golden, world_tuple

pytestmark = pytest.mark.gpu

MR, TD = 500.0, 108


def make(kn, occ, max_range=MR, td=TD, world=None):
    m = rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))  # PyOMap takes arr[row=y, col=x]
    if world is not None:
        m.set_world(*world)
    if kn == "bl":
        return rl.PyBresenhamsLine(m, max_range)
    if kn == "rm":
        return rl.PyRayMarchingGPU(m, max_range)
    c = rl.PyCDDTCast(m, max_range, td)
    if kn == "pcddt":
        c.prune()
    return c


def test_device_trig_equals_libm():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-20, 20, 400000), rng.uniform(-130, 130, 100000), rng.uniform(-1e6, 1e6, 20000),
                        rng.uniform(-1e-3, 1e-3, 1000), np.array([0.0, -0.0, np.pi, 0.75, 0.7853982, 119.99, 120.0, 1e30])
                        ]).astype(np.float32)
    s, c = api.device_sincosf(x)
    es = np.array([port.sinf(v) for v in x[:2000]], np.float32)
    assert_bit_equal(s[:2000], es, "restated sinf")
    # numpy's float32 sin/cos are not glibc's; compare against libm through the oracle library
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.sinf.restype = ctypes.c_float
    libm.sinf.argtypes = [ctypes.c_float]
    libm.cosf.restype = ctypes.c_float
    libm.cosf.argtypes = [ctypes.c_float]
    idx = rng.integers(0, len(x), 50000)
    idx[-8:] = np.arange(len(x) - 8, len(x))
    assert_bit_equal(s[idx], np.array([libm.sinf(float(v)) for v in x[idx]], np.float32), "sinf vs libm")
    assert_bit_equal(c[idx], np.array([libm.cosf(float(v)) for v in x[idx]], np.float32), "cosf vs libm")


@pytest.mark.parametrize("name", VECTOR_MAPS)
@pytest.mark.parametrize("kn", ["bl", "rm", "cddt", "pcddt"])
def test_golden_vectors(name, kn):
    g = golden(name)
    occ = wl.load_map(name)
    meth = make(kn, occ)
    q = np.ascontiguousarray(g["queries"])
    out = np.empty(len(q), np.float32)
    meth.calc_range_many_grid(q, out)
    assert_bit_equal(out, g[kn + "_grid"], kn + " grid")
    parts, angles, obs = (np.ascontiguousarray(g[k]) for k in ("particles", "angles", "obs"))
    out = np.empty(len(parts) * len(angles), np.float32)
    meth.calc_range_repeat_angles(parts, angles, out)
    assert_bit_equal(out, g[kn + "_angles"], kn + " angles")
    meth.set_sensor_model(wl.sensor_table(501))
    w = np.empty(len(parts), np.float64)
    meth.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs, w)
    assert_bit_equal(w, g[kn + "_weights_fused"], kn + " fused weights")
    w2 = np.empty(len(parts), np.float64)
    meth.eval_sensor_model(obs, np.ascontiguousarray(g[kn + "_angles"]), w2, len(angles), len(parts))
    assert_bit_equal(w2, g[kn + "_weights_two_step"], kn + " two-step weights")
    for tag in ("world", "world_rot"):
        mw = make(kn, occ, world=world_tuple(g[tag]))
        qw = np.ascontiguousarray(g["queries_" + tag])
        out = np.empty(len(qw), np.float32)
        mw.calc_range_many(qw, out)
        assert_bit_equal(out, g[kn + "_" + tag], kn + " " + tag)
    if kn == "rm":
        dt = meth.distance_transform()
        W, H = occ.shape
        assert_bit_equal(dt[:: max(1, W // 37), :: max(1, H // 41)], g["dt_sample"], "dt sample")
    if kn in ("cddt", "pcddt"):
        widths, trans, offsets, values = meth.table()
        assert (widths == g[kn + "_widths"]).all()
        assert_bit_equal(trans, g[kn + "_trans"], "translations")
        assert len(values) == int(g[kn + "_nvalues"])


@pytest.mark.parametrize("name", ["basement_hallways_5cm", "basement_fixed_rectangle", "synthetic.map", "small.map",
                                  "quad.map", "single_pixel.map"])
def test_structures_equal_oracle(name):
    """Distance transform and CDDT / PCDDT tables, whole arrays, bit for bit."""
    occ = wl.load_map(name)
    assert_bit_equal(make("rm", occ).distance_transform(), port.Oracle(port.RM, occ, MR).dt(), "dt")
    for kn, kind in (("cddt", port.CDDT), ("pcddt", port.PCDDT)):
        a = make(kn, occ).table()
        b = port.Oracle(kind, occ, MR, TD).cddt_table()
        for x, y, what in zip(a, b, ("widths", "trans", "offsets", "values")):
            assert_bit_equal(x, y, "%s %s" % (kn, what))


@pytest.mark.parametrize("kn", ["bl", "rm", "cddt", "pcddt"])
def test_fresh_queries_vs_oracle(kn):
    occ = wl.load_map("basement_hallways_5cm")
    W, H = occ.shape
    q = wl.random_queries(W, H, 200000, seed=2024)
    out = np.empty(len(q), np.float32)
    meth = make(kn, occ)
    meth.calc_range_many_grid(q, out)
    ref = port.Oracle(KINDS[kn], occ, MR, TD, threads=8).calc_range_many(q)
    assert_bit_equal(out, ref, kn)
    if kn in ("cddt", "pcddt"):  # the query index built for tables larger than L2, forced on this small one
        meth.set_spatial_sort(2)
        out2 = np.empty_like(out)
        meth.calc_range_many_grid(q, out2)
        assert_bit_equal(out2, ref, kn + " through the query index")


def test_device_pointers_and_stream():
    import torch
    occ = wl.load_map("basement_hallways_10cm")
    q = wl.random_queries(600, 600, 50000, seed=5)
    meth = make("rm", occ)
    host = np.empty(len(q), np.float32)
    meth.calc_range_many_grid(q, host)
    dq = torch.from_numpy(q).cuda()
    dout = torch.empty(len(q), dtype=torch.float32, device="cuda")
    meth.set_stream(torch.cuda.current_stream().cuda_stream)
    meth.calc_range_many_grid(dq, dout)
    torch.cuda.synchronize()
    assert_bit_equal(dout.cpu().numpy(), host, "device pointers")
    with pytest.raises(rl.RangeLibError):
        meth.calc_range_many_grid(dq, host)  # mixed host / device


def test_empty_and_ragged_inputs():
    occ = wl.load_map("basement_hallways_10cm")
    meth = make("rm", occ)
    meth.calc_range_many_grid(np.zeros((0, 3), np.float32), np.zeros(0, np.float32))
    meth.set_sensor_model(wl.sensor_table(501))
    # one particle, one beam; 3 particles x 70 beams (not a multiple of the warp size)
    for n, m in ((1, 1), (3, 70), (5, 2500)):
        parts = wl.pf_particles_uniform(occ, n, seed=n)
        angles = wl.lidar_angles(m)
        obs = np.linspace(0, 499, m).astype(np.float32)
        w = np.empty(n, np.float64)
        meth.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs, w)
        o = port.Oracle(port.RM, occ, MR)
        o.set_sensor_model(wl.sensor_table(501))
        assert_bit_equal(w, o.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs), "fused %dx%d" % (n, m))
    with pytest.raises(ValueError):
        meth.calc_range_many_grid(np.zeros((4, 3), np.float64), np.zeros(4, np.float32))
    # non-finite poses do not hang and return max_range
    bad = np.array([[np.nan, 1, 0], [1, np.inf, 0], [10, 10, np.nan]], np.float32)
    for kn in ("bl", "rm", "cddt"):
        out = np.empty(3, np.float32)
        make(kn, occ).calc_range_many_grid(bad, out)
        assert (out == MR).all()


def test_single_ray_api_and_sensor_without_table():
    occ = wl.load_map("basement_hallways_10cm")
    meth = make("cddt", occ)
    o = port.Oracle(port.CDDT, occ, MR, TD)
    for x, y, th in ((300.5, 290.25, 1.0), (100.0, 100.0, -2.0), (550.0, 20.0, 9.0)):
        assert np.float32(meth.calc_range(x, y, th)) == np.float32(o.calc_range(x, y, th))
    with pytest.raises(rl.RangeLibError):
        meth.eval_sensor_model(np.zeros(2, np.float32), np.zeros(4, np.float32), np.zeros(2, np.float64), 2, 2)


def test_dynamic_map_update_bl():
    occ = wl.synthetic_map(1024, seed=11)
    meth = make("bl", occ)
    q = wl.random_queries(1024, 1024, 50000, seed=12)
    for frame in range(3):
        for x0, y0, patch in wl.flip_blocks(occ, frame, seed=11, n_blocks=16):
            meth.update_map(patch, x0, y0)
        out = np.empty(len(q), np.float32)
        meth.calc_range_many_grid(q, out)
        assert_bit_equal(out, port.Oracle(port.BL, occ, MR, threads=8).calc_range_many(q), "frame %d" % frame)


def test_dynamic_map_update_batched():
    import torch
    q = wl.random_queries(1024, 1024, 50000, seed=14)
    for kn, kind in (("bl", port.BL), ("rm", port.RM)):
        occ = wl.synthetic_map(1024, seed=13)
        meth = make(kn, occ)
        for frame in range(2):
            blocks = wl.flip_blocks(occ, frame, seed=13, n_blocks=24, block=16)  # block aligned: no overlaps
            rects = np.array([[x0, y0, 16, 16] for x0, y0, _ in blocks], np.int32)
            flat = np.concatenate([p.ravel() for _, _, p in blocks])
            if frame == 0:
                meth.update_map_batch(flat, rects)  # host patches
            else:
                meth.update_map_batch(torch.from_numpy(flat).cuda(), rects)  # device patches
            out = np.empty(len(q), np.float32)
            meth.calc_range_many_grid(q, out)
            assert_bit_equal(out, port.Oracle(kind, occ, MR, threads=8).calc_range_many(q), "%s frame %d" % (kn, frame))


def test_dynamic_map_update_rm_and_cddt():
    occ = wl.synthetic_map(512, seed=21)
    q = wl.random_queries(512, 512, 20000, seed=22)
    rm, cd = make("rm", occ), make("pcddt", occ)
    for x0, y0, patch in wl.flip_blocks(occ, 0, seed=21, n_blocks=8):
        rm.update_map(patch, x0, y0)
        cd.update_map(patch, x0, y0)
    out = np.empty(len(q), np.float32)
    rm.calc_range_many_grid(q, out)
    assert_bit_equal(out, port.Oracle(port.RM, occ, MR, threads=8).calc_range_many(q), "rm after update")
    cd.calc_range_many_grid(q, out)
    assert_bit_equal(out, port.Oracle(port.PCDDT, occ, MR, TD, threads=8).calc_range_many(q), "pcddt after update")


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
def test_rm_persistent_kernel_large_batches(variant):
    """Batches large enough for the persistent-warp / lane re-queuing kernel, all three entry points.
    Variants: 1 default (parked rays in registers), 2 conditional load, 3 parked rays in shared memory,
    4 two rays per lane -- performance knobs, identical results."""
    occ = wl.load_map("basement_hallways_5cm")
    W, H = occ.shape
    world = (0.05, 0.0, -30.0, -30.0, 0.0, 1.0)
    meth = make("rm", occ, world=world)
    meth.set_persistent(variant)
    o = port.Oracle(port.RM, occ, MR, threads=8)
    o.set_world(*world)
    import torch
    n = 1_000_003  # not a multiple of anything
    q = wl.random_queries(W, H, n, seed=31)
    want = o.calc_range_many(q)
    out = np.empty(n, np.float32)
    meth.calc_range_many_grid(q, out)  # host arrays: staged through device memory
    assert_bit_equal(out, want, "grid (host)")
    d_out = torch.empty(n, dtype=torch.float32, device="cuda")
    meth.calc_range_many_grid(torch.from_numpy(q).cuda(), d_out)  # device arrays: one launch of the persistent kernel
    meth.synchronize()
    assert_bit_equal(d_out.cpu().numpy(), want, "grid (device)")
    qw = wl.grid_to_world(q, world[0], world[2], world[3])
    want = o.numpy_calc_range(qw)
    meth.calc_range_many(qw, out)
    assert_bit_equal(out, want, "world (host)")
    meth.calc_range_many(torch.from_numpy(qw).cuda(), d_out)
    meth.synchronize()
    assert_bit_equal(d_out.cpu().numpy(), want, "world (device)")
    parts = wl.grid_to_world(wl.random_queries(W, H, 9001, seed=32), world[0], world[2], world[3])
    angles = wl.lidar_angles(113)
    out = np.empty(len(parts) * len(angles), np.float32)
    meth.calc_range_repeat_angles(parts, angles, out)
    assert_bit_equal(out, o.numpy_calc_range_angles(parts, angles), "angles")
    meth.set_persistent(0)
    out2 = np.empty_like(out)
    meth.calc_range_repeat_angles(parts, angles, out2)
    assert_bit_equal(out2, out, "persistent vs one-ray-per-thread")


def test_rm_variants_agree():
    """The cooperative tail and the persistent kernel are pure performance knobs: identical results."""
    occ = wl.load_map("basement_hallways_5cm")
    q = wl.random_queries(1200, 1200, 300000, seed=77)
    meth = make("rm", occ)
    ref_out = port.Oracle(port.RM, occ, MR, threads=8).calc_range_many(q)
    for coop in (0, 1, 3, 8, 32):
        meth.set_coop_threshold(coop)
        out = np.empty(len(q), np.float32)
        meth.calc_range_many_grid(q, out)
        assert_bit_equal(out, ref_out, "coop=%d" % coop)
        # small batches leave most warps nearly empty: the cooperative path does almost all the work
        out = np.empty(37, np.float32)
        meth.calc_range_many_grid(q[:37], out)
        assert_bit_equal(out, ref_out[:37], "coop=%d small" % coop)


def test_cython_drop_in_module():
    """The re-pointed Cython module `range_libc` (pywrapper/RangeLibc.pyx): same names as the reference."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(rl.__file__), "pywrapper"))
    try:
        import range_libc
    except ImportError:
        pytest.skip("range_libc extension not built")
    occ = wl.load_map("basement_hallways_10cm")
    g = golden("basement_hallways_10cm")
    omap = range_libc.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
    assert omap.width() == 600 and omap.height() == 600 and not omap.error()
    for cls, kn, extra in ((range_libc.PyBresenhamsLine, "bl", ()), (range_libc.PyRayMarching, "rm", ()),
                           (range_libc.PyRayMarchingGPU, "rm", ()), (range_libc.PyCDDTCast, "cddt", (108,))):
        meth = cls(omap, 500.0, *extra)
        parts, angles, obs = (np.ascontiguousarray(g[k]) for k in ("particles", "angles", "obs"))
        out = np.empty(len(parts) * len(angles), np.float32)
        meth.calc_range_repeat_angles(parts, angles, out)
        assert_bit_equal(out, g[kn + "_angles"], kn + " angles via range_libc")
        meth.set_sensor_model(wl.sensor_table(501))
        w = np.empty(len(parts), np.float64)
        meth.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs, w)
        assert_bit_equal(w, g[kn + "_weights_fused"], kn + " fused via range_libc")
        w2 = np.empty(len(parts), np.float64)
        meth.eval_sensor_model(obs, out, w2, len(angles), len(parts))
        assert_bit_equal(w2, g[kn + "_weights_two_step"], kn + " two-step via range_libc")
        with pytest.raises(ValueError):
            meth.calc_range_many(np.zeros((4, 3), np.float64), np.zeros(4, np.float32))
    c = range_libc.PyCDDTCast(omap, 500.0, 108)
    c.prune()
    out = np.empty(len(g["queries"]), np.float32)
    # calc_range_many takes WORLD poses; with identity world params grid pose (x, y, th) is world (y, x, -th - 3pi/2)
    qw = wl.grid_to_world(g["queries"])
    c.calc_range_many(qw, out)
    o = port.Oracle(port.PCDDT, occ, MR, TD)
    assert_bit_equal(out, o.numpy_calc_range(qw), "pcddt via range_libc")


@pytest.mark.parametrize("name", ["huge_map", "gigantic_map"])
def test_big_maps_against_digests(name):
    """BASELINE config 3 sizes (10976^2, 4128x10976: sides > 4096, where the reference's distance transform
    is no longer the exact EDT): device-built distance transform, CDDT and PCDDT tables and ranges against the
    oracle's sha256 digests (tests/golden/big_digests.json, made by tests/golden/make_golden_big.py)."""
    import hashlib
    import json
    import os
    from helpers import GOLD
    dig = json.load(open(os.path.join(GOLD, "big_digests.json")))[name]

    def sha(a):
        return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()

    occ = wl.load_map(name)
    W, H = occ.shape
    assert (W, H) == (dig["width"], dig["height"])
    q = wl.random_queries(W, H, 20000, seed=777)
    out = np.empty(len(q), np.float32)
    rm = make("rm", occ)
    assert sha(rm.distance_transform()) == dig["dt_sha256"]
    rm.calc_range_many_grid(q, out)
    assert sha(out) == dig["rm_ranges_sha256"]
    del rm
    bl = make("bl", occ)
    bl.calc_range_many_grid(q, out)
    assert sha(out) == dig["bl_ranges_sha256"]
    del bl
    cd = make("cddt", occ)
    widths, trans, offsets, values = cd.table()
    assert len(values) == dig["cddt_nvalues"]
    assert sha(offsets) == dig["cddt_offsets_sha256"]
    assert sha(values) == dig["cddt_values_sha256"]
    cd.calc_range_many_grid(q, out)
    assert sha(out) == dig["cddt_ranges_sha256"]
    cd.prune()
    widths, trans, offsets, values = cd.table()
    assert len(values) == dig["pcddt_nvalues"]
    assert sha(offsets) == dig["pcddt_offsets_sha256"]
    assert sha(values) == dig["pcddt_values_sha256"]
    cd.calc_range_many_grid(q, out)
    assert sha(out) == dig["pcddt_ranges_sha256"]


def test_bl_large_batches_and_reference_nontermination():
    """Large BL batches incl. rays that start outside the map, and the ray on which the reference's float
    `_x += xstep` accumulation jumps over its loop target (the reference never returns for it; oracle and
    device both end the walk when it has left the map for good)."""
    occ = wl.load_map("basement_hallways_5cm")
    W, H = occ.shape
    world = (0.05, 0.0, -30.0, -30.0, 0.0, 1.0)
    meth = make("bl", occ, world=world)
    o = port.Oracle(port.BL, occ, MR, threads=8)
    o.set_world(*world)
    n = 400_003
    q = wl.random_queries(W, H, n, seed=41)
    assert abs(q[101113, 0] - 1188.2903) < 1e-3  # the non-terminating ray is part of this batch
    q[:64, 0] = np.linspace(-20, W + 20, 64)  # a few rays that start outside the map
    out = np.empty(n, np.float32)
    meth.calc_range_many_grid(q, out)
    assert_bit_equal(out, o.calc_range_many(q), "grid")
    qw = wl.grid_to_world(q, world[0], world[2], world[3])
    meth.calc_range_many(qw, out)
    assert_bit_equal(out, o.numpy_calc_range(qw), "world")


def test_giant_lut_cast_gpu():
    """GiantLUTCast: device-built W*H*td uint16 table and gather queries, bit-equal to the oracle; fused weights too."""
    occ = wl.load_map("basement_hallways_10cm")
    W, H = occ.shape
    td = 36
    m = rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
    glt = rl.PyGiantLUTCast(m, MR, td)
    o = port.Oracle(port.GLT, occ, MR, td)
    assert np.array_equal(glt.table(), o.glt_table())
    q = wl.random_queries(W, H, 100000, seed=3)
    q[:10, 0] = [-1, 0, 599.9, 600, 601, 5, 5, 5, 5, 5]
    q[:10, 2] = [0, 1, 2, 3, 4, -7, 7, 6.2831855, 6.28, 100]
    out = np.empty(len(q), np.float32)
    glt.calc_range_many_grid(q, out)
    assert_bit_equal(out, o.calc_range_many(q), "glt grid")
    parts = wl.pf_particles_uniform(occ, 300, seed=4)
    angles = wl.lidar_angles(45)
    obs = np.linspace(3, 480, 45).astype(np.float32)
    glt.set_sensor_model(wl.sensor_table(501))
    o.set_sensor_model(wl.sensor_table(501))
    w = np.empty(len(parts), np.float64)
    glt.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs, w)
    assert_bit_equal(w, o.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs), "glt fused")
    assert glt.memory() >= W * H * td * 2


def test_cpp_header_mirror_on_gpu(tmp_path):
    from test_abi_surface import test_cpp_header_mirror_compiles_and_fails_loudly_without_gpu as run
    run(tmp_path)


@pytest.mark.parametrize("kn", ["bl", "rm", "cddt", "pcddt"])
def test_rotated_world_fused_and_odd_parameters(kn):
    """Rotated ROS world frame through the angle-fan and fused entry points, a non-integer max_range, a sensor
    table whose width is not max_range + 1 (clamping), observations beyond the table, odd theta_discretization."""
    occ = wl.load_map("basement_fixed_rectangle")
    W, H = occ.shape
    ang = 0.4
    world = (0.07, ang, 1.5, -2.5, float(np.sin(ang)), float(np.cos(ang)))
    mr, td, K = 123.45, 7, 100
    meth = make(kn, occ, max_range=mr, td=td, world=world)
    o = port.Oracle(KINDS[kn], occ, mr, td, threads=4)
    o.set_world(*world)
    parts = wl.grid_to_world(wl.random_queries(W, H, 700, seed=61), world[0], world[2], world[3], ang)
    angles = wl.lidar_angles(33, fov=2.0)
    obs = np.linspace(-1.0, 12.0, 33).astype(np.float32)  # world units: up to 171 px >> K-1
    table = wl.sensor_table(K)
    meth.set_sensor_model(table)
    o.set_sensor_model(table)
    out = np.empty(len(parts) * len(angles), np.float32)
    meth.calc_range_repeat_angles(parts, angles, out)
    ref_ranges = o.numpy_calc_range_angles(parts, angles)
    assert_bit_equal(out, ref_ranges, kn + " angles")
    w = np.empty(len(parts), np.float64)
    meth.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs, w)
    assert_bit_equal(w, o.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs), kn + " fused")
    w2 = np.empty(len(parts), np.float64)
    meth.eval_sensor_model(obs, out, w2, len(angles), len(parts))
    assert_bit_equal(w2, o.eval_sensor_model(obs, ref_ranges, len(angles), len(parts)), kn + " two-step")


@pytest.mark.parametrize("name", ["quad.map", "single_pixel.map", "small.map"])
def test_degenerate_maps_all_kinds(name):
    occ = wl.load_map(name)
    W, H = occ.shape
    q = wl.random_queries(W, H, 2000, seed=71)
    q[:200, :2] = np.random.default_rng(1).uniform(-3, max(W, H) + 3, (200, 2))  # also outside the map
    for kn in ("bl", "rm"):
        out = np.empty(len(q), np.float32)
        make(kn, occ).calc_range_many_grid(q, out)
        assert_bit_equal(out, port.Oracle(KINDS[kn], occ, MR).calc_range_many(q), "%s %s" % (name, kn))
    qi = q[200:]  # CDDT reads map.grid[x][y] unchecked in the reference: in-map queries only
    for kn in ("cddt", "pcddt"):
        out = np.empty(len(qi), np.float32)
        make(kn, occ).calc_range_many_grid(qi, out)
        assert_bit_equal(out, port.Oracle(KINDS[kn], occ, MR, TD).calc_range_many(qi), "%s %s" % (name, kn))
    m = rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
    glt = rl.PyGiantLUTCast(m, MR, 12)
    out = np.empty(len(q), np.float32)
    glt.calc_range_many_grid(q, out)
    assert_bit_equal(out, port.Oracle(port.GLT, occ, MR, 12).calc_range_many(q), "%s glt" % name)


def test_empty_map_and_empty_batches():
    """Zero-sized maps and batches are legal inputs: every ray of an empty map leaves it at once."""
    q = np.array([[0.5, 0.5, 0.1], [3.0, -2.0, 2.0]], np.float32)
    for shape in ((0, 0), (0, 5), (7, 0)):
        m = rl.PyOMap(np.zeros(shape, bool))
        for ctor in (lambda: rl.PyRayMarchingGPU(m, 50.0), lambda: rl.PyBresenhamsLine(m, 50.0),
                     lambda: rl.PyCDDTCast(m, 50.0, 12), lambda: rl.PyGiantLUTCast(m, 50.0, 12)):
            meth = ctor()
            out = np.full(2, -1.0, np.float32)
            meth.calc_range_many_grid(q, out)
            assert (out == 50.0).all(), (shape, type(meth).__name__, out)
            meth.calc_range_many_grid(np.zeros((0, 3), np.float32), np.zeros(0, np.float32))
    occ = wl.load_map("small.map")
    meth = make("pcddt", occ)
    meth.set_sensor_model(wl.sensor_table(64))
    meth.calc_range_repeat_angles(np.zeros((0, 3), np.float32), np.zeros(4, np.float32), np.zeros(0, np.float32))
    w = np.zeros(0, np.float64)
    meth.calc_range_repeat_angles_eval_sensor_model(np.zeros((0, 3), np.float32), np.zeros(4, np.float32),
                                                    np.zeros(4, np.float32), w)


def test_bl_persistent_kernel_matches_one_ray_per_thread():
    """BL batches large enough for the persistent-warp kernel (lane re-queuing), all three entry points."""
    occ = wl.load_map("basement_hallways_5cm")
    W, H = occ.shape
    world = (0.05, 0.0, -30.0, -30.0, 0.0, 1.0)
    meth = make("bl", occ, world=world)
    o = port.Oracle(port.BL, occ, MR, threads=8)
    o.set_world(*world)
    n = 500_009
    q = wl.random_queries(W, H, n, seed=43)
    out = np.empty(n, np.float32)
    meth.calc_range_many_grid(q, out)
    assert_bit_equal(out, o.calc_range_many(q), "grid")
    parts = wl.grid_to_world(wl.random_queries(W, H, 7001, seed=44), world[0], world[2], world[3])
    angles = wl.lidar_angles(71)
    out = np.empty(len(parts) * len(angles), np.float32)
    meth.calc_range_repeat_angles(parts, angles, out)
    assert_bit_equal(out, o.numpy_calc_range_angles(parts, angles), "angles")
    meth.set_persistent(0)
    out2 = np.empty_like(out)
    meth.calc_range_repeat_angles(parts, angles, out2)
    assert_bit_equal(out2, out, "persistent vs one-ray-per-thread")


@pytest.mark.parametrize("kn", ["bl", "rm", "cddt", "pcddt"])
def test_radial_optimized(kn):
    """calc_range_many_radial_optimized (RangeLib.h:616-676; pairs from CDDTCast::calc_range_pair :1521-1649):
    bit-exact against the reference's golden vectors, including the beams the reference leaves unwritten, for
    host and device pointers; then against the oracle on a larger fresh batch."""
    import os
    import torch
    from helpers import GOLD
    g = np.load(os.path.join(GOLD, "vectors_radial.npz"))
    for name in ("basement_hallways_10cm", "basement_hallways_5cm"):
        occ = wl.load_map(name)
        for wname in ("id", "rot"):
            meth = make(kn, occ, world=world_tuple(g["world_rot"]) if wname == "rot" else None)
            ins = g[name + "/particles"] if wname == "id" else g[name + "/particles_rot"]
            for ci, (n, lo, hi) in enumerate(g["configs"]):
                n = int(n)
                want = g["%s/%s/%s/%d" % (name, kn, wname, ci)]
                outs = np.full(len(ins) * n, g["fill"], np.float32)
                meth.calc_range_many_radial_optimized(n, float(lo), float(hi), ins, outs)
                assert_bit_equal(outs, want, "%s %s %s cfg %d host" % (name, kn, wname, ci))
                d_out = torch.full((len(ins) * n,), float(g["fill"]), dtype=torch.float32, device="cuda")
                meth.calc_range_many_radial_optimized(n, float(lo), float(hi), torch.from_numpy(ins).cuda(), d_out)
                torch.cuda.synchronize()
                assert_bit_equal(d_out.cpu().numpy(), want, "%s %s %s cfg %d device" % (name, kn, wname, ci))
    occ = wl.load_map("basement_hallways_5cm")
    parts = wl.pf_particles_uniform(occ, 3000, seed=77)
    meth = make(kn, occ)
    o = port.Oracle(KINDS[kn], occ, MR, TD)
    for n, lo, hi in ((1081, -2.35619449615, 2.35619449615), (54, -0.75 * np.pi, 0.75 * np.pi)):
        got = np.full(len(parts) * n, -3.0, np.float32)
        want = got.copy()
        meth.calc_range_many_radial_optimized(n, lo, hi, parts, got)
        o.calc_range_many_radial_optimized(n, lo, hi, parts, want)
        assert_bit_equal(got, want, "%s fresh %d beams" % (kn, n))
    with pytest.raises(rl.RangeLibError):
        meth.calc_range_many_radial_optimized(1, 0.0, 1.0, parts, got)
    with pytest.raises(rl.RangeLibError):
        meth.calc_range_many_radial_optimized(10, 1.0, 1.0, parts, got)


def test_device_map_ingest_rgba_and_occupancy_grid():
    """Whole-map ingest on the device.  The RGBA conversion is checked exhaustively -- a 4096 x 4096
    image holds every (r, g, b) triple once -- against the host mirror of the reference's OMap(png, threshold)
    (mapio.occupancy_from_rgba, itself pinned to the reference's loader in the CPU suite); structures rebuilt from
    the ingested map equal those of a handle constructed from the same occupancy."""
    import torch
    from range_libc_b200 import mapio
    side = 4096
    v = np.arange(side * side, dtype=np.uint32)
    rgba = np.empty((side, side, 4), np.uint8)
    rgba[..., 0] = (v & 255).reshape(side, side)
    rgba[..., 1] = ((v >> 8) & 255).reshape(side, side)
    rgba[..., 2] = ((v >> 16) & 255).reshape(side, side)
    rgba[..., 3] = 255
    for thr in (128.0, 37.5):
        want = mapio.occupancy_from_rgba(rgba, thr)  # [W, H] x-major
        bl = rl.PyBresenhamsLine(rl.PyOMap(side, side), MR)
        bl.set_map_rgba(torch.from_numpy(rgba).cuda(), thr)
        assert np.array_equal(bl.occupancy(), want), "rgba ingest, threshold %g" % thr
    # non-square, non-multiple-of-32 image, host pointer; RM rebuilds its distance transform
    rng = np.random.default_rng(8)
    rows, cols = 301, 517  # image rows = map height, cols = map width
    img = rng.integers(0, 256, (rows, cols, 4), dtype=np.uint8)
    img[rng.random((rows, cols)) < 0.9] = 255
    occ = mapio.occupancy_from_rgba(img, 128.0)
    assert occ.shape == (cols, rows)
    rm = rl.PyRayMarchingGPU(rl.PyOMap(cols, rows), MR)
    rm.set_map_rgba(img)
    fresh = make("rm", occ)
    assert_bit_equal(rm.distance_transform(), fresh.distance_transform(), "distance transform after rgba ingest")
    q = wl.random_queries(cols, rows, 20000, seed=3)
    a, b = np.empty(len(q), np.float32), np.empty(len(q), np.float32)
    rm.calc_range_many_grid(q, a)
    fresh.calc_range_many_grid(q, b)
    assert_bit_equal(a, b, "ranges after rgba ingest")
    # OccupancyGrid data: int8 [rows = map width][cols = map height], occupied iff > 10
    data = rng.choice(np.array([-1, 0, 5, 10, 11, 100], np.int8), size=(cols, rows), p=[.2, .6, .05, .05, .05, .05])
    occ2 = (data > 10).astype(np.uint8)
    for src in (data, torch.from_numpy(data).cuda()):
        cd = rl.PyCDDTCast(rl.PyOMap(cols, rows), MR, TD)
        cd.prune()
        cd.set_map_occupancy_grid(src)
        fresh = make("pcddt", occ2)
        cd.calc_range_many_grid(q, a)
        fresh.calc_range_many_grid(q, b)
        assert_bit_equal(a, b, "PCDDT ranges after occupancy-grid ingest")
    with pytest.raises(rl.RangeLibError):
        cd.set_map_occupancy_grid(data[:-1])


@pytest.mark.parametrize("kn", ["rm", "bl", "pcddt"])
def test_fused_small_host_call_paths(kn):
    """The small host-pointer fused call (poses by one async copy straight from pinned caller memory, angles and
    observation inside the launch, weights stored into pinned host memory) against the oracle, for pinned and
    pageable caller arrays, beam counts on both sides of the in-launch limit (256), and a sub-array view."""
    import torch
    occ = wl.load_map("basement_hallways_10cm")
    world = (0.1, 0.0, -5.0, 3.0, 0.0, 1.0)
    meth = make(kn, occ, world=world)
    table = wl.sensor_table(501)
    meth.set_sensor_model(table)
    o = port.Oracle(KINDS[kn], occ, MR, TD, threads=8)
    o.set_world(*world)
    o.set_sensor_model(table)
    rng = np.random.default_rng(4)
    for n, m_beams in ((1, 1), (777, 60), (4000, 60), (50, 256), (50, 257), (9, 1080)):
        parts = wl.grid_to_world(wl.pf_particles_uniform(occ, n, seed=n), world[0], world[2], world[3])
        angles = wl.lidar_angles(m_beams) if m_beams > 1 else np.zeros(1, np.float32)
        obs = rng.uniform(0, 50.0, m_beams).astype(np.float32)
        want = o.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs)
        for pin in (False, True):
            mk = (lambda a: torch.from_numpy(a.copy()).pin_memory().numpy()) if pin else (lambda a: a.copy())
            hp, ha, ho = mk(parts), mk(angles), mk(obs)
            hw = mk(np.full(n + 3, -1.0))
            meth.calc_range_repeat_angles_eval_sensor_model(hp, ha, ho, hw[1:n + 1])  # view with an offset
            assert_bit_equal(hw[1:n + 1], want, "%s n=%d m=%d pinned=%s" % (kn, n, m_beams, pin))
            assert hw[0] == -1.0 and (hw[n + 1:] == -1.0).all()


def test_large_cloud_spatial_ordering_is_invisible():
    """Big clouds on a map whose distance transform exceeds the L2 threshold are processed tile by tile
    (rl_sort.cu); weights land at their own indices and equal the oracle's bit for bit."""
    import torch
    occ = wl.synthetic_map(4096, seed=5)  # 64 MB distance transform
    meth = make("rm", occ)
    table = wl.sensor_table(501)
    meth.set_sensor_model(table)
    n, m_beams = 40000, 6
    parts = wl.pf_particles_uniform(occ, n, seed=12)
    parts[7] = [np.nan, 1.0, 0.0]        # non-finite and out-of-map poses go through the key kernel too
    parts[8] = [-1e9, 5e9, 1.0]
    angles = wl.lidar_angles(m_beams)
    obs = np.linspace(20.0, 400.0, m_beams).astype(np.float32)
    o = port.Oracle(port.RM, occ, MR, threads=8)
    o.set_sensor_model(table)
    want = o.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs)
    got = torch.empty(n, dtype=torch.float64, device="cuda")
    l0 = rl.kernel_launches()
    meth.calc_range_repeat_angles_eval_sensor_model(torch.from_numpy(parts).cuda(), torch.from_numpy(angles).cuda(),
                                                    torch.from_numpy(obs).cuda(), got)
    meth.synchronize()
    assert rl.kernel_launches() - l0 >= 3, "the ordering pass did not run"
    assert_bit_equal(got.cpu().numpy(), want, "spatially ordered fused update")


@pytest.mark.parametrize("n,m_beams", [(12000, 60), (700, 1080), (300, 2500), (600011, 1)])
def test_deep_fused_launches_use_requeuing_kernel(n, m_beams):
    """Fused RM updates many waves deep run on fused_rm_persist_kernel (lane re-queuing inside a CTA's particle
    group, product in beam order): weights bit-equal to the oracle, rotated world frame, odd sizes."""
    import torch
    occ = wl.load_map("basement_hallways_5cm")
    world = (0.05, 0.3, -3.0, 2.0, float(np.float32(np.sin(0.3))), float(np.float32(np.cos(0.3))))
    meth = make("rm", occ, world=world)
    table = wl.sensor_table(501)
    meth.set_sensor_model(table)
    o = port.Oracle(port.RM, occ, MR, threads=8)
    o.set_world(*world)
    o.set_sensor_model(table)
    parts = wl.grid_to_world(wl.pf_particles_uniform(occ, n, seed=n), world[0], world[2], world[3], world[1])
    parts[n // 2] = [np.inf, 0.0, 0.0]
    angles = wl.lidar_angles(m_beams) if m_beams > 1 else np.array([0.1], np.float32)
    obs = np.random.default_rng(9).uniform(0, 25.0, m_beams).astype(np.float32)
    want = o.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs)
    got = np.empty(n, np.float64)
    meth.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs, got)
    assert_bit_equal(got, want, "deep fused %dx%d (host)" % (n, m_beams))
    d_got = torch.empty(n, dtype=torch.float64, device="cuda")
    meth.calc_range_repeat_angles_eval_sensor_model(torch.from_numpy(parts).cuda(), torch.from_numpy(angles).cuda(),
                                                    torch.from_numpy(obs).cuda(), d_got)
    meth.synchronize()
    assert_bit_equal(d_got.cpu().numpy(), want, "deep fused %dx%d (device)" % (n, m_beams))


def test_fuzz_small_random_maps_all_kinds():
    """Random small non-square maps x {BL, RM, CDDT, PCDDT} x odd theta discretizations / max ranges / world
    frames: grid, world, angle-fan and fused entry points against the oracle (itself fuzzed against the
    unmodified reference in tests/test_oracle_vs_ref.py), queries inside and outside the map."""
    rng = np.random.default_rng(20261018)
    for it in range(24):
        W, H = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        occ = (rng.random((W, H)) < rng.choice([0.0, 0.02, 0.1, 0.3])).astype(np.uint8)
        if rng.random() < 0.5:
            occ[int(rng.integers(0, W)), :] = 1
        mr = float(rng.choice([3.5, 20.0, 50.0, 500.0]))
        td = int(rng.choice([4, 7, 16, 108, 361]))
        ang = float(rng.uniform(-3, 3))
        world = (float(rng.choice([1.0, 0.05, 2.5])), ang, float(rng.uniform(-5, 5)), float(rng.uniform(-5, 5)),
                 float(np.float32(np.sin(ang))), float(np.float32(np.cos(ang))))
        q = wl.random_queries(max(W, 2), max(H, 2), 300, seed=it)
        q[:40, :2] += rng.uniform(-8, 8, (40, 2)).astype(np.float32)  # some start outside the map
        qw = wl.grid_to_world(q, world[0], world[2], world[3], world[1])
        parts = qw[:50]
        angles = wl.lidar_angles(int(rng.integers(1, 40)) + 1)
        obs = rng.uniform(0, mr * world[0], len(angles)).astype(np.float32)
        K = int(rng.integers(2, 60))
        table = rng.uniform(0.05, 1.0, (K, K))
        for kn in ("bl", "rm", "cddt", "pcddt"):
            what = "iter %d %s map %dx%d mr %g td %d" % (it, kn, W, H, mr, td)
            meth = make(kn, occ, max_range=mr, td=td, world=world)
            o = port.Oracle(KINDS[kn], occ, mr, td)
            o.set_world(*world)
            out = np.empty(len(q), np.float32)
            meth.calc_range_many_grid(q, out)
            assert_bit_equal(out, o.calc_range_many(q), what + " grid")
            if kn in ("cddt", "pcddt") and it % 2:  # odd iterations: everything below through the CDDT query index
                meth.set_spatial_sort(2)
                meth.calc_range_many_grid(q, out)
                assert_bit_equal(out, o.calc_range_many(q), what + " grid, query index")
            meth.calc_range_many(qw, out)
            assert_bit_equal(out, o.numpy_calc_range(qw), what + " world")
            fan = np.empty(len(parts) * len(angles), np.float32)
            meth.calc_range_repeat_angles(parts, angles, fan)
            assert_bit_equal(fan, o.numpy_calc_range_angles(parts, angles), what + " angles")
            meth.set_sensor_model(table)
            o.set_sensor_model(table)
            w = np.empty(len(parts), np.float64)
            meth.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs, w)
            assert_bit_equal(w, o.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs), what + " fused")
            meth.eval_sensor_model(obs, fan, w, len(angles), len(parts))
            assert_bit_equal(w, o.eval_sensor_model(obs, fan, len(angles), len(parts)), what + " two-step")


def test_peer_store_epilogue_single_gpu_all_launch_shapes():
    """The multi-GPU epilogue (weights stored through peer pointers at an offset into a gathered array) with this
    GPU as its own only peer, for each fused launch shape: small (fused_kernel + cooperative tail), deep with many
    particles per group (fused_rm_persist_kernel), deep with ~1000-beam particles (fused_kernel), and a big cloud on
    a map beyond the L2 threshold (processing order permuted by rl_sort.cu)."""
    import torch
    cases = [("basement_hallways_5cm", 900, 60), ("basement_hallways_5cm", 12000, 60),
             ("basement_hallways_5cm", 700, 1080), (4096, 40000, 16)]
    for name, n, m_beams in cases:
        occ = wl.synthetic_map(name, seed=5) if isinstance(name, int) else wl.load_map(name)
        meth = make("rm", occ)
        table = wl.sensor_table(501)
        meth.set_sensor_model(table)
        parts = wl.pf_particles_uniform(occ, n, seed=3)
        angles = wl.lidar_angles(m_beams)
        obs = np.linspace(10.0, 300.0, m_beams).astype(np.float32)
        o = port.Oracle(port.RM, occ, MR, threads=8)
        o.set_sensor_model(table)
        want = o.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs)
        offset = 17
        gathered = torch.full((n + 40,), -1.0, dtype=torch.float64, device="cuda")
        meth.calc_range_repeat_angles_eval_sensor_model_peers(
            torch.from_numpy(parts).cuda(), torch.from_numpy(angles).cuda(), torch.from_numpy(obs).cuda(),
            [gathered.data_ptr()], offset)
        meth.synchronize()
        got = gathered.cpu().numpy()
        assert_bit_equal(got[offset:offset + n], want, "peer epilogue %s %dx%d" % (name, n, m_beams))
        assert (got[:offset] == -1.0).all() and (got[offset + n:] == -1.0).all()
