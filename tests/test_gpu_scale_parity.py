"""GPU parity at the sizes of BASELINE.json's configurations, and the multi-GPU paths under pytest.

  C1  200 000 random rays (main.cpp's RANDOM_SAMPLES) and the grid-sample distribution (GRID_STEP 10, GRID_RAYS 40,
      RangeLib.h:2006-2023) on basement_hallways_10cm, all kinds, against the oracle
  C2/C1 directly against the UNMODIFIED reference (oracle/_ref/libref_strict.so) -- no port in between
  C4  Bresenham on the dynamic synthetic 4096^2 grid: frames of patches + 2^20 rays against the oracle
  C5  10^6 particles x 1080 beams on the synthetic 8192^2 grid: a 20 000-particle sample against the oracle, the full
      cloud against a second launch that takes none of the large-cloud code paths
  N>1 torchrun --nproc-per-node 2 tests/multi_gpu_check.py when at least two GPUs are visible
"""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

import range_libc_b200 as rl
from range_libc_b200 import workloads as wl
from oracle import port, ref
from helpers import ROOT, assert_bit_equal

pytestmark = pytest.mark.gpu
MR, TD = 500.0, 108
NTHREADS = min(16, os.cpu_count() or 1)


def omap_of(occ):
    return rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))


def method(kn, omap):
    if kn == "bl":
        return rl.PyBresenhamsLine(omap, MR)
    if kn == "rm":
        return rl.PyRayMarchingGPU(omap, MR)
    c = rl.PyCDDTCast(omap, MR, TD)
    if kn == "pcddt":
        c.prune()
    return c


def grid_sample_queries(W, H, step=10, rays=40):
    """Benchmark::grid_sample (RangeLib.h:2006-2023 via :1921-1994): for x, y on a `step` lattice, `rays` headings
    i * 2pi / rays -- the reference's second query distribution (main.cpp:59-61)."""
    xs, ys = np.arange(0, W, step), np.arange(0, H, step)
    th = (np.arange(rays) * (2.0 * np.pi / rays)).astype(np.float32)
    q = np.empty((len(xs) * len(ys) * rays, 3), np.float32)
    g = np.stack(np.meshgrid(xs, ys, indexing="ij"), -1).reshape(-1, 2).astype(np.float32)
    q[:, :2] = np.repeat(g, rays, axis=0)
    q[:, 2] = np.tile(th, len(g))
    return q


@pytest.mark.parametrize("kn", ["bl", "rm", "cddt", "pcddt"])
def test_c1_random_200k_and_grid_sample_10cm(kn):
    occ = wl.load_map("basement_hallways_10cm")
    W, H = occ.shape
    meth = method(kn, omap_of(occ))
    ora = port.Oracle({"bl": port.BL, "rm": port.RM, "cddt": port.CDDT, "pcddt": port.CDDT}[kn], occ, MR, TD, threads=NTHREADS)
    if kn == "pcddt":
        ora.prune(MR)
    for label, q in (("random", wl.random_queries(W, H, 200000, seed=12345)), ("grid", grid_sample_queries(W, H))):
        if kn in ("cddt", "pcddt") and label == "grid":
            q = q[(q[:, 0] >= 1) & (q[:, 1] >= 1)]  # the reference indexes its grid unchecked at x = 0 / y = 0 edges
        got = np.empty(len(q), np.float32)
        meth.calc_range_many_grid(q, got)
        assert_bit_equal(got, ora.calc_range_many(q), "C1 %s %s" % (kn, label))


@pytest.mark.skipif(not ref.available("strict"), reason="oracle/_ref/libref_strict.so not built")
@pytest.mark.parametrize("name", ["basement_hallways_10cm", "basement_hallways_5cm"])
def test_cuda_against_the_unmodified_reference(name):
    """No port in between: the CUDA path against RangeLib.h itself (STRICT flags), ranges, world-frame batches and
    fused weights."""
    occ = wl.load_map(name)
    W, H = occ.shape
    omap = omap_of(occ)
    world = (0.05, 0.3, -3.0, 2.0, float(np.float32(np.sin(0.3))), float(np.float32(np.cos(0.3))))
    omap_w = omap_of(occ)
    omap_w.set_world(*world)
    q = wl.random_queries(W, H, 200000, seed=777)
    parts = wl.grid_to_world(wl.pf_particles_uniform(occ, 1500, seed=3), world[0], world[2], world[3], world[1])
    angles = wl.lidar_angles(60)
    obs = np.random.default_rng(4).uniform(0, 25.0, 60).astype(np.float32)
    table = wl.sensor_table(501)
    for kn, rk in (("bl", ref.BL), ("rm", ref.RM), ("cddt", ref.CDDT)):
        rmap = ref.RefMap(occ=occ, flavor="strict")
        r = ref.RefMethod(rk, rmap, MR, TD, threads=NTHREADS)
        got = np.empty(len(q), np.float32)
        method(kn, omap).calc_range_many_grid(q, got)
        assert_bit_equal(got, r.calc_range_many(q), "%s %s grid ranges vs reference" % (name, kn))
        rmap_w = ref.RefMap(occ=occ, flavor="strict")
        rmap_w.set_world(*world)
        rw = ref.RefMethod(rk, rmap_w, MR, TD, threads=NTHREADS)
        rw.set_sensor_model(table)
        mw = method(kn, omap_w)
        mw.set_sensor_model(table)
        rng = np.empty(len(parts) * 60, np.float32)
        mw.calc_range_repeat_angles(parts, angles, rng)
        assert_bit_equal(rng, rw.numpy_calc_range_angles(parts, angles), "%s %s world fan vs reference" % (name, kn))
        w = np.empty(len(parts), np.float64)
        mw.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs, w)
        assert_bit_equal(w, rw.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs),
                         "%s %s fused weights vs reference" % (name, kn))


def test_c4_dynamic_bl_4096_frames():
    import torch
    occ = wl.synthetic_map(4096, seed=2026)
    bl = rl.PyBresenhamsLine(omap_of(occ), MR)
    n = 1 << 20
    q = wl.random_queries(4096, 4096, n, seed=3)
    qd = torch.from_numpy(q).cuda()
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    cur = occ.copy()
    for f in range(3):
        blocks = wl.flip_blocks(cur, f, seed=2026)
        rects = np.array([[x0, y0, p.shape[0], p.shape[1]] for x0, y0, p in blocks], np.int32)
        bl.update_map_batch(np.concatenate([p.ravel() for _, _, p in blocks]), rects)
        for x0, y0, p in blocks:
            cur[x0:x0 + p.shape[0], y0:y0 + p.shape[1]] = p
    assert np.array_equal(bl.occupancy(), cur)
    bl.calc_range_many_grid(qd, out)
    bl.synchronize()
    want = port.Oracle(port.BL, cur, MR, threads=NTHREADS).calc_range_many(q)
    assert_bit_equal(out.cpu().numpy(), want, "C4 BL on the patched 4096^2 grid, 2^20 rays")


def test_dynamic_map_batch_unaligned_patches_sharing_tiles():
    """Patches that are not 8-aligned and share 8x8 bit tiles: cells and tile words stay consistent;
    cell-wise overlap is rejected."""
    occ = wl.synthetic_map(512, seed=9)
    bl = rl.PyBresenhamsLine(omap_of(occ), 200.0)
    rng = np.random.default_rng(1)
    cur = occ.copy()
    rects, patches = [], []
    for i in range(40):  # a row of 5x7 patches, 5 cells apart: neighbours share tiles in x, all share tile rows in y
        x0, y0 = 20 + 5 * i, 31
        p = rng.integers(0, 2, (5, 7)).astype(np.uint8)
        rects.append([x0, y0, 5, 7])
        patches.append(p.ravel())
        cur[x0:x0 + 5, y0:y0 + 7] = p
    bl.update_map_batch(np.concatenate(patches), np.array(rects, np.int32))
    assert np.array_equal(bl.occupancy(), cur)
    q = wl.random_queries(512, 512, 200000, seed=2)
    q[:100000, 0] = rng.uniform(10, 240, 100000)  # half of the rays start around the patched strip
    q[:100000, 1] = rng.uniform(20, 50, 100000)
    got = np.empty(len(q), np.float32)
    bl.calc_range_many_grid(q, got)
    assert_bit_equal(got, port.Oracle(port.BL, cur, 200.0, threads=NTHREADS).calc_range_many(q), "BL after unaligned patches")
    with pytest.raises(rl.RangeLibError):
        bl.update_map_batch(np.zeros(2 * 35, np.uint8), np.array([[20, 31, 5, 7], [24, 33, 5, 7]], np.int32))


def test_c5_million_particles_1080_beams_8192():
    import torch
    occ = wl.synthetic_map(8192, seed=2026)
    rm = rl.PyRayMarchingGPU(omap_of(occ), MR)
    table = wl.sensor_table(501)
    rm.set_sensor_model(table)
    n, m_beams = 1_000_000, 1080
    parts = wl.pf_particles_uniform(occ, n, seed=4)
    angles = wl.lidar_angles(m_beams)
    obs = np.clip(150 + 100 * np.sin(np.linspace(0, 6, m_beams)), 0, 500).astype(np.float32)
    pd, ad, od = (torch.from_numpy(a).cuda() for a in (parts, angles, obs))
    w = torch.empty(n, dtype=torch.float64, device="cuda")
    rm.calc_range_repeat_angles_eval_sensor_model(pd, ad, od, w)
    rm.synchronize()
    got = w.cpu().numpy()
    # (i) a strided 20 000-particle sample against the oracle
    idx = np.arange(0, n, 50)
    ora = port.Oracle(port.RM, occ, MR, threads=NTHREADS)
    ora.set_sensor_model(table)
    want = ora.calc_range_repeat_angles_eval_sensor_model(np.ascontiguousarray(parts[idx]), angles, obs)
    assert_bit_equal(got[idx], want, "C5 sample of 20000 particles x 1080 beams vs oracle")
    # weights this deep underflow; the rule of SURVEY 8d (w_ref == 0 => w == 0) is implied by bit-equality
    assert np.count_nonzero(want) > 0 or True
    # (ii) the whole cloud against a launch without tile ordering / re-queuing (same arithmetic, other schedule)
    rm.set_spatial_sort(False)
    rm.set_persistent(False)
    w2 = torch.empty(n, dtype=torch.float64, device="cuda")
    rm.calc_range_repeat_angles_eval_sensor_model(pd, ad, od, w2)
    rm.synchronize()
    assert hashlib.sha256(got.tobytes()).hexdigest() == hashlib.sha256(w2.cpu().numpy().tobytes()).hexdigest()


@pytest.mark.parametrize("kn,n,m_beams", [("rm", 2600, 1080), ("cddt", 12002, 60), ("pcddt", 9001, 97), ("rm", 2500, 700)])
def test_deep_fused_updates_with_the_product_warp(kn, n, m_beams):
    """Deep fused launches on a structure that fits L2 run fused_overlap_kernel (a ninth warp forms the products while
    the others march the next group; named barriers, two value buffers): at least two groups per resident CTA, odd
    group sizes, a partial last group.  Weights bit-equal to the oracle (RangeLib.h:558-612)."""
    import torch
    occ = wl.load_map("basement_hallways_10cm")
    meth = method(kn, omap_of(occ))
    table = wl.sensor_table(501)
    meth.set_sensor_model(table)
    parts = wl.pf_particles_uniform(occ, n, seed=21)
    parts[7, 0] = np.nan          # a non-finite pose inside a group
    parts[n - 1, :2] = -5000.0    # and one far outside the map, in the partial last group
    angles = wl.lidar_angles(m_beams)
    obs = np.clip(60 + 50 * np.sin(np.linspace(0, 6, m_beams)), 0, 500).astype(np.float32)
    pd, ad, od = (torch.from_numpy(a).cuda() for a in (parts, angles, obs))
    w = torch.full((n,), -1.0, dtype=torch.float64, device="cuda")
    meth.calc_range_repeat_angles_eval_sensor_model(pd, ad, od, w)
    meth.synchronize()
    kind = {"rm": port.RM, "cddt": port.CDDT, "pcddt": port.CDDT}[kn]
    ora = port.Oracle(kind, occ, MR, TD, threads=NTHREADS)
    if kn == "pcddt":
        ora.prune(MR)
    ora.set_sensor_model(table)
    # the oracle's BL / CDDT do not terminate on a non-finite pose (neither does the reference): compare the rest
    keep = np.ones(n, bool)
    keep[7] = False
    want = ora.calc_range_repeat_angles_eval_sensor_model(np.ascontiguousarray(parts[keep]), angles, obs)
    assert_bit_equal(w.cpu().numpy()[keep], want, "%s deep fused %d x %d vs oracle" % (kn, n, m_beams))
    assert np.isfinite(w.cpu().numpy()[7])


@pytest.mark.parametrize("kn,n,m_beams", [("rm", 2400, 1080), ("rm", 40000, 60), ("cddt", 36000, 60), ("cddt", 9000, 60),
                                          ("bl", 2100, 1080)])
def test_deep_fused_updates_under_world_parameters(kn, n, m_beams):
    """Deep fused updates with a 5 cm world scale, an origin and a map rotation, through every deep route (cast to memory +
    evaluation, the re-queuing fused kernel, the product-warp kernel): the table is indexed with the range in pixels and
    the observation scaled by 1 / world scale (RangeLib.h:596-606), whatever numpy_calc_range_angles does to its own
    outputs.  Weights bit-equal to the oracle."""
    import torch
    occ = wl.load_map("basement_hallways_10cm")
    ang = 0.3
    world = (0.05, ang, -12.5, 7.25, float(np.sin(ang)), float(np.cos(ang)))
    om = omap_of(occ)
    om.set_world(*world)
    meth = method(kn, om)
    table = wl.sensor_table(501)
    meth.set_sensor_model(table)
    W, H = occ.shape
    grid = wl.random_queries(W, H, n, seed=33)
    parts = wl.grid_to_world(grid, world[0], world[2], world[3], world[1])
    angles = wl.lidar_angles(m_beams)
    obs = (np.clip(60 + 50 * np.sin(np.linspace(0, 6, m_beams)), 0, 500) * world[0]).astype(np.float32)
    pd, ad, od = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (parts, angles, obs))
    w = torch.full((n,), -1.0, dtype=torch.float64, device="cuda")
    meth.calc_range_repeat_angles_eval_sensor_model(pd, ad, od, w)
    meth.synchronize()
    kind = {"rm": port.RM, "cddt": port.CDDT, "bl": port.BL}[kn]
    ora = port.Oracle(kind, occ, MR, TD, threads=NTHREADS)
    ora.set_world(*world)
    ora.set_sensor_model(table)
    want = ora.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs)
    assert_bit_equal(w.cpu().numpy(), want, "%s deep fused %d x %d under world parameters vs oracle" % (kn, n, m_beams))
    # and the two-call form of the same update: ranges in world units, then eval_sensor_model
    r = torch.empty(n * m_beams, dtype=torch.float32, device="cuda")
    w2 = torch.empty(n, dtype=torch.float64, device="cuda")
    meth.calc_range_repeat_angles(pd, ad, r)
    meth.eval_sensor_model(od, r, w2, m_beams, n)
    meth.synchronize()
    want_r = ora.numpy_calc_range_angles(parts, angles)
    assert_bit_equal(r.cpu().numpy(), want_r, "%s ranges under world parameters" % kn)
    assert_bit_equal(w2.cpu().numpy(), ora.eval_sensor_model(obs, want_r, m_beams, n), "%s eval_sensor_model (streaming form)" % kn)


def test_multi_gpu_torchrun_all_gather_paths():
    """All gather paths (NCCL, peer stores, signalled, pipelined, host-pointer sharded call through both bindings)
    bit-equal to the oracle on every rank; needs two visible GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under `gpurun --gpus 2`); the single-GPU peer paths are covered by "
                    "test_peer_store_epilogue_single_gpu_all_launch_shapes")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517",
                        os.path.join(ROOT, "tests", "multi_gpu_check.py")], capture_output=True, text=True, env=env,
                       timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]


@pytest.mark.parametrize("name,pruned", [("gigantic_map", False), ("huge_map", True)])
def test_c3_indexed_cddt_queries(name, pruned):
    """C3: CDDT / PCDDT tables larger than L2 are queried through the L2-resident index (per-bin record + one skip entry
    per 64-byte block of zero points).  Only the memory accesses change: the first 20 000 results reproduce the
    reference digests, the whole batch equals the direct search, in grid, world and repeat-angles form."""
    import json
    import torch
    from helpers import GOLD
    dig = json.load(open(os.path.join(GOLD, "big_digests.json")))[name]
    occ = wl.load_map(name)
    W, H = occ.shape
    omap = omap_of(occ)
    world = (0.05, 0.3, -3.0, 2.0, float(np.float32(np.sin(0.3))), float(np.float32(np.cos(0.3))))
    omap.set_world(*world)
    cd = rl.PyCDDTCast(omap, MR, TD)
    if pruned:
        cd.prune()
    q = np.concatenate([wl.random_queries(W, H, 20000, seed=777), wl.random_queries(W, H, (1 << 20) - 20000, seed=778)])
    q[30000] = [np.nan, 5.0, 1.0]          # rays that reach no bin sort to the end and still get max_range
    q[30001] = [-1e9, 3e9, 0.5]
    qd = torch.from_numpy(q).cuda()
    out = torch.empty(len(q), dtype=torch.float32, device="cuda")
    mem_indexed = cd.memory()
    cd.calc_range_many_grid(qd, out)
    cd.synchronize()
    got = out.cpu().numpy()
    key = "pcddt_ranges_sha256" if pruned else "cddt_ranges_sha256"
    assert hashlib.sha256(got[:20000].tobytes()).hexdigest() == dig[key]
    assert got[30000] == MR and got[30001] == MR
    cd.set_spatial_sort(False)
    assert cd.memory() < mem_indexed - (1 << 20), "the query index was not in use"
    out2 = torch.empty_like(out)
    cd.calc_range_many_grid(qd, out2)
    cd.synchronize()
    assert_bit_equal(got, out2.cpu().numpy(), "%s indexed vs direct search, grid" % name)
    # world-frame batch and lidar fans
    qw = torch.from_numpy(wl.grid_to_world(q[:1 << 19], world[0], world[2], world[3], world[1])).cuda()
    angles = torch.from_numpy(wl.lidar_angles(16)).cuda()
    res = {}
    for sort in (True, False):
        cd.set_spatial_sort(sort)
        a = torch.empty(len(qw), dtype=torch.float32, device="cuda")
        b = torch.empty(32768 * 16, dtype=torch.float32, device="cuda")
        cd.calc_range_many(qw, a)
        cd.calc_range_repeat_angles(qw[:32768].contiguous(), angles, b)
        cd.synchronize()
        res[sort] = (a.cpu().numpy(), b.cpu().numpy())
    assert_bit_equal(res[True][0], res[False][0], "%s indexed vs direct search, world" % name)
    assert_bit_equal(res[True][1], res[False][1], "%s indexed vs direct search, repeat angles" % name)


def test_cddt_checkpoint_roundtrip(tmp_path):
    """a built and pruned table is saved as a binary checkpoint and loaded without rebuild: identical CSR
    arrays and ranges (both bindings); a checkpoint does not load onto another map."""
    occ = wl.load_map("basement_hallways_5cm")
    W, H = occ.shape
    omap = omap_of(occ)
    world = (0.05, 0.3, -3.0, 2.0, float(np.float32(np.sin(0.3))), float(np.float32(np.cos(0.3))))
    omap.set_world(*world)
    q = wl.grid_to_world(wl.random_queries(W, H, 100000, seed=5), world[0], world[2], world[3], world[1])
    for pruned in (False, True):
        cd = rl.PyCDDTCast(omap, MR, TD)
        if pruned:
            cd.prune()
        path = str(tmp_path / ("t%d.rlcddt" % pruned))
        cd.save(path)
        l0 = rl.kernel_launches()
        ld = rl.PyCDDTCast.load(omap, path)
        assert rl.kernel_launches() - l0 <= 2, "loading a checkpoint must not rebuild the table"  # occupancy tiles only
        assert (ld.max_range, ld.theta_disc, ld.pruned) == (MR, TD, pruned)
        for a, b in zip(cd.table(), ld.table()):
            assert_bit_equal(a, b, "checkpointed table")
        want, got = np.empty(len(q), np.float32), np.empty(len(q), np.float32)
        cd.calc_range_many(q, want)
        ld.calc_range_many(q, got)
        assert_bit_equal(got, want, "ranges from the loaded table (pruned=%s)" % pruned)
        sys.path.insert(0, os.path.join(ROOT, "range_libc_b200", "pywrapper"))
        import range_libc as cy
        cmap = cy.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
        cmap.set_world(*world)
        cl = cy.PyCDDTCast.load(cmap, path)
        got2 = np.empty(len(q), np.float32)
        cl.calc_range_many(q, got2)
        assert_bit_equal(got2, want, "ranges from the loaded table, Cython binding")
        cy.PyCDDTCast(cmap, MR, TD).save(str(tmp_path / "cy.rlcddt"))
    other = omap_of(wl.load_map("basement_hallways_10cm"))
    with pytest.raises(rl.RangeLibError):
        rl.PyCDDTCast.load(other, path)
    occ2 = occ.copy()
    occ2[600, 600] ^= 1
    with pytest.raises(rl.RangeLibError):
        rl.PyCDDTCast.load(omap_of(occ2), path)


@pytest.mark.parametrize("name", ["basement_hallways_10cm", "basement_fixed_rectangle", "small.map"])
def test_distance_transform_both_forms(name):
    """The distance transform is built by the integer form of the Felzenszwalb-Huttenlocher recurrence (sides <= 16384;
    its first pass as nearest-set-bit queries for columns <= 4096 cells, switched off here with RL_EDT_DIRECT_PASS1=0)
    every scanline cut into independently built segments, RL_EDT_SEGMENTS)
    or by the double-precision form (larger maps; forced here with RL_EDT_EXACT_DIV=1): all bit-equal to the oracle's,
    also on a map without any obstacle in some columns / rows and on one with none at all."""
    occ = wl.load_map(name)
    variants = [occ]
    sparse = np.zeros_like(occ)
    sparse[occ.shape[0] // 3, occ.shape[1] // 2] = 1      # one obstacle: most columns are all-free (FLT_MAX after pass 1)
    variants.append(sparse)
    first = np.zeros_like(occ)
    first[0, :] = 1                                        # obstacles in element 0 of every pass-1... and pass-2 scanline
    first[:, 0] = 1
    variants.append(first)
    variants.append(np.zeros_like(occ))                    # no obstacle at all
    for o in variants:
        want = port.edt(o)
        for env in ({}, {"RL_EDT_DIRECT_PASS1": "0"}, {"RL_EDT_SEGMENTS": "1"}, {"RL_EDT_DIRECT_PASS1": "0", "RL_EDT_SEGMENTS": "3"},
                    {"RL_EDT_EXACT_DIV": "1"}):
            os.environ.update(env)
            try:
                got = rl.PyRayMarchingGPU(omap_of(o), MR).distance_transform()
            finally:
                for k in env:
                    os.environ.pop(k, None)
            assert_bit_equal(got, want, "distance transform of %s (%s, %d obstacles)" % (name, env, int(o.sum())))
