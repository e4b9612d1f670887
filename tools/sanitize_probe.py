"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import range_libc_b200 as rl  # noqa: E402
from range_libc_b200 import workloads as wl  # noqa: E402

occ = wl.load_map("basement_hallways_10cm")[150:406, 180:420].copy()
W, H = occ.shape
omap = rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
q = wl.random_queries(W, H, 3000, seed=1)
parts = wl.pf_particles_uniform(occ, 40, seed=2)
angles = wl.lidar_angles(60)
obs = np.linspace(3, 200, 60).astype(np.float32)
table = wl.sensor_table(257)
out = np.empty(len(q), np.float32)
w = np.empty(len(parts), np.float64)
rng = np.empty(len(parts) * len(angles), np.float32)
for ctor in (lambda: rl.PyRayMarchingGPU(omap, 256.0), lambda: rl.PyBresenhamsLine(omap, 256.0),
             lambda: rl.PyCDDTCast(omap, 256.0, 24), lambda: rl.PyGiantLUTCast(omap, 256.0, 8)):
    m = ctor()
    if isinstance(m, rl.PyCDDTCast):
        m.prune()
    m.set_sensor_model(table)
    m.calc_range_many_grid(q, out)
    m.calc_range_many(q, out)
    m.calc_range_repeat_angles(parts, angles, rng)
    m.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs, w)
    m.eval_sensor_model(obs, rng, w, len(angles), len(parts))
    m.update_map_batch(np.ones(64, np.uint8), np.array([[8, 8, 8, 8]], np.int32))
    m.calc_range_many_grid(q, out)
    print(type(m).__name__, "ok", float(out.mean()))

# paths added later in round 1 (not yet run under the sanitizer: the round's GPU budget was spent)
import torch  # noqa: E402

rm = rl.PyRayMarchingGPU(omap, 256.0)
rm.set_sensor_model(table)
cd = rl.PyCDDTCast(omap, 256.0, 24)
fan = np.full(len(parts) * 37, -1.0, np.float32)
for m in (rm, cd):  # radial_kernel / cddt_cast_pair
    m.calc_range_many_radial_optimized(37, -1.0, 2.0, parts, fan)
# deep fused update: fused_rm_persist_kernel (lane re-queuing inside particle groups)
big = wl.pf_particles_uniform(occ, 12000, seed=3)
wb = np.empty(len(big), np.float64)
rm.calc_range_repeat_angles_eval_sensor_model(big, angles, obs, wb)
# large independent batch on device pointers: rm_persist_kernel (parked rays in registers)
qd = torch.from_numpy(wl.random_queries(W, H, 600000, seed=4)).cuda()
od = torch.empty(len(qd), dtype=torch.float32, device="cuda")
rm.calc_range_many_grid(qd, od)
rm.synchronize()
# whole-map ingest kernels
img = np.random.default_rng(5).integers(0, 256, (H, W, 4), dtype=np.uint8)
rm.set_map_rgba(img, 128.0)
cd.set_map_occupancy_grid(np.random.default_rng(6).integers(-1, 101, (W, H)).astype(np.int8))
# spatial ordering (rl_sort.cu): needs a structure beyond the L2 threshold, i.e. a map of >= 3548^2 cells
occ_big = wl.synthetic_map(3600, seed=7)
rmb = rl.PyRayMarchingGPU(rl.PyOMap(np.ascontiguousarray(occ_big.T.astype(bool))), 256.0)
rmb.set_sensor_model(table)
cloud = wl.pf_particles_uniform(occ_big, 33000, seed=8)
wc = np.empty(len(cloud), np.float64)
rmb.calc_range_repeat_angles_eval_sensor_model(cloud, angles[:4].copy(), obs[:4].copy(), wc)
print("late round-1 paths ok", float(fan.max()), float(wb.mean()), float(od.mean().item()), float(wc.mean()))

# ---- round 2 additions --------------------------------------------------------------------------------------------
import tempfile  # noqa: E402

# CDDT query index (forced onto a small table) + cddt_batch_kernel (batches >= 65536 rays), plain and pruned
cq = torch.from_numpy(wl.random_queries(W, H, 70000, seed=9)).cuda()
co = torch.empty(len(cq), dtype=torch.float32, device="cuda")
for pruned in (False, True):
    ci = rl.PyCDDTCast(omap, 256.0, 24)
    if pruned:
        ci.prune()
    ci.calc_range_many_grid(cq, co)      # direct search in the batch kernel
    ci.set_spatial_sort(2)               # index: bin records + 16-bit skip codes
    ci.calc_range_many_grid(cq, co)
    ci.calc_range_many_grid(q, out)      # small batch through cast_kernel with the index
    with tempfile.TemporaryDirectory() as d:  # binary checkpoint
        ci.save(os.path.join(d, "t.rlcddt"))
        cl = rl.PyCDDTCast.load(omap, os.path.join(d, "t.rlcddt"))
        cl.calc_range_many_grid(q, out)
    ci.synchronize()
# distance transform variants: integer form with direct pass 1 / envelope pass 1 / segments, double-precision form
for env in ({}, {"RL_EDT_DIRECT_PASS1": "0"}, {"RL_EDT_DIRECT_PASS1": "0", "RL_EDT_SEGMENTS": "3"}, {"RL_EDT_EXACT_DIV": "1"}):
    os.environ.update(env)
    rl.PyRayMarchingGPU(omap, 256.0).distance_transform()
    for k in env:
        os.environ.pop(k, None)
# persistent kernels with claimed pieces: lidar fans (rm_persist_kernel<ANGLES>) and BL
fp = torch.from_numpy(wl.pf_particles_uniform(occ, 12000, seed=10)).cuda()
fa = torch.from_numpy(angles).cuda()
fo = torch.empty(len(fp) * len(angles), dtype=torch.float32, device="cuda")
rm2 = rl.PyRayMarchingGPU(omap, 256.0)
rm2.calc_range_repeat_angles(fp, fa, fo)
bl2 = rl.PyBresenhamsLine(omap, 256.0)
bl2.calc_range_many_grid(qd, od)
bl2.synchronize()
# particle-filter steps
pw = np.random.default_rng(11).uniform(0, 1, 5000)
pp = wl.pf_particles_uniform(occ, 5000, seed=12)
po = np.empty_like(pp)
rm2.normalize_weights(pw, 1.0 / 2.2)
rm2.resample(pp, pw, po, 0.3)
rm2.motion_update(po, 0.2, 0.1, 0.05, np.zeros_like(po))
rm2.synchronize()
print("round-2 paths ok", float(co.mean().item()), float(fo.mean().item()), float(po.mean()))

# round 2, second half: fused_overlap_kernel (ninth warp forms the products behind named barriers, two value buffers):
# RM with many beams (one particle per group) and CDDT / PCDDT with 60 beams (four particles per group, partial last group)
ma = wl.lidar_angles(700)
mo = np.linspace(3, 200, 700).astype(np.float32)
dp = wl.pf_particles_uniform(occ, 2300, seed=13)
dw = np.empty(len(dp), np.float64)
rm2.set_sensor_model(table)
rm2.calc_range_repeat_angles_eval_sensor_model(dp, ma, mo, dw)
cdp = wl.pf_particles_uniform(occ, 9003, seed=14)
cw = np.empty(len(cdp), np.float64)
for pruned in (False, True):
    c2 = rl.PyCDDTCast(omap, 256.0, 24)
    if pruned:
        c2.prune()
    c2.set_sensor_model(table)
    c2.calc_range_repeat_angles_eval_sensor_model(cdp, angles, obs, cw)
print("fused_overlap paths ok", float(dw.mean()), float(cw.mean()))

# deep update as two kernels (launch_fused_twostep): fan cast into the scratch array + eval_overlap_kernel, RM and CDDT,
# and the streaming eval_sensor_model on a batch of >= 64 particles per SM
tp = wl.pf_particles_uniform(occ, 2750, seed=15)
tw = np.empty(len(tp), np.float64)
rm2.calc_range_repeat_angles_eval_sensor_model(tp, ma, mo, tw)
c3 = rl.PyCDDTCast(omap, 256.0, 24)
c3.set_sensor_model(table)
cp3 = wl.pf_particles_uniform(occ, 32000, seed=16)
cw3 = np.empty(len(cp3), np.float64)
c3.calc_range_repeat_angles_eval_sensor_model(cp3, angles, obs, cw3)
er = np.random.default_rng(17).uniform(0, 250, 12000 * 60).astype(np.float32)
ew = np.empty(12000, np.float64)
rm2.eval_sensor_model(obs, er, ew, 60, 12000)
print("two-kernel deep update paths ok", float(tw.mean()), float(cw3.mean()), float(ew.mean()))
