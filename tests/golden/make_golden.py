"""Generates the committed fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE
(oracle/_ref/libref_strict.so, built from /root/reference by oracle/Makefile).

    python tests/golden/make_golden.py            # maps + vectors
    python tests/golden/make_golden.py --big      # additionally the 10976^2 / 4128x10976 maps

Outputs
  tests/golden/maps/<name>.occ.xz + maps.json   occupancy grids exactly as the reference's
                                                OMap(png, threshold) decodes them (x-major bits)
  tests/golden/vectors_<name>.npz               seeded queries + the reference's answers for
                                                BL / RM / CDDT / PCDDT, world and grid
                                                coordinates, the angle fan, sensor-model weights;
                                                sha256 of the distance transform and CDDT tables
Needs /root/reference (maps) -- run in the authoring container only.
"""
import argparse
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from range_libc_b200 import workloads as wl  # noqa: E402

REF_MAPS = "/root/reference/maps"
GOLD = os.path.join(ROOT, "tests", "golden")

SMALL = [("basement_hallways_10cm", 128), ("basement_hallways_5cm", 128), ("small.map", 128), ("quad.map", 128),
         ("single_pixel.map", 128), ("basement_fixed_rectangle", 128), ("synthetic.map", 1)]
BIG = [("gigantic_map", 128), ("huge_map", 128)]
VECTOR_MAPS = ["basement_hallways_10cm", "basement_hallways_5cm", "small.map", "basement_fixed_rectangle"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def vectors_for(name, occ, n=3000, n_particles=64, n_angles=30, max_range=500.0, td=108):
    W, H = occ.shape
    out = {}
    q = wl.random_queries(W, H, n, seed=4242)
    # a few adversarial headings: cardinal directions and exact bin boundaries
    q[:16, 2] = np.array([0, np.pi / 2, np.pi, 1.5 * np.pi, 2 * np.pi, -np.pi / 2, 7.0, -7.0, 1e-3, np.pi - 1e-6,
                          np.pi / 108, 3 * np.pi / 108, 12.5, -12.5, 0.25 * np.pi, 0.75 * np.pi], np.float32)
    out["queries"] = q
    world = dict(scale=0.05, angle=0.0, ox=-30.0, oy=-30.0, sin_a=0.0, cos_a=1.0)
    qw = wl.grid_to_world(q, world["scale"], world["ox"], world["oy"])
    out["queries_world"] = qw
    out["world"] = np.array([world[k] for k in ("scale", "angle", "ox", "oy", "sin_a", "cos_a")], np.float32)
    # rotated world frame as the ROS constructor would set it (RangeLibc.pyx:160-166)
    ang = -0.3
    world_rot = np.array([0.1, ang, 2.5, -4.0, np.sin(ang), np.cos(ang)], np.float32)
    out["world_rot"] = world_rot
    qr = wl.grid_to_world(q, 0.1, 2.5, -4.0, ang)
    out["queries_world_rot"] = qr
    parts = wl.pf_particles_uniform(occ, n_particles, seed=99)
    angles = wl.lidar_angles(n_angles)
    out["particles"] = parts
    out["angles"] = angles
    table = wl.sensor_table(int(max_range) + 1)
    obs = (np.random.default_rng(5).uniform(0, max_range, n_angles)).astype(np.float32)
    out["obs"] = obs
    rmap = ref.RefMap(occ=occ)
    for kind, kn in [(ref.BL, "bl"), (ref.RM, "rm"), (ref.CDDT, "cddt"), (ref.PCDDT, "pcddt")]:
        rmap.set_world()
        meth = ref.RefMethod(kind, rmap, max_range, td)
        # NB the reference copies the OMap (and its world params) at construction, so a fresh
        # method is built per world setting.
        out[kn + "_grid"] = meth.calc_range_many(q)
        out[kn + "_angles"] = meth.numpy_calc_range_angles(parts, angles)
        meth.set_sensor_model(table)
        out[kn + "_weights_fused"] = meth.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs)
        out[kn + "_weights_two_step"] = meth.eval_sensor_model(obs, out[kn + "_angles"], n_angles, n_particles)
        if kind == ref.RM:
            out["dt_sha256"] = np.array(sha(meth.dt()))
            out["dt_sample"] = meth.dt()[:: max(1, W // 37), :: max(1, H // 41)].copy()
        if kind in (ref.CDDT, ref.PCDDT):
            widths, trans, offsets, values = meth.cddt_table(td)
            out[kn + "_widths"] = widths
            out[kn + "_trans"] = trans
            out[kn + "_offsets_sha256"] = np.array(sha(offsets))
            out[kn + "_values_sha256"] = np.array(sha(values))
            out[kn + "_nvalues"] = np.array(len(values))
        rmap.set_world(**world)
        meth_w = ref.RefMethod(kind, rmap, max_range, td)
        out[kn + "_world"] = meth_w.numpy_calc_range(qw)
        rmap.set_world(*[float(v) for v in world_rot])
        meth_r = ref.RefMethod(kind, rmap, max_range, td)
        out[kn + "_world_rot"] = meth_r.numpy_calc_range(qr)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    args = ap.parse_args()
    assert ref.available("strict"), "build oracle/_ref first (make -f oracle/Makefile)"
    for name, thr in SMALL + (BIG if args.big else []):
        rmap = ref.RefMap(png=os.path.join(REF_MAPS, name + ".png"), threshold=thr)
        wl.save_map(name, rmap.occ())
        print("map", name, rmap.width, rmap.height)
    for name in VECTOR_MAPS:
        occ = wl.load_map(name)
        v = vectors_for(name, occ)
        np.savez_compressed(os.path.join(GOLD, "vectors_%s.npz" % name), **v)
        print("vectors", name, {k: (v[k].shape if hasattr(v[k], "shape") else v[k]) for k in list(v)[:4]})


if __name__ == "__main__":
    main()
