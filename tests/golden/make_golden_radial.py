"""Golden vectors for calc_range_many_radial_optimized (RangeLib.h:616-676), produced by RUNNING THE UNMODIFIED
REFERENCE (oracle/_ref/libref_strict.so).  Output: tests/golden/vectors_radial.npz.

The reference writes a pair's second beam at a + index_offset without checking it against num_rays, so some
configurations run past the particle's row (into the next row, which that particle then rewrites, and past the
end of the buffer for the last particle): the reference is given a padded buffer and only the N*num_rays prefix
is kept.  Beams the reference never writes keep the fill value -7.

    python tests/golden/make_golden_radial.py        (authoring container only: needs oracle/_ref)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from range_libc_b200 import workloads as wl  # noqa: E402

FILL = -7.0
CONFIGS = [(60, -0.75 * np.pi, 0.75 * np.pi), (100, -np.pi, np.pi), (37, -1.0, 2.0), (7, 0.0, 6.0),
           (271, -2.35619449615, 2.35619449615)]
WORLD_ROT = (0.05, 0.3, -3.0, 2.0, float(np.float32(np.sin(0.3))), float(np.float32(np.cos(0.3))))
KINDS = [("bl", ref.BL), ("rm", ref.RM), ("cddt", ref.CDDT), ("pcddt", ref.PCDDT)]
MAPS = ["basement_hallways_10cm", "basement_hallways_5cm"]
N_PART = 48


def main():
    assert ref.available("strict")
    out = {"fill": np.float32(FILL), "configs": np.array(CONFIGS, np.float64), "world_rot": np.array(WORLD_ROT, np.float32)}
    for name in MAPS:
        occ = wl.load_map(name)
        parts = wl.pf_particles_uniform(occ, N_PART, seed=31)
        out[name + "/particles"] = parts
        out[name + "/particles_rot"] = wl.grid_to_world(parts, WORLD_ROT[0], WORLD_ROT[2], WORLD_ROT[3], WORLD_ROT[1])
        for kn, kind in KINDS:
            for wname in ("id", "rot"):
                rmap = ref.RefMap(occ=occ)
                if wname == "rot":
                    rmap.set_world(*WORLD_ROT)
                meth = ref.RefMethod(kind, rmap, 500.0, 108)
                ins = parts if wname == "id" else out[name + "/particles_rot"]
                for ci, (n, lo, hi) in enumerate(CONFIGS):
                    buf = np.full(N_PART * n + 4 * n + 4096, FILL, np.float32)
                    meth.calc_range_many_radial_optimized(n, lo, hi, ins, buf)
                    out["%s/%s/%s/%d" % (name, kn, wname, ci)] = buf[: N_PART * n].copy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "vectors_radial.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
