"""Golden digests for the big-map configurations (BASELINE configs 3 and 5 sizes), produced by the
CPU oracle restatement (oracle/rangelib_oracle.c, itself pinned bit-for-bit against the unmodified
reference on the small maps; the reference's own PCDDT prune of gigantic_map takes ~24 min and
> 5 GB).  Writes tests/golden/big_digests.json:
  sha256 of the distance transform, of the CDDT and PCDDT CSR offsets / values, and the reference
  ranges of 20000 seeded queries per method.

    python tests/golden/make_golden_big.py [gigantic_map huge_map ...]
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import port  # noqa: E402
from range_libc_b200 import workloads as wl  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "big_digests.json")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    names = sys.argv[1:] or ["huge_map", "gigantic_map"]
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in names:
        occ = wl.load_map(name)
        W, H = occ.shape
        q = wl.random_queries(W, H, 20000, seed=777)
        ent = {"width": W, "height": H}
        t = time.time()
        rm = port.Oracle(port.RM, occ, 500.0, threads=8)
        ent["dt_sha256"] = sha(rm.dt())
        ent["rm_ranges_sha256"] = sha(rm.calc_range_many(q))
        print(name, "dt", time.time() - t, flush=True)
        del rm
        t = time.time()
        bl = port.Oracle(port.BL, occ, 500.0, threads=8)
        ent["bl_ranges_sha256"] = sha(bl.calc_range_many(q))
        del bl
        cd = port.Oracle(port.CDDT, occ, 500.0, 108, threads=8)
        widths, trans, offsets, values = cd.cddt_table()
        ent["cddt_nvalues"] = int(len(values))
        ent["cddt_offsets_sha256"] = sha(offsets)
        ent["cddt_values_sha256"] = sha(values)
        ent["cddt_ranges_sha256"] = sha(cd.calc_range_many(q))
        print(name, "cddt", time.time() - t, len(values), flush=True)
        t = time.time()
        cd.prune(500.0)
        widths, trans, offsets, values = cd.cddt_table()
        ent["pcddt_nvalues"] = int(len(values))
        ent["pcddt_offsets_sha256"] = sha(offsets)
        ent["pcddt_values_sha256"] = sha(values)
        ent["pcddt_ranges_sha256"] = sha(cd.calc_range_many(q))
        ent["pcddt_unassigned_index_pixels"] = int(cd.prune_unassigned())
        print(name, "pcddt", time.time() - t, len(values), flush=True)
        del cd
        res[name] = ent
        json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
