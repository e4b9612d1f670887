"""Latency of one blocking host-pointer particle-filter update (4000 x 60, RM, basement_hallways_5cm) through the
drop-in Cython module, for pinned and pageable caller arrays.  RL_HOST_DIRECT=0 in the environment selects the
general marshalling path (inputs packed into the staging buffer, angles / observation copied with them) for A/B."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "range_libc_b200", "pywrapper"))
import bench  # noqa: E402
import range_libc as cy  # noqa: E402
from oracle import port  # noqa: E402
from range_libc_b200 import workloads as wl  # noqa: E402


def main():
    occ = wl.load_map(bench.MAP)
    sets, angles, obs = bench.make_inputs(occ, 64)
    table = wl.sensor_table(bench.K_TABLE)
    rm = cy.PyRayMarchingGPU(cy.PyOMap(np.ascontiguousarray(occ.T.astype(bool))), bench.MAX_RANGE)
    rm.set_sensor_model(table)
    o = port.Oracle(port.RM, occ, bench.MAX_RANGE, threads=os.cpu_count())
    o.set_sensor_model(table)
    want = [o.calc_range_repeat_angles_eval_sensor_model(sets[i], angles, obs) for i in range(4)]
    for label, pin in (("pinned", True), ("pageable", False)):
        mk = (lambda a: torch.from_numpy(a.copy()).pin_memory().numpy()) if pin else (lambda a: a.copy())
        hs = [mk(sets[i]) for i in range(64)]
        ha, ho = mk(angles), mk(obs)
        hw = mk(np.zeros(bench.N_PART, np.float64))
        ok = True
        for i in range(4):
            rm.calc_range_repeat_angles_eval_sensor_model(hs[i], ha, ho, hw)
            ok = ok and np.array_equal(hw.view(np.uint64), want[i].view(np.uint64))
        for i in range(50):
            rm.calc_range_repeat_angles_eval_sensor_model(hs[i % 64], ha, ho, hw)
        n = 3000
        t0 = time.perf_counter()
        for i in range(n):
            rm.calc_range_repeat_angles_eval_sensor_model(hs[i % 64], ha, ho, hw)
        dt = (time.perf_counter() - t0) / n
        print("RL_HOST_DIRECT=%s %-8s %.2f us per update  %.2f G rays/s  bit-equal-to-oracle=%s" % (
            os.environ.get("RL_HOST_DIRECT", "1"), label, dt * 1e6, bench.N_PART * bench.N_BEAMS / dt / 1e9, ok), flush=True)


if __name__ == "__main__":
    main()
