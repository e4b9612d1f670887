"""CPU simulation of the cooperative tail's probe schedule on the longest rays of bench.py's particle sets
(development aid for the open item in DESIGN.md section 10).

  current   every batch is anchored at the exact sample where the previous one ran out: its 64 probes
            (t_b + 0.75 j) are loaded, THEN replayed -- one L2 round trip per batch on the ray's critical path.
  grid      probes form one continuous grid t_0 + 0.75 j; window k+1 is requested while window k is replayed, so
            only a sample that neither the current nor the next window holds (a clipped corner, or a step that
            jumps more than a window) costs a synchronous round trip (re-anchor at the exact sample).

Prints, for rays with more than 60 steps: steps, batches (= exposed round trips) of the current scheme and
exposed round trips of the grid scheme."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import port  # noqa: E402
from range_libc_b200 import workloads as wl  # noqa: E402

f32 = np.float32
SP, NP = f32(0.75), 64


def cell(x0, y0, dx, dy, t):
    return int(f32(x0 + f32(dx * t))), int(f32(y0 + f32(dy * t)))


def simulate(dt, W, H, x0, y0, dx, dy, t, mr, scheme):
    """returns (steps, exposed round trips)"""
    steps = trips = 0
    anchor, window, tried_next = t, 0, False
    while True:
        if scheme == "current":
            anchor, window = t, 0
        lo = anchor + f32(window * NP) * SP
        probes = {}
        for j in range(NP if scheme == "current" else 2 * NP):  # grid: current + prefetched window
            s = f32(lo + f32(j) * SP)
            probes[cell(x0, y0, dx, dy, s)] = True
        trips += 1 if (scheme == "current" or not tried_next) else 0
        progressed = False
        while t < mr:
            c = cell(x0, y0, dx, dy, t)
            if not (0 <= c[0] < W and 0 <= c[1] < H):
                return steps, trips
            if c not in probes:
                break
            d = dt[c]
            if d <= 0:
                return steps, trips
            t = f32(t + max(f32(d * f32(0.999)), f32(1.0)))
            steps += 1
            progressed = True
        else:
            return steps, trips
        if scheme == "grid":
            if progressed or not tried_next:
                # slide to the window that holds t (its loads were requested a window ago): not exposed,
                # unless t jumped beyond the prefetched window
                k = int((t - anchor) / (SP * NP))
                exposed = k > window + 1
                window, tried_next = k, not exposed
                if exposed:
                    anchor, window, tried_next = t, 0, False
            else:
                anchor, window, tried_next = t, 0, False  # clipped corner: re-anchor at the exact sample


def main():
    occ = wl.load_map(bench.MAP)
    sets, angles, _ = bench.make_inputs(occ, 4)
    o = port.Oracle(port.RM, occ, 500.0)
    dt = o.dt()
    W, H = occ.shape
    rows = []
    for si in range(4):
        P = sets[si]
        rot = f32(-3.0 * np.pi / 2.0)
        th = ((-P[:, 2] + rot)[:, None] - angles[None, :]).astype(f32)
        rays = np.empty((P.shape[0] * len(angles), 3), f32)
        rays[:, 0] = np.repeat(P[:, 1], len(angles))
        rays[:, 1] = np.repeat(P[:, 0], len(angles))
        rays[:, 2] = th.ravel()
        c = o.rm_step_counts(rays)
        for r in np.argsort(-c)[:40]:
            x0, y0, thh = rays[r]
            dx, dy = f32(port.cosf(thh)), f32(port.sinf(thh))
            s1, b1 = simulate(dt, W, H, x0, y0, dx, dy, f32(0), f32(500.0), "current")
            s2, b2 = simulate(dt, W, H, x0, y0, dx, dy, f32(0), f32(500.0), "grid")
            rows.append((c[r], s1, b1, s2, b2))
    rows = np.array(rows)
    print("rays simulated: %d   reference steps: mean %.0f max %d" % (len(rows), rows[:, 0].mean(), rows[:, 0].max()))
    print("current scheme: batches (exposed L2 round trips) mean %.1f max %d" % (rows[:, 2].mean(), rows[:, 2].max()))
    print("grid scheme:    exposed L2 round trips           mean %.1f max %d" % (rows[:, 4].mean(), rows[:, 4].max()))
    assert (rows[:, 1] == rows[:, 3]).all(), "both schedules must take the reference's steps"


if __name__ == "__main__":
    main()
