"""Synthetic workloads for tests and bench.py: map fixtures, query / particle generators and the
particle-filter sensor table.  Pure numpy; nothing here touches the GPU or the oracle.

Query distributions follow the reference's benchmark drivers
(/root/reference/includes/RangeLib.h:2025-2059 random_sample: x~U(1,W-1), y~U(1,H-1), theta~U(0,2pi);
 /root/reference/pywrapper/test.py:46 for the +-0.75*pi lidar fan).
"""
import json
import lzma
import os

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAP_DIR = os.path.join(_ROOT, "tests", "golden", "maps")


def save_map(name, occ, map_dir=MAP_DIR):
    """occ: uint8 [W, H], x-major (occ[x, y] = OMap::grid[x][y])."""
    os.makedirs(map_dir, exist_ok=True)
    occ = np.ascontiguousarray(occ, dtype=np.uint8)
    with open(os.path.join(map_dir, name + ".occ.xz"), "wb") as f:
        f.write(lzma.compress(np.packbits(occ).tobytes(), preset=9))
    idx_path = os.path.join(map_dir, "maps.json")
    idx = json.load(open(idx_path)) if os.path.exists(idx_path) else {}
    idx[name] = {"width": int(occ.shape[0]), "height": int(occ.shape[1]), "occupied": int(occ.sum())}
    json.dump(idx, open(idx_path, "w"), indent=1, sort_keys=True)


def map_names(map_dir=MAP_DIR):
    return sorted(json.load(open(os.path.join(map_dir, "maps.json"))).keys())


def load_map(name, map_dir=MAP_DIR):
    """Returns occ uint8 [W, H] x-major."""
    idx = json.load(open(os.path.join(map_dir, "maps.json")))[name]
    W, H = idx["width"], idx["height"]
    raw = lzma.decompress(open(os.path.join(map_dir, name + ".occ.xz"), "rb").read())
    occ = np.unpackbits(np.frombuffer(raw, np.uint8))[: W * H].reshape(W, H)
    assert int(occ.sum()) == idx["occupied"]
    return np.ascontiguousarray(occ)


def synthetic_map(size, seed=2026, n_segments=None, n_discs=None):
    """Border walls + seeded random axis-aligned/diagonal wall segments and discs, ~1-2 % occupied
    (SURVEY.md 8d, configs C4/C5).  Returns uint8 [size, size]."""
    rng = np.random.default_rng(seed)
    occ = np.zeros((size, size), np.uint8)
    t = max(2, size // 1024)
    occ[:t, :] = occ[-t:, :] = occ[:, :t] = occ[:, -t:] = 1
    n_segments = n_segments if n_segments is not None else size // 8
    n_discs = n_discs if n_discs is not None else size // 32
    for _ in range(n_segments):
        x0, y0 = rng.integers(0, size, 2)
        length = int(rng.integers(size // 64, size // 8))
        kind = int(rng.integers(0, 3))
        w = int(rng.integers(1, 4))
        if kind == 0:
            occ[x0:x0 + length, y0:y0 + w] = 1
        elif kind == 1:
            occ[x0:x0 + w, y0:y0 + length] = 1
        else:
            k = np.arange(length)
            sgn = 1 if rng.integers(0, 2) else -1
            xs = np.clip(x0 + k, 0, size - 1)
            ys = np.clip(y0 + sgn * k, 0, size - 1)
            for d in range(w + 1):
                occ[xs, np.clip(ys + d, 0, size - 1)] = 1
    yy, xx = np.mgrid[-40:41, -40:41]
    for _ in range(n_discs):
        cx, cy = rng.integers(41, size - 41, 2)
        r = int(rng.integers(3, 40))
        mask = (xx * xx + yy * yy) <= r * r
        occ[cx - 40:cx + 41, cy - 40:cy + 41] |= mask.astype(np.uint8)
    return occ


def flip_blocks(occ, frame, seed=2026, n_blocks=64, block=16):
    """Per-frame occupancy update for the dynamic-map config (C4): toggles n_blocks seeded, block-aligned
    (hence non-overlapping) block x block patches.  Returns list of (x0, y0, patch uint8 [block, block]) and
    applies them to `occ`."""
    rng = np.random.default_rng(seed * 1000003 + frame)
    W, H = occ.shape
    gx, gy = (W - 2) // block, (H - 2) // block
    picks = rng.choice(gx * gy, size=min(n_blocks, gx * gy), replace=False)
    out = []
    for k in picks:
        x0 = 1 + int(k // gy) * block
        y0 = 1 + int(k % gy) * block
        patch = (1 - occ[x0:x0 + block, y0:y0 + block]).astype(np.uint8)
        occ[x0:x0 + block, y0:y0 + block] = patch
        out.append((x0, y0, np.ascontiguousarray(patch)))
    return out


def random_queries(W, H, n, seed=12345):
    """float32 [n, 3] (x, y, theta) in GRID coordinates, the reference's random_sample distribution."""
    rng = np.random.default_rng(seed)
    q = np.empty((n, 3), np.float32)
    q[:, 0] = rng.uniform(1.0, max(W - 1.0, 1.0), n)
    q[:, 1] = rng.uniform(1.0, max(H - 1.0, 1.0), n)
    q[:, 2] = rng.uniform(0.0, 2.0 * np.pi, n)
    return q


def grid_to_world(q_grid, scale=1.0, ox=0.0, oy=0.0, angle=0.0):
    """Inverse of the reference's world->grid conversion (RangeLib.h:464-475):
         x' = (xw-ox)/s, y' = (yw-oy)/s, (x, y) = R(angle) (x', y'), calc_range(y, x, -thw - angle - 3pi/2).
    Given grid poses (gx, gy, gth) as calc_range should see them, returns world poses that map
    onto them up to float rounding.  `angle` is OMap::world_angle (sin/cos = sin/cos(angle))."""
    q = np.asarray(q_grid, np.float64)
    c, s = np.cos(angle), np.sin(angle)
    x, y = q[:, 1], q[:, 0]  # calc_range(y, x, .): grid-x is the rotated y, grid-y the rotated x
    xp = c * x + s * y
    yp = -s * x + c * y
    w = np.empty(q.shape, np.float32)
    w[:, 0] = xp * scale + ox
    w[:, 1] = yp * scale + oy
    w[:, 2] = (-angle - 1.5 * np.pi) - q[:, 2]
    return w


def lidar_angles(m, fov=0.75 * np.pi):
    return np.linspace(-fov, fov, m).astype(np.float32)


def free_cells(occ):
    xs, ys = np.nonzero(occ == 0)
    return xs, ys


def pf_particles_uniform(occ, n, seed=7):
    """'global init' particles: uniform over free cells, theta~U(0,2pi); WORLD == grid frame
    ordering expected by numpy_calc_range_angles under identity world params: (x_w, y_w) with the
    reference's x/y swap means column 0 is grid-y... we simply sample both axes from free cells
    transposed so that the *swapped* pose lies in free space."""
    rng = np.random.default_rng(seed)
    xs, ys = free_cells(occ)
    pick = rng.integers(0, len(xs), n)
    p = np.empty((n, 3), np.float32)
    # numpy_calc_range_angles calls calc_range(y, x, .): world x -> grid y, world y -> grid x
    p[:, 0] = ys[pick] + rng.uniform(0.05, 0.95, n)
    p[:, 1] = xs[pick] + rng.uniform(0.05, 0.95, n)
    p[:, 2] = rng.uniform(0.0, 2.0 * np.pi, n)
    return p


def pf_particles_tracking(occ, n, seed=7, sigma_xy=10.0, sigma_th=0.2, dt=None):
    """'tracking' particles: Gaussian cloud about a free hallway pose (a free cell whose distance to
    the nearest wall is 5..25 px if `dt` is given).  Returns (particles [n,3], centre pose [3])."""
    rng = np.random.default_rng(seed)
    if dt is not None:
        xs, ys = np.nonzero((dt > 5.0) & (dt <= 25.0))
    else:
        xs, ys = free_cells(occ)
    k = int(rng.integers(0, len(xs)))
    centre = np.array([ys[k] + 0.5, xs[k] + 0.5, rng.uniform(0, 2 * np.pi)], np.float32)
    p = np.empty((n, 3), np.float32)
    p[:, 0] = centre[0] + rng.normal(0, sigma_xy, n)
    p[:, 1] = centre[1] + rng.normal(0, sigma_xy, n)
    p[:, 2] = centre[2] + rng.normal(0, sigma_th, n)
    W, H = occ.shape
    p[:, 0] = np.clip(p[:, 0], 1.0, H - 2.0)
    p[:, 1] = np.clip(p[:, 1], 1.0, W - 2.0)
    return p, centre


def sensor_table(K=501, sigma=8.0, z_short=0.01, z_max=0.07, z_rand=0.12, z_hit=0.75):
    """K x K float64 table[observed r][expected d], the usual racecar particle-filter recipe
    (table generation is downstream of the reference -- any strictly positive table
    exercises the same lookup/product).  Columns normalised to sum 1."""
    r = np.arange(K, dtype=np.float64)[:, None]
    d = np.arange(K, dtype=np.float64)[None, :]
    p = z_hit * np.exp(-((r - d) ** 2) / (2.0 * sigma * sigma)) / (sigma * np.sqrt(2.0 * np.pi))
    with np.errstate(divide="ignore", invalid="ignore"):
        short = np.where((r < d) & (d > 0), 2.0 * z_short * (d - r) / np.where(d > 0, d, 1.0), 0.0)
    p = p + short
    p = p + np.where(r == K - 1, z_max, 0.0)
    p = p + np.where(r < K - 1, z_rand / (K - 1.0), 0.0)
    p = p / p.sum(axis=0, keepdims=True)
    return np.ascontiguousarray(p)
