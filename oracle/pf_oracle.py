"""CPU checker for the particle-filter steps of range_libc_b200/csrc/rl_pf.cu.

TEST INFRASTRUCTURE ONLY -- imported by tests/ (and nothing in the product path).

PARITY UNPINNED: these steps are NOT in the reference (range_libc ends at the per-particle weights,
/root/reference/includes/RangeLib.h:558-612); they restate what its downstream user, the mit-racecar particle_filter
MCL loop named in the reference's README.md:57, does between two sensor updates (weights ** (1 / squash), normalise,
resample, odometry motion model), with the definitions written down in rl_pf.cu.  There is no reference source or
golden vector to pin them against; the resampling and motion steps are specified so that they are bit-reproducible.
"""
import ctypes

import numpy as np

_libm = ctypes.CDLL("libm.so.6")
_libm.sinf.restype = ctypes.c_float
_libm.sinf.argtypes = [ctypes.c_float]
_libm.cosf.restype = ctypes.c_float
_libm.cosf.argtypes = [ctypes.c_float]


def normalize(weights, inv_squash=1.0):
    """(normalised weights, sum of the squashed weights) in float64"""
    w = np.asarray(weights, np.float64)
    if inv_squash != 1.0:
        w = np.power(w, inv_squash)
    s = float(np.sum(w))
    return w / s, s


def resample_indices(weights, u0):
    """systematic resampling in 2^-40 fixed point (rl_pf.cu): index of the particle copied to every output slot"""
    w = np.asarray(weights, np.float64)
    n = len(w)
    v = w * 1099511627776.0
    f = np.where(v > 0.0, np.trunc(np.where(v > 0.0, v, 0.0)), 0.0).astype(np.uint64)
    c = np.cumsum(f, dtype=np.uint64)
    total = np.float64(c[-1])
    thr = (np.float64(u0) + np.arange(n, dtype=np.float64)) / np.float64(n) * total
    idx = np.searchsorted(c.astype(np.float64), thr, side="right")
    return np.minimum(idx, n - 1)


def resample(particles, weights, u0):
    return np.asarray(particles, np.float32)[resample_indices(weights, u0)]


def motion(particles, dx, dy, dtheta, noise=None):
    """float32 odometry step with libm's sinf / cosf, operation order of pf_motion_kernel"""
    p = np.asarray(particles, np.float32)
    th = p[:, 2]
    sn = np.array([_libm.sinf(float(t)) for t in th], np.float32)
    cs = np.array([_libm.cosf(float(t)) for t in th], np.float32)
    dx, dy, dtheta = np.float32(dx), np.float32(dy), np.float32(dtheta)
    x = p[:, 0] + (cs * dx - sn * dy)
    y = p[:, 1] + (sn * dx + cs * dy)
    t = th + dtheta
    if noise is not None:
        nz = np.asarray(noise, np.float32)
        x, y, t = x + nz[:, 0], y + nz[:, 1], t + nz[:, 2]
    return np.stack([x, y, t], 1).astype(np.float32)
