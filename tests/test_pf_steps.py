"""Particle-filter steps either side of the sensor update (range_libc_b200/csrc/rl_pf.cu).
These are NOT in the reference; the checker is oracle/pf_oracle.py (parity unpinned, see its header).  The CPU tests
pin the oracle's own properties, the GPU tests compare the CUDA path with it through the C ABI (both bindings):
resampling and motion bit-exact, normalisation within 1e-12 relative (floating-point sum order)."""
import os
import sys

import numpy as np
import pytest

from oracle import pf_oracle
from helpers import ROOT, assert_bit_equal


def cloud(n, seed):
    rng = np.random.default_rng(seed)
    p = np.stack([rng.uniform(-30, 30, n), rng.uniform(-30, 30, n), rng.uniform(-7, 7, n)], 1).astype(np.float32)
    w = rng.uniform(0.0, 1.0, n) ** 8           # a few heavy particles, many light ones
    if n > 10:
        w[rng.integers(0, n, max(1, n // 50))] = 0.0  # and some with no weight at all
    return p, w


def test_oracle_resampling_properties():
    p, w = cloud(5000, 1)
    wn, s = pf_oracle.normalize(w, 1.0 / 2.2)
    assert abs(wn.sum() - 1.0) < 1e-12 and s > 0
    idx = pf_oracle.resample_indices(wn, 0.37)
    assert (np.diff(idx) >= 0).all(), "systematic resampling visits the particles in order"
    counts = np.bincount(idx, minlength=len(wn))
    assert (counts[wn == 0.0] == 0).all(), "a particle without weight is never drawn"
    assert (np.abs(counts - wn * len(wn)) < 1.0 + 1e-6).all(), "copies = n * w within one (low-variance resampling)"
    # uniform weights and u0 = 0.5: every particle exactly once
    assert (pf_oracle.resample_indices(np.full(1000, 1e-3), 0.5) == np.arange(1000)).all()


def test_oracle_motion_is_the_planar_odometry_step():
    p = np.array([[1.0, 2.0, 0.0], [0.0, 0.0, np.pi / 2]], np.float32)
    q = pf_oracle.motion(p, 1.0, 0.0, 0.25)
    assert np.allclose(q, [[2.0, 2.0, 0.25], [0.0, 1.0, np.pi / 2 + 0.25]], atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 37, 4000, 250000])
def test_pf_steps_on_gpu_match_the_oracle(n):
    import torch
    import range_libc_b200 as rl
    from range_libc_b200 import workloads as wl
    occ = wl.load_map("basement_hallways_10cm")
    m = rl.PyBresenhamsLine(rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool))), 100.0)
    p, w = cloud(n, n)
    # normalisation (host buffers, in place)
    for inv in (1.0, 1.0 / 2.2):
        got = w.copy()
        s = m.normalize_weights(got, inv, want_sum=True)
        want, ws = pf_oracle.normalize(w, inv)
        assert abs(s - ws) <= 1e-12 * ws
        assert np.allclose(got, want, rtol=1e-12, atol=0.0)
    wn = pf_oracle.normalize(w, 1.0 / 2.2)[0]
    # resampling: host buffers, then device buffers on the caller's stream
    for u0 in (0.0, 0.37, 0.999999):
        out = np.empty_like(p)
        m.resample(p, wn, out, u0)
        assert_bit_equal(out, pf_oracle.resample(p, wn, u0), "resample n=%d u0=%g" % (n, u0))
    dp, dw = torch.from_numpy(p).cuda(), torch.from_numpy(wn).cuda()
    dout = torch.empty_like(dp)
    m.set_stream(torch.cuda.current_stream().cuda_stream)
    m.resample(dp, dw, dout, 0.61)
    assert_bit_equal(dout.cpu().numpy(), pf_oracle.resample(p, wn, 0.61), "resample, device buffers")
    with pytest.raises(rl.RangeLibError):
        m.resample(p, wn, np.empty_like(p), 1.0)
    # motion update with and without noise
    noise = np.random.default_rng(7).normal(0, 0.05, p.shape).astype(np.float32)
    for nz in (None, noise):
        got = p.copy()
        m.motion_update(got, 0.31, -0.07, 0.043, nz)
        assert_bit_equal(got, pf_oracle.motion(p, 0.31, -0.07, 0.043, nz), "motion n=%d noise=%s" % (n, nz is not None))
    dq = torch.from_numpy(p).cuda()
    m.motion_update(dq, 0.31, -0.07, 0.043, torch.from_numpy(noise).cuda())
    assert_bit_equal(dq.cpu().numpy(), pf_oracle.motion(p, 0.31, -0.07, 0.043, noise), "motion, device buffers")
    # the Cython drop-in forwards to the same entry points
    if n == 4000:
        sys.path.insert(0, os.path.join(ROOT, "range_libc_b200", "pywrapper"))
        import range_libc as cy
        cm = cy.PyBresenhamsLine(cy.PyOMap(np.ascontiguousarray(occ.T.astype(bool))), 100.0)
        got = w.copy()
        s = cm.normalize_weights(got, 1.0 / 2.2)
        assert abs(s - pf_oracle.normalize(w, 1.0 / 2.2)[1]) <= 1e-12 * s and np.allclose(got, wn, rtol=1e-12, atol=0.0)
        out = np.empty_like(p)
        cm.resample(p, wn, out, 0.37)
        assert_bit_equal(out, pf_oracle.resample(p, wn, 0.37), "resample, Cython")
        got = p.copy()
        cm.motion_update(got, 0.31, -0.07, 0.043, noise)
        assert_bit_equal(got, pf_oracle.motion(p, 0.31, -0.07, 0.043, noise), "motion, Cython")


@pytest.mark.gpu
def test_pf_loop_keeps_particles_on_the_device():
    """One MCL iteration entirely on device buffers: motion -> fused sensor update -> squash + normalise -> resample.
    The weights equal the oracle's fused update of the moved particles, the resampled cloud the oracle's resampling."""
    import torch
    import range_libc_b200 as rl
    from range_libc_b200 import workloads as wl
    from oracle import port
    occ = wl.load_map("basement_hallways_10cm")
    rm = rl.PyRayMarchingGPU(rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool))), 200.0)
    table = wl.sensor_table(201)
    rm.set_sensor_model(table)
    rm.set_stream(torch.cuda.current_stream().cuda_stream)
    parts = wl.pf_particles_uniform(occ, 3000, seed=5)
    angles = wl.lidar_angles(30)
    obs = np.full(30, 40.0, np.float32)
    d_p = torch.from_numpy(parts).cuda()
    d_a, d_o = torch.from_numpy(angles).cuda(), torch.from_numpy(obs).cuda()
    d_w = torch.empty(len(parts), dtype=torch.float64, device="cuda")
    d_q = torch.empty_like(d_p)
    rm.motion_update(d_p, 0.5, 0.0, 0.02)
    rm.calc_range_repeat_angles_eval_sensor_model(d_p, d_a, d_o, d_w)
    rm.normalize_weights(d_w, 1.0 / 2.2)
    rm.resample(d_p, d_w, d_q, 0.25)
    moved = pf_oracle.motion(parts, 0.5, 0.0, 0.02)
    assert_bit_equal(d_p.cpu().numpy(), moved, "moved particles")
    o = port.Oracle(port.RM, occ, 200.0, 0)
    o.set_sensor_model(table)
    w_ref = o.calc_range_repeat_angles_eval_sensor_model(moved, angles, obs)
    wn = pf_oracle.normalize(w_ref, 1.0 / 2.2)[0]
    got_w = d_w.cpu().numpy()
    assert np.allclose(got_w, wn, rtol=1e-12, atol=0.0)
    assert_bit_equal(d_q.cpu().numpy(), pf_oracle.resample(moved, got_w, 0.25), "resampled cloud")
