"""Run under torchrun on >= 2 GPUs:  torchrun --nproc-per-node 2 tests/multi_gpu_check.py
Checks the sharded fused sensor update (peer-store epilogue and the NCCL path) against the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import range_libc_b200 as rl  # noqa: E402
from range_libc_b200 import parallel, workloads as wl  # noqa: E402
from oracle import port  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    occ = wl.load_map("basement_hallways_5cm")
    n_total, M = 4001, 60
    particles = wl.pf_particles_uniform(occ, n_total, seed=5)
    angles_h = wl.lidar_angles(M)
    obs_h = np.linspace(10, 400, M).astype(np.float32)
    table = wl.sensor_table(501)
    omap = rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
    rm = rl.PyRayMarchingGPU(omap, 500.0, device=local)
    rm.set_sensor_model(table)
    rm.set_stream(torch.cuda.current_stream().cuda_stream)
    angles = torch.from_numpy(angles_h).to(dev)
    obs = torch.from_numpy(obs_h).to(dev)
    lo, hi = parallel.particle_slice(n_total, rank, world)
    mine = torch.from_numpy(particles[lo:hi]).to(dev)
    ora = port.Oracle(port.RM, occ, 500.0, threads=4)
    ora.set_sensor_model(table)
    ref = ora.calc_range_repeat_angles_eval_sensor_model(particles, angles_h, obs_h)
    results = {}

    def compute(local_particles, out_local):
        rm.calc_range_repeat_angles_eval_sensor_model(local_particles, angles, obs, out_local)

    upd = parallel.ShardedSensorUpdate(n_total, compute, device=dev)
    w = upd.update(mine)
    torch.cuda.synchronize()
    results["nccl"] = np.array_equal(w.cpu().numpy().view(np.uint64), ref.view(np.uint64))
    try:
        peer = parallel.PeerStoreSensorUpdate(n_total, rm, angles, obs, device=dev)
        w2 = peer.update(mine)
        torch.cuda.synchronize()
        results["peer"] = np.array_equal(w2.cpu().numpy().view(np.uint64), ref.view(np.uint64))
    except Exception as ex:  # noqa: BLE001
        results["peer"] = "unavailable: %s" % str(ex).splitlines()[0]
    try:
        sig = parallel.SignalledSensorUpdate(n_total, rm, angles, obs, device=dev)
        ok_all = True
        for it in range(6):  # several epochs back to back: both buffers, flag reuse
            shift = torch.tensor([0.25 * it, -0.5 * it, 0.01 * it], device=dev)
            w3 = sig.update(mine + shift, wait=True)
            torch.cuda.synchronize()
            if it in (0, 5):
                ref_it = ora.calc_range_repeat_angles_eval_sensor_model(
                    (torch.from_numpy(particles).to(dev) + shift).cpu().numpy(), angles_h, obs_h)
                ok_all &= bool(np.array_equal(w3.cpu().numpy().view(np.uint64), ref_it.view(np.uint64)))
        results["signalled"] = ok_all
    except Exception as ex:  # noqa: BLE001
        results["signalled"] = "unavailable: %s" % str(ex).splitlines()[0]
    try:
        pipe = parallel.PipelinedPeerStoreUpdate(n_total, rm, angles, obs, device=dev)
        ok_all = True
        outs = []
        for it in range(5):
            shift = torch.tensor([0.25 * it, -0.5 * it, 0.01 * it], device=dev)
            buf, ev = pipe.update(mine + shift)
            if it in (3, 4):  # the last use of each buffer
                outs.append((it, buf, ev))
        pipe.finish()
        torch.cuda.synchronize()
        for it, buf, ev in outs:
            shift = torch.tensor([0.25 * it, -0.5 * it, 0.01 * it], device=dev)
            ref_it = ora.calc_range_repeat_angles_eval_sensor_model(
                (torch.from_numpy(particles).to(dev) + shift).cpu().numpy(), angles_h, obs_h)
            ok_all &= bool(np.array_equal(buf.cpu().numpy().view(np.uint64), ref_it.view(np.uint64)))
        results["pipelined"] = ok_all
    except Exception as ex:  # noqa: BLE001
        results["pipelined"] = "unavailable: %s" % str(ex).splitlines()[0]
    # the sharded update through HOST buffers (one blocking call per rank), both bindings; 1080 beams: the
    # one-particle-per-CTA shape of BASELINE config 5
    try:
        sys.path.insert(0, os.path.join(ROOT, "range_libc_b200", "pywrapper"))
        import range_libc as cy
        for label, mod, m_beams in (("host_ctypes", rl, M), ("host_cython", cy, M), ("host_cython_1080", cy, 1080)):
            cmap = mod.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
            meth = mod.PyRayMarchingGPU(cmap, 500.0)
            meth.set_sensor_model(table)
            n_sub = n_total if m_beams == M else 301
            a_h = wl.lidar_angles(m_beams)
            o_h = np.linspace(10, 400, m_beams).astype(np.float32)
            host = parallel.HostShardedSensorUpdate(n_sub, meth, device=dev)
            ok_all = True
            for it in range(3):  # three epochs: both buffers, flag reuse
                pts = particles[:n_sub].copy()
                pts[:, 0] += 0.25 * it
                w_all = np.zeros(n_sub, np.float64)
                host.update(np.ascontiguousarray(pts[host.lo:host.hi]), a_h, o_h, w_all)
                ref_it = ora.calc_range_repeat_angles_eval_sensor_model(pts, a_h, o_h)
                ok_all &= bool(np.array_equal(w_all.view(np.uint64), ref_it.view(np.uint64)))
            results[label] = ok_all
            dist.barrier()
    except Exception as ex:  # noqa: BLE001
        results["host"] = "FAILED: %s" % str(ex).splitlines()[0]
    print("rank %d/%d: %s" % (rank, world, results), flush=True)
    ok = results["nccl"] is True and all(v is True or (isinstance(v, str) and v.startswith("unavailable")) for v in results.values())
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
