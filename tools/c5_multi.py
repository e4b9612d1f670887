"""BASELINE config 5 across GPUs: 10^6 particles x 1080 beams, RM fused with the sensor model on a synthetic 8192^2
grid, particles sharded over the ranks (map, distance transform and table replicated), weights all-gathered by the
fused kernel's peer stores over NVLink + a symmetric-memory barrier.  Launch with torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/c5_multi.py

Rank 0 prints one JSON line: time per update (max over ranks, CUDA events) and whole-job rays/s."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import range_libc_b200 as rl  # noqa: E402
from range_libc_b200 import parallel, workloads as wl  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_total, beams = 1000000, 1080
    occ = wl.synthetic_map(8192, seed=2026)
    rm = rl.PyRayMarchingGPU(rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool))), 500.0, device=local)
    rm.set_sensor_model(wl.sensor_table(501))
    st = torch.cuda.current_stream()
    rm.set_stream(st.cuda_stream)
    lo, hi = parallel.particle_slice(n_total, rank, world)
    parts = torch.from_numpy(wl.pf_particles_uniform(occ, n_total, seed=4)[lo:hi].copy()).to(dev)
    angles = torch.from_numpy(wl.lidar_angles(beams)).to(dev)
    obs = torch.from_numpy(np.clip(150 + 100 * np.sin(np.linspace(0, 6, beams)), 0, 500).astype(np.float32)).to(dev)
    if world > 1:
        upd = parallel.PeerStoreSensorUpdate(n_total, rm, angles, obs, device=dev)
        step = lambda: upd.update(parts)  # noqa: E731
    else:
        w = torch.empty(n_total, dtype=torch.float64, device=dev)
        step = lambda: rm.calc_range_repeat_angles_eval_sensor_model(parts, angles, obs, w)  # noqa: E731
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    k = 10
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(k):
        step()
    b.record(st)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / k
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({"workload": "C5: 1e6 particles x 1080 beams RM fused, synthetic 8192^2, sharded", "n_gpus": world,
                          "ms_per_update": ms, "rays_per_s": n_total * beams / (ms * 1e-3),
                          "gather": "fused peer stores + symmetric-memory barrier" if world > 1 else "none"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
