"""Multi-process host logic of the sharded particle-filter update, world_size 2 and 3 over gloo on CPU.
The compute callable is the CPU oracle here (tests may use it); on GPUs it is the fused CUDA kernel."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from range_libc_b200 import parallel, workloads as wl  # noqa: E402


def test_particle_slices_cover_and_partition():
    for n in (0, 1, 7, 4000, 1000003):
        for world in (1, 2, 3, 8):
            edges = [parallel.particle_slice(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (a, b), (c, d) in zip(edges, edges[1:]):
                assert b == c and b - a >= d - c >= 0
            assert sum(parallel.shard_sizes(n, world)) == n


def _worker(rank, world, port, n_total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import port as oracle_port
    occ = wl.load_map("basement_hallways_10cm")
    particles = wl.pf_particles_uniform(occ, n_total, seed=5)
    angles = wl.lidar_angles(24)
    obs = np.linspace(10, 400, 24).astype(np.float32)
    ora = oracle_port.Oracle(oracle_port.RM, occ, 500.0)
    ora.set_sensor_model(wl.sensor_table(501))

    def compute(local, out_local):
        out_local.copy_(torch.from_numpy(ora.calc_range_repeat_angles_eval_sensor_model(local, angles, obs)))

    upd = parallel.ShardedSensorUpdate(n_total, compute)
    lo, hi = parallel.particle_slice(n_total, rank, world)
    w = upd.update(particles[lo:hi])
    np.save(os.path.join(out_dir, "w%d.npy" % rank), w.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 64), (2, 65), (3, 100)])
def test_sharded_update_gloo(tmp_path, world, n_total):
    port = 29500 + (os.getpid() % 2000) + world * 7 + n_total % 5
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    from oracle import port as oracle_port
    occ = wl.load_map("basement_hallways_10cm")
    particles = wl.pf_particles_uniform(occ, n_total, seed=5)
    ora = oracle_port.Oracle(oracle_port.RM, occ, 500.0)
    ora.set_sensor_model(wl.sensor_table(501))
    ref = ora.calc_range_repeat_angles_eval_sensor_model(particles, wl.lidar_angles(24),
                                                         np.linspace(10, 400, 24).astype(np.float32))
    for r in range(world):
        w = np.load(os.path.join(str(tmp_path), "w%d.npy" % r))
        assert np.array_equal(w.view(np.uint64), ref.view(np.uint64)), "rank %d" % r
