"""Multi-GPU particle-filter sensor update: one process per GPU, particles sharded across ranks, the
map / distance transform / sensor table replicated, per-particle weights gathered on every rank
(SURVEY.md section 8e).  Plain ray batches need no collective; only the weights are exchanged.

Gather paths:
  * "peer":      the epilogue of the kernel that forms the weights stores each one straight into every rank's gathered
                 array over NVLink (rl_calc_range_repeat_angles_eval_sensor_model_peers); the arrays
                 live in torch symmetric memory, and a symmetric-memory barrier orders the step.
  * "signalled": the same stores plus in-kernel epoch flags: one kernel per rank and step, no barrier launch.
  * "host":      HostShardedSensorUpdate -- the signalled path behind ONE blocking call with host (numpy) buffers,
                 rl_calc_range_repeat_angles_eval_sensor_model_sharded: what a multi-process particle filter
                 written against the reference's Python API calls.
  * "nccl":      local fused kernel, then torch.distributed.all_gather_into_tensor (NCCL on GPUs;
                 gloo in the CPU tests, where the compute callable is injected).

Streams: a method handle launches on its own private stream unless told otherwise.  Every GPU class here binds the
handle to torch's CURRENT stream at each update() (method.set_stream), so the symmetric-memory barrier, the events
and the consumer -- all on torch's current stream -- are ordered after the kernel's peer stores, and the kernel after
whatever produced its inputs.
The host-side logic here (slicing, gather bookkeeping, uneven shards) has no GPU dependency so that
it can be exercised with world_size-2 gloo tests.
"""
import numpy as np


def particle_slice(n_total, rank, world):
    """Contiguous shard [lo, hi) of rank `rank`; the first n_total % world ranks get one extra."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_total, world):
    return [particle_slice(n_total, r, world)[1] - particle_slice(n_total, r, world)[0] for r in range(world)]


class ShardedSensorUpdate:
    """weights_all = update(local_particles): every rank returns the weights of ALL particles.

    compute(local_particles, out_local) must fill out_local (1-D float64 tensor view of this rank's
    slice) with the fused range + sensor-model weights of the local particles."""

    def __init__(self, n_total, compute, group=None, device=None, dtype=None):
        import torch
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = n_total
        self.lo, self.hi = particle_slice(n_total, self.rank, self.world)
        self.sizes = shard_sizes(n_total, self.world)
        self.compute = compute
        self.even = len(set(self.sizes)) == 1
        self.weights_all = torch.empty(n_total, dtype=dtype or torch.float64, device=device)
        if not self.even:  # all_gather with uneven shards: pad to the largest shard
            self.pad = max(self.sizes)
            self.stage_local = torch.zeros(self.pad, dtype=self.weights_all.dtype, device=device)
            self.stage_all = torch.empty(self.pad * self.world, dtype=self.weights_all.dtype, device=device)

    def local_view(self):
        return self.weights_all[self.lo:self.hi]

    def update(self, local_particles):
        self.compute(local_particles, self.local_view())
        if self.world == 1:
            return self.weights_all
        if self.even:
            self.dist.all_gather_into_tensor(self.weights_all, self.local_view(), group=self.group)
        else:
            self.stage_local[: self.hi - self.lo].copy_(self.local_view())
            self.dist.all_gather_into_tensor(self.stage_all, self.stage_local, group=self.group)
            for r in range(self.world):
                lo, hi = particle_slice(self.n_total, r, self.world)
                self.weights_all[lo:hi].copy_(self.stage_all[r * self.pad: r * self.pad + (hi - lo)])
        return self.weights_all


def _bind_current_stream(method):
    """Make the handle launch on torch's current stream (see the module docstring)."""
    import torch
    method.set_stream(torch.cuda.current_stream().cuda_stream)


class PeerStoreSensorUpdate:
    """Fused compute + all-gather: the weights array lives in torch symmetric memory and the RM kernel
    writes each rank's slice into every peer directly.  GPU only.  update() runs on torch's current stream."""

    def __init__(self, n_total, method, angles, obs, group=None, device=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.dist = dist
        self.group = group or dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.n_total = n_total
        self.lo, self.hi = particle_slice(n_total, self.rank, self.world)
        self.method, self.angles, self.obs = method, angles, obs
        self.weights_all = symm_mem.empty(n_total, dtype=torch.float64, device=device)
        self.handle = symm_mem.rendezvous(self.weights_all, self.group)
        self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]

    def update(self, local_particles):
        _bind_current_stream(self.method)
        self.method.calc_range_repeat_angles_eval_sensor_model_peers(local_particles, self.angles, self.obs,
                                                                     self.peer_ptrs, self.lo)
        self.handle.barrier()  # every rank's stores have landed before anyone reads weights_all
        return self.weights_all


def _signalled_buffers(n_total, method, group, device):
    """Two gathered-weight buffers + one flag array in symmetric memory, registered with the handle."""
    import torch
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bufs, handles = [], []
    for _ in range(2):
        t = symm_mem.empty(n_total, dtype=torch.float64, device=device)
        bufs.append(t)
        handles.append(symm_mem.rendezvous(t, group))
    flags = symm_mem.empty(max(world, 2), dtype=torch.int64, device=device)
    flags.zero_()
    flag_handle = symm_mem.rendezvous(flags, group)
    torch.cuda.synchronize()
    flag_handle.barrier()  # every rank's flags are zero before anyone can signal
    torch.cuda.synchronize()
    method.peers_init([int(p) for p in handles[0].buffer_ptrs], [int(p) for p in handles[1].buffer_ptrs],
                      [int(p) for p in flag_handle.buffer_ptrs], rank)
    return bufs, handles, flags, flag_handle


class SignalledSensorUpdate:
    """Fused compute + all-gather + synchronisation in ONE kernel per rank and step
    (rl_calc_range_repeat_angles_eval_sensor_model_signalled): weights are double buffered in symmetric
    memory, completion is signalled through peer-written epoch flags, and the only extra launch is the
    consumer-side wait (update(..., wait=True)) before the gathered weights are read.  GPU only; runs on torch's
    current stream.  The consumer of step k must be ordered before update() of step k+2 (same stream, or an event):
    that launch overwrites the buffer it reads."""

    def __init__(self, n_total, method, angles, obs, group=None, device=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group or dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.n_total = n_total
        self.lo, self.hi = particle_slice(n_total, self.rank, self.world)
        self.method, self.angles, self.obs = method, angles, obs
        _bind_current_stream(method)
        self.bufs, self.handles, self.flags, self.flag_handle = _signalled_buffers(n_total, method, self.group, device)

    def update(self, local_particles, wait=True):
        _bind_current_stream(self.method)
        b = self.method.calc_range_repeat_angles_eval_sensor_model_signalled(local_particles, self.angles, self.obs, self.lo)
        if wait:
            self.method.peers_wait()
        return self.bufs[b]


class HostShardedSensorUpdate:
    """The sharded update through HOST buffers in one blocking call per rank
    (rl_calc_range_repeat_angles_eval_sensor_model_sharded): local particles in (numpy, ideally pinned), weights of
    ALL particles out (numpy).  `method` may be the ctypes mirror (range_libc_b200.Py*) or the Cython drop-in
    (range_libc.Py*): both expose peers_init and calc_range_repeat_angles_eval_sensor_model_sharded.  The handle keeps
    its private stream; the call synchronises it before returning."""

    def __init__(self, n_total, method, group=None, device=None):
        import torch.distributed as dist
        self.group = group or dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.n_total = n_total
        self.lo, self.hi = particle_slice(n_total, self.rank, self.world)
        self.method = method
        self.bufs, self.handles, self.flags, self.flag_handle = _signalled_buffers(n_total, method, self.group, device)

    def update(self, local_particles, angles, obs, weights_all):
        self.method.calc_range_repeat_angles_eval_sensor_model_sharded(local_particles, angles, obs, weights_all, self.lo)
        return weights_all


class PipelinedPeerStoreUpdate:
    """PeerStoreSensorUpdate with the barrier of step k overlapped with the compute of step k+1, for THROUGHPUT
    workloads whose steps are independent (a particle filter that resamples on step k's weights before it can start
    step k+1 gains nothing: use SignalledSensorUpdate / PeerStoreSensorUpdate).  The gathered weights are TRIPLE
    buffered in symmetric memory; the fused kernel of step k stores into buffer k % 3 on the main stream, and the
    symmetric-memory barrier that closes step k runs on a side stream while the kernel of step k+1 is computing.
    Reuse rule: the kernel of step k may overwrite buffer k % 3 (last used by step k-3) once the barrier of step k-2
    has completed -- every rank has then launched its step k-2, so, PROVIDED each rank's consumer of step j is ordered
    before its update() of step j+1 (same stream, or an event the main stream waits on), every consumer of step k-3
    has finished.  update() returns (buffer, event): the buffer holds the complete gather once the event has fired.
    GPU only; the main stream is torch's current stream."""

    NBUF = 3

    def __init__(self, n_total, method, angles, obs, group=None, device=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.torch = torch
        self.group = group or dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.n_total = n_total
        self.lo, self.hi = particle_slice(n_total, self.rank, self.world)
        self.method, self.angles, self.obs = method, angles, obs
        self.bufs, self.handles, self.ptrs = [], [], []
        for _ in range(self.NBUF):
            t = symm_mem.empty(n_total, dtype=torch.float64, device=device)
            h = symm_mem.rendezvous(t, self.group)
            self.bufs.append(t)
            self.handles.append(h)
            self.ptrs.append([int(p) for p in h.buffer_ptrs])
        self.side = torch.cuda.Stream(device=device)
        self.done = {}  # step -> event: the barrier closing that step has completed
        self.k = 0

    def update(self, local_particles):
        torch = self.torch
        main = torch.cuda.current_stream()
        _bind_current_stream(self.method)
        k = self.k
        b = k % self.NBUF
        if k - 2 in self.done:
            main.wait_event(self.done[k - 2])
        self.done.pop(k - 3, None)
        self.method.calc_range_repeat_angles_eval_sensor_model_peers(local_particles, self.angles, self.obs,
                                                                     self.ptrs[b], self.lo)
        ready = torch.cuda.Event()
        ready.record(main)
        self.side.wait_event(ready)
        with torch.cuda.stream(self.side):
            self.handles[b].barrier()
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.done[k] = ev
        self.k += 1
        return self.bufs[b], ev

    def reset(self):
        """forget the events of earlier steps.  Call only when all earlier work has completed on every rank (e.g.
        after a device synchronise + process-group barrier), and around CUDA-graph capture: a capturing stream
        must not wait on events recorded outside the capture, nor eager work on events recorded inside it."""
        self.done = {}
        self.k = 0

    def finish(self):
        """join the side stream: every gather issued so far is complete once the main stream gets here"""
        main = self.torch.cuda.current_stream()
        for ev in self.done.values():
            main.wait_event(ev)
