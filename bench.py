#!/usr/bin/env python
"""Benchmark of the hot path on the BASELINE.json configuration
  "Particle-filter sensor model: 4000 particles x 60 beams calc_range_repeat_angles +
   eval_sensor_model on basement_hallways_5cm, 1 B200"  (configs[1]).

A step = one particle-filter sensor update: RM ray casts for N x M (particle, beam) pairs fused
with the sensor-table lookup and the per-particle product
(RangeMethod::calc_range_repeat_angles_eval_sensor_model, /root/reference/includes/RangeLib.h:558-612).
Metric: ray casts / s (one sensor-model evaluation per ray, so it is also evals / s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Multi-GPU (torchrun, one rank per GPU): particles shard across ranks (weak scaling: 4000 per rank),
the map/DT/table are replicated, and the per-particle weights are all-gathered every step.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAP = "basement_hallways_5cm"
N_PART, N_BEAMS, MAX_RANGE, K_TABLE = 4000, 60, 500.0, 501
METRIC = "ray casts/sec (RM, fused PF sensor-model update: calc_range_repeat_angles + eval_sensor_model)"
L2_BYTES = 126 * 1024 * 1024


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of fused_kernel<RM> from the committed ncu
    --set full capture of this same command (profiles/ncu_fused_r01.txt; cold-cache, serialised launches)."""
    p = os.path.join(ROOT, "profiles", "ncu_fused_r01.txt")
    if not os.path.exists(p):
        return None, None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, n, sectors = 0.0, 0, []
    for line in open(p):
        f = line.split()
        if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(f[1].replace(",", "")) * scale.get(f[2], 1.0)
            n += f[0] == "dram__bytes_read.sum"
        if len(f) >= 2 and f[0] == "lts__t_sectors.sum":
            sectors.append(float(f[1].replace(",", "")))
    if not n:
        return None, None
    return tot / n, (float(np.mean(sectors)) * 32.0 if sectors else None)


def make_inputs(occ, n_sets, n_part=N_PART, seed=2026):
    """n_sets particle clouds: even sets 'global init' (uniform over free cells), odd sets a 'tracking'
    Gaussian cloud around a free pose -- the two regimes of a particle filter."""
    from range_libc_b200 import workloads as wl
    sets = np.empty((n_sets, n_part, 3), np.float32)
    base_u = wl.pf_particles_uniform(occ, n_part * 8, seed=seed)
    rng = np.random.default_rng(seed)
    xs, ys = wl.free_cells(occ)
    for s in range(n_sets):
        if s % 2 == 0:
            sets[s] = base_u[rng.integers(0, len(base_u), n_part)]
            sets[s, :, 2] = rng.uniform(0, 2 * np.pi, n_part)
        else:
            k = int(rng.integers(0, len(xs)))
            sets[s, :, 0] = np.clip(ys[k] + 0.5 + rng.normal(0, 10.0, n_part), 1.0, occ.shape[1] - 2.0)
            sets[s, :, 1] = np.clip(xs[k] + 0.5 + rng.normal(0, 10.0, n_part), 1.0, occ.shape[0] - 2.0)
            sets[s, :, 2] = rng.uniform(0, 2 * np.pi) + rng.normal(0, 0.2, n_part)
    angles = wl.lidar_angles(N_BEAMS)
    obs = np.clip(120.0 + 80.0 * np.sin(np.linspace(0, 3.0, N_BEAMS)), 0, MAX_RANGE).astype(np.float32)
    return sets, angles, obs


class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the bench runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        nv = self.nv
        names = {}
        for nm in ("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap", "HwPowerBrakeSlowdown"):
            v = getattr(nv, "nvmlClocksThrottleReason" + nm, None)
            if v is not None:
                names[v] = nm
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv:
            self.t.start()

    def stop(self):
        self.stop_flag = True
        if self.nv and self.t.is_alive():
            self.t.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_reference_run(occ, sets, angles, obs, table, steps, warmup, threads, budget_s=None):
    """The reference's own CPU implementation of the step (oracle/_ref, unmodified RangeLib.h built with its
    shipped optimisation flags) or, if that library is absent, our C port.  Returns (rays/s, kind, steps run)."""
    from oracle import port, ref
    if ref.available("shipped"):
        rmap = ref.RefMap(occ=occ, flavor="shipped")
        meth = ref.RefMethod(ref.RM, rmap, MAX_RANGE, threads=threads)
        kind = "reference"
    else:
        meth = port.Oracle(port.RM, occ, MAX_RANGE, threads=threads)
        kind = "port"
    meth.set_sensor_model(table)
    for i in range(warmup):
        meth.calc_range_repeat_angles_eval_sensor_model(sets[i % len(sets)], angles, obs)
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        meth.calc_range_repeat_angles_eval_sensor_model(sets[i % len(sets)], angles, obs)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done * N_PART * N_BEAMS / dt, kind, done, dt


def reference_cuda_run(occ, sets, angles, obs, table, budget_s=5.0):
    """The reference's own CUDA path (RayMarchingGPU + includes/kernels.cu recompiled for sm_100a, oracle/_ref/
    libref_cuda.so) on the same step, the way the reference documents it for a particle filter: ranges on the GPU
    through numpy_calc_range_angles with host buffers (its only interface: cudaMemcpy in, kernel, cudaMemcpy out,
    device sync), then RangeMethod::eval_sensor_model on one host thread (its fused GPU variant prints
    "unimplemented", kernels.cu:281-284).  Returns None when the library is absent."""
    from oracle import ref
    if not ref.available("cuda"):
        return None
    rmap = ref.RefMap(occ=occ, flavor="cuda")
    meth = ref.RefMethod(ref.RMGPU, rmap, MAX_RANGE)
    meth.set_sensor_model(table)
    n_rays = N_PART * N_BEAMS

    def timed(fn):
        for i in range(3):
            fn(i)
        t0, done = time.perf_counter(), 0
        while time.perf_counter() - t0 < budget_s / 2 and done < 5000:
            fn(done)
            done += 1
        return (time.perf_counter() - t0) / done, done

    ranges = [None]

    def cast(i):
        ranges[0] = meth.numpy_calc_range_angles(sets[i % len(sets)], angles)

    def step(i):
        cast(i)
        meth.eval_sensor_model(obs, ranges[0], N_BEAMS, N_PART)

    t_cast, n1 = timed(cast)
    t_step, n2 = timed(step)
    return {"ranges_only_rays_per_s": n_rays / t_cast, "ranges_only_ms": t_cast * 1e3,
            "value": n_rays / t_step, "ms_per_step": t_step * 1e3, "unit": "rays/s", "steps": n2,
            "what": "reference RayMarchingGPU.numpy_calc_range_angles (kernels.cu recompiled for sm_100a, CHUNK_SIZE 262144, "
                    "NUM_THREADS 256, host buffers) + RangeMethod::eval_sensor_model on 1 host thread; compare with e2e"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    from range_libc_b200 import workloads as wl
    occ = wl.load_map(MAP)
    sets, angles, obs = make_inputs(occ, 16)
    table = wl.sensor_table(K_TABLE)
    threads = os.cpu_count() or 1
    steps = min(args.steps, 400)
    v, kind, done, dt = cpu_reference_run(occ, sets, angles, obs, table, steps, min(args.warmup, 5), threads, budget_s=120.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "rays/s", "n_gpus": args.gpus, "steps": done,
        "warmup": min(args.warmup, 5), "ms_per_step": dt / done * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 ranges, f64 weights", "data": "synthetic",
        "config": {"workload": "PF sensor update %dx%d RM fused, %s, max_range %g, K=%d" % (
            N_PART, N_BEAMS, MAP, MAX_RANGE, K_TABLE)},
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": threads, "kind": kind,
                         "sample": "%d full steps of the same workload, %d host threads slicing particles around the "
                                   "reference's own single-threaded loop" % (done, threads)},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary throughput lines")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import range_libc_b200 as rl
    from range_libc_b200 import workloads as wl

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: rangelib_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W_ = max(args.warmup, 3)
    K_ = max(args.steps, 1)

    occ = wl.load_map(MAP)
    omap = rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
    rm = rl.PyRayMarchingGPU(omap, MAX_RANGE, device=local_rank)
    table = wl.sensor_table(K_TABLE)
    rm.set_sensor_model(table)
    stream = torch.cuda.current_stream()
    rm.set_stream(stream.cuda_stream)

    # inputs: rotate through particle sets whose total size exceeds L2, so no step finds its inputs cached
    n_sets = L2_BYTES // (N_PART * 12) + 64
    sets_h, angles_h, obs_h = make_inputs(occ, n_sets, seed=2026 + rank)
    sets = torch.from_numpy(sets_h).to(dev)
    angles = torch.from_numpy(angles_h).to(dev)
    obs = torch.from_numpy(obs_h).to(dev)
    weights_all = torch.empty(world * N_PART, dtype=torch.float64, device=dev)
    my_w = weights_all[rank * N_PART:(rank + 1) * N_PART]

    set_views = [sets[i] for i in range(n_sets)]

    # multi-GPU: the fused kernel stores its weights into every rank's gathered array over NVLink (symmetric
    # memory) and a symmetric-memory barrier closes the step; NCCL all-gather is the fallback
    peer, gather_mode = None, "none (single GPU)"
    if world > 1:
        gather_mode = "nccl all_gather_into_tensor"
        # "peer" (fused peer stores + symmetric-memory barrier) measured 31.1 us/step at N=2 against 34.7 us for
        # "signal" (one kernel per step with in-kernel epoch flags: every CTA pays a system-scope fence) when the
        # steps are replayed from a CUDA graph; launched eagerly "signal" is the faster one (33.7 vs 36.0 us).
        want = os.environ.get("RL_BENCH_GATHER", "pipelined")
        if want in ("peer", "signal", "pipelined"):
            try:
                from range_libc_b200 import parallel
                if want == "pipelined":
                    peer = parallel.PipelinedPeerStoreUpdate(world * N_PART, rm, angles, obs, device=dev)
                    gather_mode = ("fused kernel epilogue: peer stores over NVLink into double-buffered symmetric memory; the "
                                   "symmetric-memory barrier closing step k runs on a side stream and overlaps the compute "
                                   "of step k+1")
                elif want == "signal":
                    peer = parallel.SignalledSensorUpdate(world * N_PART, rm, angles, obs, device=dev)
                    gather_mode = ("one kernel per step: fused compute + peer stores over NVLink (symmetric memory, double "
                                   "buffered) + in-kernel epoch flags; no barrier launch, no NCCL call")
                else:
                    peer = parallel.PeerStoreSensorUpdate(world * N_PART, rm, angles, obs, device=dev)
                    gather_mode = "fused kernel epilogue: peer stores over NVLink into symmetric memory + symm barrier"
                peer.update(set_views[0])
                torch.cuda.synchronize()
            except Exception as ex:  # noqa: BLE001
                peer = None
                gather_mode = "nccl all_gather_into_tensor (symmetric memory unavailable: %s)" % str(ex).splitlines()[0][:100]
    signalled = peer is not None and hasattr(peer, "flags")
    pipelined = peer is not None and hasattr(peer, "finish")

    def step(i):
        if peer is not None:
            if signalled:  # the next step's kernel waits in-kernel for this step's gather; see finish()
                peer.update(set_views[i % n_sets], wait=False)
            else:
                peer.update(set_views[i % n_sets])
            return
        rm.calc_range_repeat_angles_eval_sensor_model(set_views[i % n_sets], angles, obs, my_w)
        if world > 1:
            dist.all_gather_into_tensor(weights_all, my_w)

    def finish():
        """close the last step: every rank's slice of the last gather has arrived (signalled mode)"""
        if signalled:
            rm.peers_wait()
        if pipelined:
            peer.finish()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(W_):
        step(i)
    finish()
    barrier()

    # The K timed steps are captured once into a CUDA graph (launch-bound inner loop: each step is a
    # ~10 us kernel) and replayed inside the event-bracketed region; eager launches are the fallback.
    graph, mode = None, "eager"
    try:
        if world > 1 and peer is None:
            raise RuntimeError("NCCL collectives are launched eagerly (no graph capture)")
        if pipelined:
            peer.reset()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(stream)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side, capture_error_mode="thread_local"):
            rm.set_stream(torch.cuda.current_stream().cuda_stream)
            for i in range(K_):
                step(W_ + i)
            finish()
        rm.set_stream(stream.cuda_stream)
        if pipelined:
            peer.reset()
        graph.replay()  # untimed: instantiation / first-run costs
        barrier()
        mode = "cuda_graph"
    except Exception as ex:  # noqa: BLE001
        graph = None
        rm.set_stream(stream.cuda_stream)
        torch.cuda.synchronize()
        if pipelined:
            peer.reset()
        mode = "eager (graph capture failed: %s)" % str(ex).splitlines()[0][:120]

    l0 = rl.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    if graph is not None:
        graph.replay()
    else:
        for i in range(K_):
            step(W_ + i)
        finish()
    e1.record(stream)
    barrier()
    launches = (rl.kernel_launches() - l0) if graph is None else K_
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    rays_per_step = N_PART * N_BEAMS * world
    value = rays_per_step * K_ / (ms * 1e-3)

    # the same K steps launched eagerly through the Python API (host launch overhead included)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(K_):
        step(W_ + i)
    finish()
    e1.record(stream)
    barrier()
    eager_ms = e0.elapsed_time(e1) / K_

    # average duration of the dominant kernel (fused_kernel<RM>): inside the replayed graph the K kernel nodes
    # run back to back on the launch stream, so the event-bracketed region divided by K is the per-launch
    # device time; in eager mode fall back to one event pair per launch
    if graph is not None and world == 1:
        kernel_ms = ms / K_
    else:
        kt = []
        for i in range(min(K_, 200)):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            rm.calc_range_repeat_angles_eval_sensor_model(set_views[(W_ + i) % n_sets], angles, obs, my_w)
            b.record(stream)
            b.synchronize()
            kt.append(a.elapsed_time(b))
        kernel_ms = float(np.mean(kt))

    # cold-L2 variant: flush L2 (write a 256 MB buffer) before every step, time steps individually
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    cold = []
    for i in range(min(K_, 50)):
        flush.fill_(i & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        rm.calc_range_repeat_angles_eval_sensor_model(sets[(W_ + i) % n_sets], angles, obs, my_w)
        b.record(stream)
        b.synchronize()
        cold.append(a.elapsed_time(b))
    cold_ms = float(np.mean(cold))
    del flush

    # end to end through the public API with HOST buffers: pinned inputs, H2D + kernel + D2H inside the timed region.
    # The public API is the drop-in Cython module `range_libc` (same names as the reference's module); the ctypes
    # mirror is used if the extension has not been built.
    rm.set_stream(None)
    e2e_api = "range_libc_b200.PyRayMarchingGPU.calc_range_repeat_angles_eval_sensor_model(numpy host arrays)"
    rm_e2e = rm
    try:
        sys.path.insert(0, os.path.join(ROOT, "range_libc_b200", "pywrapper"))
        import range_libc as cy
        cy_map = cy.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
        rm_e2e = cy.PyRayMarchingGPU(cy_map, MAX_RANGE)
        rm_e2e.set_sensor_model(table)
        e2e_api = "range_libc.PyRayMarchingGPU.calc_range_repeat_angles_eval_sensor_model(numpy host arrays) [Cython drop-in]"
    except ImportError:
        pass
    n_host_sets = 64
    host_sets = [torch.from_numpy(sets_h[i].copy()).pin_memory().numpy() for i in range(n_host_sets)]
    host_angles = torch.from_numpy(angles_h.copy()).pin_memory().numpy()
    host_obs = torch.from_numpy(obs_h.copy()).pin_memory().numpy()
    host_w = torch.empty(N_PART, dtype=torch.float64).pin_memory().numpy()
    for i in range(W_):
        rm_e2e.calc_range_repeat_angles_eval_sensor_model(host_sets[i % n_host_sets], host_angles, host_obs, host_w)
    barrier()
    ke = min(K_, 2000)
    t0 = time.perf_counter()
    for i in range(ke):
        rm_e2e.calc_range_repeat_angles_eval_sensor_model(host_sets[i % n_host_sets], host_angles, host_obs, host_w)
        if world > 1:
            my_w.copy_(torch.from_numpy(host_w), non_blocking=True)
            dist.all_gather_into_tensor(weights_all, my_w)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = rays_per_step * ke / e2e_s
    clocks = sampler.stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    traffic, l2_bytes = profiled_traffic()
    algo_bytes = 12 * N_PART + 8 * N_BEAMS + 8 * N_PART  # (12 N + 8 M + 8 N) per launch
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": K_, "warmup": W_,
        "ms_per_step": ms / K_, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 ranges, f64 weights", "data": "synthetic",
        "config": {"workload": "PF sensor update %dx%d RM fused (calc_range_repeat_angles_eval_sensor_model), %s 1200x1200, "
                               "max_range %g px, K=%d table%s" % (N_PART, N_BEAMS, MAP, MAX_RANGE, K_TABLE,
                                                                  ", per rank + all-gather of weights" if world > 1 else ""),
                   "particles": "alternating global-init (uniform over free cells) and tracking (sigma 10 px / 0.2 rad) clouds",
                   "l2": "inputs rotate through %d particle sets (%.0f MB > 126 MB L2); the map structures (5.76 MB distance "
                         "transform, 2 MB table) stay L2-resident as they do in deployment; see value_cold_l2 for a flushed-L2 "
                         "measurement" % (n_sets, n_sets * N_PART * 12 / 1e6)},
        "gpu_launches": int(launches),
        "launch_mode": mode,
        "weight_gather": gather_mode,
        "ms_per_step_eager": eager_ms,
        "kernel_ms": kernel_ms,
        "value_cold_l2": N_PART * N_BEAMS / (cold_ms * 1e-3),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": 12 * N_PART + 8 * N_BEAMS,
                "d2h_bytes_per_step": 8 * N_PART, "ms_per_step": e2e_s / ke * 1e3, "steps": ke,
                "api": e2e_api},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "fused_kernel<RM>", "algorithmic_bytes_per_launch": algo_bytes,
                     "peak_source": peak_src, "l2_sector_bytes_per_launch": l2_bytes,
                     "l2_gather": None if not l2_bytes else {
                         "achieved_gsectors_per_s": l2_bytes / 32.0 / (kernel_ms * 1e-3) / 1e9,
                         "peak_gsectors_per_s": 291.0,
                         "frac": l2_bytes / 32.0 / (kernel_ms * 1e-3) / 291e9,
                         "what": "32-byte sectors through L2 per launch (ncu lts__t_sectors, committed capture) over the "
                                 "measured kernel time, against the chip's random-gather rate: one L1-miss request per "
                                 "clock per SM = 148 x 1.965 GHz, which tools/gather_bench.cu reproduces (287-292 G/s). "
                                 "The large-batch RM kernel runs at ~90 % of it (profiles/ncu_cast_r01.txt); this 240k-ray "
                                 "launch fits the chip once and is bound by launch + set-up + its longest dependent chain"},
                     "note": "0.335 algorithmic B/ray: this path is bound by the latency of the longest sphere-tracing "
                             "chain in the launch (dependent L2 reads, ~143 ns each on B200) and by L2 sector traffic, not by "
                             "HBM (DESIGN.md section 4); traffic = DRAM bytes per launch from the committed ncu capture "
                             "(cold cache: mostly the first touch of the 5.76 MB distance transform)"},
    }
    if world == 1:
        threads = os.cpu_count() or 1
        v, kind, done, dt = cpu_reference_run(occ, sets_h, angles_h, obs_h, table, 100000, 3, threads, budget_s=10.0)
        line["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": threads, "kind": kind,
                                "sample": "%d full steps (%.1f s) of the same 4000x60 workload on %d host threads" % (
                                    done, dt, threads)}
        v1, kind1, done1, dt1 = cpu_reference_run(occ, sets_h, angles_h, obs_h, table, 100000, 1, 1, budget_s=5.0)
        line["cpu_baseline"]["single_thread_value"] = v1
        try:
            line["reference_cuda_baseline"] = reference_cuda_run(occ, sets_h, angles_h, obs_h, table)
        except Exception as ex:  # noqa: BLE001
            line["reference_cuda_baseline"] = {"unavailable": str(ex).splitlines()[0][:120]}
        if not args.no_extra:
            line["extra"] = extra_throughput(rl, wl, occ, omap, dev, stream, peak)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _time_launches(fn, stream, iters=10, warm=3):
    import torch
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)) * 1e-3


def extra_throughput(rl, wl, occ, omap, dev, stream, peak):
    """Secondary numbers (not the judged line), one per BASELINE config, all device resident:
    large-batch ray casts/s per method (16 algorithmic B/ray; north_star's 50 G rays/s RM target), CDDT/PCDDT
    build + query on gigantic_map (C3), BL on a dynamic 4096^2 grid (C4), fused RM + sensor model for
    1M particles x 1080 beams on an 8192^2 grid (C5, one GPU's worth)."""
    import torch
    out = {}
    W, H = occ.shape
    N = 1 << 24
    q = torch.from_numpy(wl.random_queries(W, H, N, seed=1)).to(dev)
    r = torch.empty(N, dtype=torch.float32, device=dev)

    def pcddt():
        m = rl.PyCDDTCast(omap, MAX_RANGE, 108)
        m.prune()
        return m

    for nm, ctor in (("rm", lambda: rl.PyRayMarchingGPU(omap, MAX_RANGE)), ("cddt", lambda: rl.PyCDDTCast(omap, MAX_RANGE, 108)),
                     ("pcddt", pcddt), ("bl", lambda: rl.PyBresenhamsLine(omap, MAX_RANGE))):
        try:
            m = ctor()
            m.set_stream(stream.cuda_stream)
            n = N if nm != "bl" else N // 4
            t = _time_launches(lambda: m.calc_range_many_grid(q[:n], r[:n]), stream)
            out[nm + "_random_rays_per_s"] = n / t
            out[nm + "_random_hbm_frac"] = 16.0 * n / t / 1e9 / peak
            del m
        except Exception as ex:  # noqa: BLE001
            out[nm + "_error"] = str(ex)[:200]
    out["workload"] = "2^24 uniformly random grid-coordinate queries on %s (BL: 2^22), max_range 500, device resident" % MAP
    del q, r
    # lidar-scan batches through calc_range_repeat_angles (numpy_calc_range_angles): the beams of one pose are
    # neighbours in a warp, so their distance-map reads share 128-byte lines -- unlike the uniformly random rays
    # above, whose rate is capped by L1 tag lookups (one line per clock per SM, 290 G lines/s chip-wide)
    try:
        rm = rl.PyRayMarchingGPU(omap, MAX_RANGE)
        rm.set_stream(stream.cuda_stream)
        scans = {}
        dt_host = rm.distance_transform()
        for label, n_p, n_b, maker in (("uniform_262144x60", 262144, 60, wl.pf_particles_uniform),
                                       ("tracking_262144x60", 262144, 60, lambda o, n, seed: wl.pf_particles_tracking(o, n, seed=seed, dt=dt_host)[0]),
                                       ("uniform_16384x1080", 16384, 1080, wl.pf_particles_uniform)):
            parts = torch.from_numpy(maker(occ, n_p, seed=11)).to(dev)
            ang = torch.from_numpy(wl.lidar_angles(n_b)).to(dev)
            ranges = torch.empty(n_p * n_b, dtype=torch.float32, device=dev)
            t = _time_launches(lambda: rm.calc_range_repeat_angles(parts, ang, ranges), stream)
            scans[label] = {"rays_per_s": n_p * n_b / t, "hbm_frac": (12.0 * n_p + 4.0 * n_b + 4.0 * n_p * n_b) / t / 1e9 / peak}
            del parts, ang, ranges
        out["rm_scan_batches"] = scans
        del rm
    except Exception as ex:  # noqa: BLE001
        out["rm_scan_error"] = str(ex)[:200]
    # C3: CDDT / PCDDT on gigantic_map (10976^2), theta_discretization 108
    try:
        big = wl.load_map("gigantic_map")
        bmap = rl.PyOMap(np.ascontiguousarray(big.T.astype(bool)))
        t0 = time.perf_counter()
        cd = rl.PyCDDTCast(bmap, MAX_RANGE, 108)
        t_build = time.perf_counter() - t0
        cd.set_stream(stream.cuda_stream)
        n = 1 << 24
        qb = torch.from_numpy(wl.random_queries(big.shape[0], big.shape[1], n, seed=2)).to(dev)
        rb = torch.empty(n, dtype=torch.float32, device=dev)
        t_q = _time_launches(lambda: cd.calc_range_many_grid(qb, rb), stream, iters=5)
        t0 = time.perf_counter()
        cd.prune()
        t_prune = time.perf_counter() - t0
        t_qp = _time_launches(lambda: cd.calc_range_many_grid(qb, rb), stream, iters=5)
        out["c3_gigantic_map"] = {"cddt_build_s": t_build, "pcddt_prune_s": t_prune, "cddt_rays_per_s": n / t_q,
                                  "pcddt_rays_per_s": n / t_qp, "table_bytes_after_prune": cd.memory(),
                                  "note": "build/prune include upload of the 120 MB grid and all device work; CPU reference: "
                                          "CDDT build ~15 s, prune ~24 min (SURVEY.md section 6)"}
        del cd, qb, rb, bmap, big
    except Exception as ex:  # noqa: BLE001
        out["c3_error"] = str(ex)[:200]
    # C4: Bresenham on a dynamic 4096^2 grid: per frame 64 16x16 occupancy patches + 2^20 rays
    try:
        occ4 = wl.synthetic_map(4096, seed=2026)
        m4 = rl.PyOMap(np.ascontiguousarray(occ4.T.astype(bool)))
        bl = rl.PyBresenhamsLine(m4, MAX_RANGE)
        bl.set_stream(stream.cuda_stream)
        n = 1 << 20
        q4 = torch.from_numpy(wl.random_queries(4096, 4096, n, seed=3)).to(dev)
        r4 = torch.empty(n, dtype=torch.float32, device=dev)
        frames = []
        for f in range(8):
            blocks = wl.flip_blocks(occ4, f, seed=2026)
            rects = np.array([[x0, y0, p.shape[0], p.shape[1]] for x0, y0, p in blocks], np.int32)
            frames.append((torch.from_numpy(np.concatenate([p.ravel() for _, _, p in blocks])).to(dev), rects))

        def frame(f):
            patches, rects = frames[f % 8]
            bl.update_map_batch(patches, rects)  # all 64 patches of the frame in one launch
            bl.calc_range_many_grid(q4, r4)

        for f in range(3):
            frame(f)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for f in range(16):
            frame(f)
        b.record(stream)
        b.synchronize()
        t = a.elapsed_time(b) * 1e-3 / 16
        out["c4_dynamic_bl_4096"] = {"ms_per_frame": t * 1e3, "rays_per_s": n / t, "patches_per_frame": 64,
                                     "note": "update + query per frame; the CPU reference copies the whole map per BresenhamsLine"}
        del bl, q4, r4, m4
    except Exception as ex:  # noqa: BLE001
        out["c4_error"] = str(ex)[:200]
    # C5: fused RM + sensor model, 1M particles x 1080 beams on a synthetic 8192^2 grid (single GPU share of the config)
    try:
        occ5 = wl.synthetic_map(8192, seed=2026)
        m5 = rl.PyOMap(np.ascontiguousarray(occ5.T.astype(bool)))
        t0 = time.perf_counter()
        rm5 = rl.PyRayMarchingGPU(m5, MAX_RANGE)
        t_dt = time.perf_counter() - t0
        rm5.set_sensor_model(wl.sensor_table(K_TABLE))
        rm5.set_stream(stream.cuda_stream)
        n5, m_beams = 1_000_000, 1080
        p5 = torch.from_numpy(wl.pf_particles_uniform(occ5, n5, seed=4)).to(dev)
        a5 = torch.from_numpy(wl.lidar_angles(m_beams)).to(dev)
        o5 = torch.from_numpy(np.clip(150 + 100 * np.sin(np.linspace(0, 6, m_beams)), 0, 500).astype(np.float32)).to(dev)
        w5 = torch.empty(n5, dtype=torch.float64, device=dev)
        t = _time_launches(lambda: rm5.calc_range_repeat_angles_eval_sensor_model(p5, a5, o5, w5), stream, iters=3, warm=1)
        out["c5_rm_fused_8192"] = {"particles": n5, "beams": m_beams, "s_per_update": t, "rays_per_s": n5 * m_beams / t,
                                   "dt_build_s": t_dt, "note": "268 MB distance transform (> L2): sector traffic is HBM traffic here"}
        del rm5, p5, w5, m5
    except Exception as ex:  # noqa: BLE001
        out["c5_error"] = str(ex)[:200]
    return out


if __name__ == "__main__":
    main()
