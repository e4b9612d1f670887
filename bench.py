#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's configurations.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload auto|c2|c5] [--no-extra]

A step = one particle-filter sensor update: RM ray casts for N x M (particle, beam) pairs fused with the
sensor-table lookup and the per-particle product
(RangeMethod::calc_range_repeat_angles_eval_sensor_model, /root/reference/includes/RangeLib.h:558-612).
Metric: ray casts / s (one sensor-model evaluation per ray, so it is also evals / s).

Workloads (--workload auto picks by the number of GPUs):
  c2  (N = 1)  BASELINE configs[1]: 4000 particles x 60 beams on basement_hallways_5cm, one B200 -- the configuration
               the metric is quoted on.  The other configurations are reported in `extra`, each with its roofline
               fractions and a CPU baseline timed in the same run.
  c5  (N > 1)  BASELINE configs[4]: 1 000 000 particles x 1080 beams on a synthetic 8192^2 grid, STRONG scaling: the
               particles are sharded over the ranks (one process per GPU, map / distance transform / table replicated),
               every step ends with all ranks holding all weights (a deep update is a lane-re-queuing cast into a
               scratch array followed by the evaluation kernel, whose epilogue stores each weight into every rank's
               gathered array over NVLink and signals completion; no NCCL call, no cross-step pipelining).
               Rank 0 also times the same update on its GPU alone (`strong_scaling_base`), and every rank checks the
               gathered weights against a single-GPU recomputation of all shards (`gather_verified`).
`--impl reference` times the reference's own CPU implementation (oracle/_ref, unmodified RangeLib.h) on the same
configuration, on all host threads, on a bounded sample of the workload.
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
C5_SIZE, C5_PART, C5_BEAMS, C5_SEED = 8192, 1_000_000, 1080, 2026
METRIC = "ray casts/sec (RM, fused PF sensor-model update: calc_range_repeat_angles + eval_sensor_model)"
DTYPE = "f32 ranges, f64 weights"
L2_BYTES = 126 * 1024 * 1024


def workload_config(name):
    """The `config` dict of the JSON line: identical in both arms (ours / reference)."""
    if name == "c2":
        return {"workload": "C2 (BASELINE configs[1]): PF sensor update %dx%d RM fused "
                            "(calc_range_repeat_angles_eval_sensor_model), %s 1200x1200, max_range %g px, K=%d table" % (
                                N_PART, N_BEAMS, MAP, MAX_RANGE, K_TABLE),
                "particles": "alternating global-init (uniform over free cells) and tracking (sigma 10 px / 0.2 rad) clouds",
                "l2": "GPU arm: inputs larger than L2 -- the particle sets rotate through > 126 MB, and warm-up, the untimed "
                      "graph replay and the timed steps use disjoint sets, so no timed step finds its particles cached; the map "
                      "structures (5.76 MB distance transform, 2 MB table) stay L2-resident, as in deployment (value_cold_l2: "
                      "L2 flushed before every step)"}
    return {"workload": "C5 (BASELINE configs[4]): PF sensor update %dx%d RM fused "
                        "(calc_range_repeat_angles_eval_sensor_model), synthetic %dx%d grid (seed %d), max_range %g px, "
                        "K=%d table, particles sharded over the GPUs, all ranks end each step with all weights" % (
                            C5_PART, C5_BEAMS, C5_SIZE, C5_SIZE, C5_SEED, MAX_RANGE, K_TABLE),
            "particles": "uniform over free cells, theta uniform (global localisation)",
            "l2": "GPU arm: the 268 MB distance transform exceeds the 126 MB L2 (inputs larger than L2); the 12 MB of poses are "
                  "device resident for `value` and cross PCIe every step for `e2e`",
            "parallelism": "particles sharded over the ranks in contiguous slices, map / distance transform / table replicated"}


def profiled_update(name, kernel_substrs):
    """DRAM bytes / L2 sector bytes / seconds of one UPDATE made of several kernels: the per-kernel means of a committed
    warm capture (profiled), summed.  None when any of them is absent."""
    parts = [profiled(name, k) for k in kernel_substrs]
    if any(p is None for p in parts):
        return None
    return {"dram_bytes": sum(p["dram_bytes"] for p in parts), "l2_sector_bytes": sum(p["l2_sector_bytes"] for p in parts),
            "seconds": sum(p["seconds"] for p in parts), "launches": sum(p["launches"] for p in parts),
            "per_kernel_seconds": {k: p["seconds"] for k, p in zip(kernel_substrs, parts)},
            "per_kernel_issue_slots_busy": {k: p["issue_slots_busy"] for k, p in zip(kernel_substrs, parts)},
            "source": parts[0]["source"]}


C5_KERNELS = ("rm_persist_kernel", "eval_overlap_kernel")  # the two kernels of a deep fused RM update


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled(name, kernel_substr=None):
    """dram bytes, L2 sectors and duration per launch from a committed WARM ncu --set full capture
    (profiles/r02/ncu_<name>.txt, `--cache-control none`, made by tools/gpu_profile_r02.sh).  None when absent."""
    p = os.path.join(ROOT, "profiles", "r02", "ncu_%s.txt" % name)
    if not os.path.exists(p):
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
    launches, cur = [], None
    for line in open(p):
        f = line.split()
        if line.startswith("## "):
            cur = {"kernel": line[3:].strip(), "dram": 0.0}
            launches.append(cur)
        elif cur is not None and len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            cur["dram"] += float(f[1].replace(",", "")) * scale.get(f[2], 1.0)
        elif cur is not None and len(f) >= 2 and f[0] == "lts__t_sectors.sum":
            cur["sectors"] = float(f[1].replace(",", ""))
        elif cur is not None and len(f) >= 3 and f[0] == "gpu__time_duration.sum":
            cur["seconds"] = float(f[1].replace(",", "")) * tscale.get(f[2], 1.0)
        elif cur is not None and len(f) >= 2 and f[0] == "smsp__issue_active.avg.pct_of_peak_sustained_active":
            cur["issue"] = float(f[1].replace(",", "")) / 100.0
    launches = [l for l in launches if kernel_substr is None or kernel_substr in l["kernel"]]
    if not launches:
        return None
    return {"dram_bytes": float(np.mean([l["dram"] for l in launches])),
            "l2_sector_bytes": float(np.mean([l.get("sectors", 0.0) for l in launches])) * 32.0,
            "seconds": float(np.mean([l.get("seconds", 0.0) for l in launches])), "launches": len(launches),
            "issue_slots_busy": float(np.mean([l.get("issue", 0.0) for l in launches])),
            "source": "profiles/r02/ncu_%s.txt (ncu --set full --cache-control none, warm)" % name}


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


def c5_inputs(n_part=C5_PART):
    """Map, particles, beam angles and observation of config 5 (the same arrays in both arms and on every rank)."""
    from range_libc_b200 import workloads as wl
    occ = wl.synthetic_map(C5_SIZE, seed=C5_SEED)
    parts = wl.pf_particles_uniform(occ, n_part, seed=4)
    angles = wl.lidar_angles(C5_BEAMS)
    obs = np.clip(150 + 100 * np.sin(np.linspace(0, 6, C5_BEAMS)), 0, 500).astype(np.float32)
    return occ, parts, angles, obs


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


# ------------------------------------------------------------------------------------------------------------------
# the reference's own CPU implementation (test infrastructure; never on the measured GPU path)
# ------------------------------------------------------------------------------------------------------------------
def cpu_method(kind_name, occ, threads, theta_disc=108, pruned=False):
    """A reference range method on the host: oracle/_ref (unmodified RangeLib.h, shipped flags) or, if that library is
    absent, our C port.  Returns (method, kind)."""
    from oracle import port, ref
    if ref.available("shipped"):
        k = {"rm": ref.RM, "bl": ref.BL, "cddt": ref.CDDT}[kind_name]
        meth = ref.RefMethod(k, ref.RefMap(occ=occ, flavor="shipped"), MAX_RANGE, theta_disc, threads=threads)
        kind = "reference"
    else:
        k = {"rm": port.RM, "bl": port.BL, "cddt": port.CDDT}[kind_name]
        meth = port.Oracle(k, occ, MAX_RANGE, theta_disc, threads=threads)
        kind = "port"
    if pruned:
        meth.prune(MAX_RANGE)
    return meth, kind


def cpu_fused_run(occ, sets, angles, obs, table, steps, warmup, threads, budget_s=None):
    """Times `steps` fused updates (particle sets rotate) on the host.  Returns (rays/s, kind, steps run, seconds)."""
    meth, kind = cpu_method("rm", occ, threads)
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
    return done * sets[0].shape[0] * len(angles) / dt, kind, done, dt


def cpu_cast_rate(kind_name, occ, q, threads, pruned=False, budget_s=4.0):
    """rays/s of the host method on (a prefix of) the grid-coordinate batch q, with its construction time."""
    t0 = time.perf_counter()
    meth, kind = cpu_method(kind_name, occ, threads, pruned=pruned)
    t_build = time.perf_counter() - t0
    n = min(len(q), 20000)
    t0 = time.perf_counter()
    meth.calc_range_many(q[:n])
    rate = n / (time.perf_counter() - t0)
    n = int(min(len(q), max(n, rate * budget_s)))
    t0 = time.perf_counter()
    meth.calc_range_many(q[:n])
    dt = time.perf_counter() - t0
    return {"rays_per_s": n / dt, "rays": n, "threads": threads, "kind": kind, "construction_s": t_build}


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


def run_reference(args, rank, world, workload):
    """--impl reference: the reference's CPU implementation of the step on all host threads (rank 0 only)."""
    if rank != 0:
        return
    from range_libc_b200 import workloads as wl
    table = wl.sensor_table(K_TABLE)
    threads = os.cpu_count() or 1
    warm = min(args.warmup, 3)
    if workload == "c2":
        occ = wl.load_map(MAP)
        sets, angles, obs = make_inputs(occ, 16)
        steps = min(args.steps, 400)
        v, kind, done, dt = cpu_fused_run(occ, sets, angles, obs, table, steps, warm, threads, budget_s=120.0)
        sample = ("%d full steps of the same 4000x60 workload (16 rotating particle sets), %d host threads slicing particles "
                  "around the reference's own single-threaded loop" % (done, threads))
        scaling = "weak"
    else:
        occ, parts, angles, obs = c5_inputs(50000 * 4)
        sets = [np.ascontiguousarray(parts[i * 50000:(i + 1) * 50000]) for i in range(4)]
        steps = min(args.steps, 40)
        v, kind, done, dt = cpu_fused_run(occ, sets, angles, obs, table, steps, min(warm, 1), threads, budget_s=150.0)
        sample = ("%d steps, each a 50 000-particle x 1080-beam sample (5 %% of one update of the 10^6-particle cloud, same "
                  "map / beams / table), %d host threads slicing particles around the reference's single-threaded loop; "
                  "rays/s of the sample = rays/s of the workload (particles are independent)" % (done, threads))
        scaling = "strong"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "rays/s", "n_gpus": args.gpus, "steps": done,
        "warmup": warm, "ms_per_step": dt / done * 1e3, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": workload_config(workload),
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# ours
# ------------------------------------------------------------------------------------------------------------------
def cython_module():
    """The drop-in `range_libc` extension (same names as the reference's module); None if it has not been built."""
    try:
        p = os.path.join(ROOT, "range_libc_b200", "pywrapper")
        if p not in sys.path:
            sys.path.insert(0, p)
        import range_libc as cy
        return cy
    except ImportError:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "c2", "c5"])
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary per-configuration blocks")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = args.workload if args.workload != "auto" else ("c2" if max(world, args.gpus) == 1 else "c5")
    if args.impl == "reference":
        run_reference(args, rank, world, workload)
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: rangelib_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if workload == "c2":
        if world > 1:
            raise SystemExit("workload c2 is the single-GPU configuration; use --workload c5 (or auto) with several GPUs")
        run_c2(args, local_rank)
    else:
        run_c5(args, rank, local_rank, world)


def run_c2(args, local_rank):
    import torch
    import range_libc_b200 as rl
    from range_libc_b200 import workloads as wl

    dev = torch.device("cuda", local_rank)
    W_ = max(args.warmup, 3)
    K_ = max(args.steps, 1)
    occ = wl.load_map(MAP)
    omap = rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
    rm = rl.PyRayMarchingGPU(omap, MAX_RANGE, device=local_rank)
    table = wl.sensor_table(K_TABLE)
    rm.set_sensor_model(table)
    stream = torch.cuda.current_stream()
    rm.set_stream(stream.cuda_stream)

    # inputs: particle sets whose total size exceeds L2; warm-up, the untimed graph replay and the timed steps use
    # disjoint sets, and L2 is flushed before the timed region, so no timed step finds its particles cached
    n_sets = max(L2_BYTES // (N_PART * 12) + 64, W_ + 2 * K_ + 64)
    sets_h, angles_h, obs_h = make_inputs(occ, min(n_sets, 4096), seed=2026)
    n_sets = len(sets_h)
    sets = torch.from_numpy(sets_h).to(dev)
    angles = torch.from_numpy(angles_h).to(dev)
    obs = torch.from_numpy(obs_h).to(dev)
    w = torch.empty(N_PART, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step(i):
        rm.calc_range_repeat_angles_eval_sensor_model(sets[i % n_sets], angles, obs, w)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(W_):
        step(i)
    torch.cuda.synchronize()

    # The K timed steps are captured once into a CUDA graph (launch-bound inner loop: each step is a ~20 us kernel)
    # and replayed inside the event-bracketed region; eager launches are the fallback.
    def capture(first):
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(stream)
        g = torch.cuda.CUDAGraph()
        l0 = rl.kernel_launches()
        with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
            rm.set_stream(torch.cuda.current_stream().cuda_stream)
            for i in range(K_):
                step(first + i)
        rm.set_stream(stream.cuda_stream)
        return g, rl.kernel_launches() - l0

    graph, mode, launches = None, "eager", None
    try:
        warm_graph, _ = capture(W_ + K_)     # same shape, other particle sets: pays instantiation / first-run costs
        graph, launches = capture(W_)
        warm_graph.replay()
        torch.cuda.synchronize()
        mode = "cuda_graph"
    except Exception as ex:  # noqa: BLE001
        graph = None
        rm.set_stream(stream.cuda_stream)
        torch.cuda.synchronize()
        mode = "eager (graph capture failed: %s)" % str(ex).splitlines()[0][:120]

    # the timed steps' particle sets have been touched by nothing since their upload (warm-up and the warm graph used
    # other sets; the upload itself streamed > L2 bytes through the cache afterwards), the map structures are L2-warm
    torch.cuda.synchronize()
    l0 = rl.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # K = 20 steps of ~23 us are over before the host has finished enqueueing them: a ~0.5 ms spin kernel ahead of the
    # first event keeps the stream busy while the events and the graph are enqueued, so the bracket holds the K steps
    # and not the host's launch call (device time, as the timing rules ask; `timing` in the line says so).  (0.1 ms was
    # not enough on a box whose host needed 37 us per eager launch.)
    torch.cuda._sleep(1_000_000)
    e0.record(stream)
    if graph is not None:
        graph.replay()
    else:
        for i in range(K_):
            step(W_ + i)
    e1.record(stream)
    torch.cuda.synchronize()
    if graph is None:
        launches = rl.kernel_launches() - l0
    ms = e0.elapsed_time(e1)
    value = N_PART * N_BEAMS * K_ / (ms * 1e-3)
    kernel_ms = ms / K_ if launches == K_ else None  # one kernel per step: the bracket divided by K is its duration

    # the same K steps launched eagerly through the Python API (host launch overhead included)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(K_):
        step(W_ + 2 * K_ + i)
    e1.record(stream)
    torch.cuda.synchronize()
    eager_ms = e0.elapsed_time(e1) / K_
    if kernel_ms is None:
        kernel_ms = eager_ms

    # cold-L2 variant: flush L2 before every step, time steps individually
    cold = []
    for i in range(min(K_, 50)):
        flush.fill_(i & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        step(W_ + i)
        b.record(stream)
        b.synchronize()
        cold.append(a.elapsed_time(b))
    cold_ms = float(np.mean(cold))
    del flush

    # end to end through the public API with HOST buffers: pinned inputs, H2D + kernel + D2H inside the timed region.
    # The public API is the drop-in Cython module `range_libc` (same names as the reference's module); the ctypes
    # mirror is used if the extension has not been built.
    rm.set_stream(None)
    cy = cython_module()
    if cy is not None:
        rm_e2e = cy.PyRayMarchingGPU(cy.PyOMap(np.ascontiguousarray(occ.T.astype(bool))), MAX_RANGE)
        rm_e2e.set_sensor_model(table)
        e2e_api = "range_libc.PyRayMarchingGPU.calc_range_repeat_angles_eval_sensor_model(numpy host arrays) [Cython drop-in]"
    else:
        rm_e2e = rm
        e2e_api = "range_libc_b200.PyRayMarchingGPU.calc_range_repeat_angles_eval_sensor_model(numpy host arrays)"
    n_host_sets = 64
    host_sets = [torch.from_numpy(sets_h[i].copy()).pin_memory().numpy() for i in range(n_host_sets)]
    host_angles = torch.from_numpy(angles_h.copy()).pin_memory().numpy()
    host_obs = torch.from_numpy(obs_h.copy()).pin_memory().numpy()
    host_w = torch.empty(N_PART, dtype=torch.float64).pin_memory().numpy()
    for i in range(W_):
        rm_e2e.calc_range_repeat_angles_eval_sensor_model(host_sets[i % n_host_sets], host_angles, host_obs, host_w)
    torch.cuda.synchronize()
    ke = min(K_, 2000)
    t0 = time.perf_counter()
    for i in range(ke):
        rm_e2e.calc_range_repeat_angles_eval_sensor_model(host_sets[(W_ + i) % n_host_sets], host_angles, host_obs, host_w)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_value = N_PART * N_BEAMS * ke / e2e_s
    clocks = sampler.stop()

    peak, peak_src = load_peaks()
    prof = profiled("c2", "fused_kernel")
    algo_bytes = 12 * N_PART + 8 * N_BEAMS + 8 * N_PART  # (12 N + 8 M + 8 N) per launch
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    cfg = workload_config("c2")
    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": 1, "steps": K_, "warmup": W_,
        "ms_per_step": ms / K_, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPE, "data": "synthetic", "config": cfg,
        "gpu_launches": int(launches), "launch_mode": mode, "weight_gather": "none (single GPU)",
        "particle_sets": "%d sets, %.0f MB; value_cold_l2 flushes L2 before every step" % (n_sets, n_sets * N_PART * 12 / 1e6),
        "timing": "CUDA events on the launching stream around one replay of the K-step graph; a spin kernel enqueued ahead "
                  "of the first event hides the host's enqueue time; ms_per_step_eager = the same steps launched one by one",
        "ms_per_step_eager": eager_ms, "kernel_ms": kernel_ms,
        "value_cold_l2": N_PART * N_BEAMS / (cold_ms * 1e-3),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": 12 * N_PART + 8 * N_BEAMS,
                "d2h_bytes_per_step": 8 * N_PART, "ms_per_step": e2e_s / ke * 1e3, "steps": ke, "api": e2e_api},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": prof["dram_bytes"] if prof else None, "kernel": "fused_kernel<RM>",
                     "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src,
                     "traffic_source": prof["source"] if prof else None,
                     "l2_gather": None if not prof else {
                         "l2_sector_bytes_per_launch": prof["l2_sector_bytes"],
                         "achieved_gsectors_per_s": prof["l2_sector_bytes"] / 32.0 / (kernel_ms * 1e-3) / 1e9,
                         "peak_gsectors_per_s": 291.0,
                         "frac": prof["l2_sector_bytes"] / 32.0 / (kernel_ms * 1e-3) / 291e9,
                         "what": "32-byte sectors through L2 per launch (ncu lts__t_sectors, warm capture) over the measured "
                                 "kernel time, against the chip's random-gather rate: one L1-miss request per clock per SM = "
                                 "148 x 1.965 GHz (tools/gather_bench.cu measures 287-292 G/s)"},
                     "note": "0.335 algorithmic B/ray: a 240k-ray launch fits the chip once and is bound by launch + set-up + "
                             "its longest dependent sphere-tracing chain (L2 reads, ~143 ns each), not by HBM (DESIGN.md "
                             "section 4); the HBM fraction is reported because the contract asks for it"},
    }
    threads = os.cpu_count() or 1
    v, kind, done, dt = cpu_fused_run(occ, sets_h, angles_h, obs_h, table, 100000, 3, threads, budget_s=10.0)
    line["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": threads, "kind": kind,
                            "sample": "%d full steps (%.1f s) of the same 4000x60 workload on %d host threads" % (
                                done, dt, threads)}
    v1, _, _, _ = cpu_fused_run(occ, sets_h, angles_h, obs_h, table, 100000, 1, 1, budget_s=5.0)
    line["cpu_baseline"]["single_thread_value"] = v1
    try:
        line["reference_cuda_baseline"] = reference_cuda_run(occ, sets_h, angles_h, obs_h, table)
    except Exception as ex:  # noqa: BLE001
        line["reference_cuda_baseline"] = {"unavailable": str(ex).splitlines()[0][:120]}
    if not args.no_extra:
        line["extra"] = extra_throughput(rl, wl, occ, omap, dev, stream, peak, threads)
    print(json.dumps(line), flush=True)


def run_c5(args, rank, local_rank, world):
    """Config 5, strong scaling: 10^6 particles x 1080 beams sharded over `world` GPUs."""
    import torch
    import torch.distributed as dist
    import range_libc_b200 as rl
    from range_libc_b200 import parallel, workloads as wl

    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W_ = max(args.warmup, 3)
    K_ = max(min(args.steps, 200), 1)
    occ, parts_h, angles_h, obs_h = c5_inputs()
    table = wl.sensor_table(K_TABLE)
    omap = rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
    rm = rl.PyRayMarchingGPU(omap, MAX_RANGE, device=local_rank)
    rm.set_sensor_model(table)
    stream = torch.cuda.current_stream()
    rm.set_stream(stream.cuda_stream)
    lo, hi = parallel.particle_slice(C5_PART, rank, world)
    mine = torch.from_numpy(parts_h[lo:hi]).to(dev)
    angles = torch.from_numpy(angles_h).to(dev)
    obs = torch.from_numpy(obs_h).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gather_mode = "none (single GPU)"
    upd = None
    if world > 1:
        upd = parallel.SignalledSensorUpdate(C5_PART, rm, angles, obs, device=dev)
        gather_mode = ("per rank and step: tile sort of the shard, fan cast into a scratch array, evaluation kernel whose "
                       "epilogue stores every weight into each rank's gathered array over NVLink (symmetric memory, double "
                       "buffered) and raises in-kernel epoch flags, then a wait kernel; no NCCL call, no cross-step pipelining")
        w_local = None
    else:
        w_local = torch.empty(C5_PART, dtype=torch.float64, device=dev)

    def step():
        if upd is not None:
            return upd.update(mine, wait=True)
        rm.calc_range_repeat_angles_eval_sensor_model(mine, angles, obs, w_local)
        return w_local

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(W_):
        step()
    barrier()
    l0 = rl.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K_):
        gathered = step()
    e1.record(stream)
    barrier()
    launches = rl.kernel_launches() - l0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    rays_per_step = C5_PART * C5_BEAMS
    value = rays_per_step * K_ / (ms * 1e-3)

    # every rank: the gathered weights of the last step against a single-GPU recomputation of ALL shards
    verified = None
    if world > 1:
        all_parts = torch.from_numpy(parts_h).to(dev)
        w_check = torch.empty(C5_PART, dtype=torch.float64, device=dev)
        rm.calc_range_repeat_angles_eval_sensor_model(all_parts, angles, obs, w_check)
        torch.cuda.synchronize()
        ok = bool(torch.equal(gathered.view(torch.int64), w_check.view(torch.int64)))
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        verified = bool(flag.item())
        del all_parts, w_check

    # end to end through the drop-in API with HOST buffers: this rank's particles in (pinned), ALL weights out
    cy = cython_module()
    mod = cy if cy is not None else rl
    m_e2e = mod.PyRayMarchingGPU(mod.PyOMap(np.ascontiguousarray(occ.T.astype(bool))), MAX_RANGE)
    m_e2e.set_sensor_model(table)
    h_parts = torch.from_numpy(parts_h[lo:hi].copy()).pin_memory().numpy()
    h_angles = torch.from_numpy(angles_h.copy()).pin_memory().numpy()
    h_obs = torch.from_numpy(obs_h.copy()).pin_memory().numpy()
    h_w = torch.empty(C5_PART, dtype=torch.float64).pin_memory().numpy()
    if world > 1:
        host_upd = parallel.HostShardedSensorUpdate(C5_PART, m_e2e, device=dev)
        e2e_api = ("%s.PyRayMarchingGPU.calc_range_repeat_angles_eval_sensor_model_sharded(numpy host arrays): H2D of this "
                   "rank's particles, cast + signalled evaluation kernel with peer stores, wait, D2H of all weights"
                   % ("range_libc [Cython drop-in]" if cy is not None else "range_libc_b200"))

        def e2e_step():
            host_upd.update(h_parts, h_angles, h_obs, h_w)
    else:
        e2e_api = "%s.PyRayMarchingGPU.calc_range_repeat_angles_eval_sensor_model(numpy host arrays)" % (
            "range_libc [Cython drop-in]" if cy is not None else "range_libc_b200")

        def e2e_step():
            m_e2e.calc_range_repeat_angles_eval_sensor_model(h_parts, h_angles, h_obs, h_w)
    for _ in range(2):
        e2e_step()
    barrier()
    ke = min(K_, 50)
    t0 = time.perf_counter()
    for _ in range(ke):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = rays_per_step * ke / e2e_s
    e2e_ok = None
    if world > 1 and verified is not None:
        e2e_ok = bool(np.array_equal(h_w.view(np.uint64), gathered.cpu().numpy().view(np.uint64)))
    clocks = sampler.stop()

    # strong-scaling base: the same update on rank 0's GPU alone, in the same run (the other ranks wait)
    base = None
    if world > 1:
        if rank == 0:
            all_parts = torch.from_numpy(parts_h).to(dev)
            w1 = torch.empty(C5_PART, dtype=torch.float64, device=dev)
            for _ in range(2):
                rm.calc_range_repeat_angles_eval_sensor_model(all_parts, angles, obs, w1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(5):
                rm.calc_range_repeat_angles_eval_sensor_model(all_parts, angles, obs, w1)
            b.record(stream)
            b.synchronize()
            base_ms = a.elapsed_time(b) / 5
            hp = torch.from_numpy(parts_h.copy()).pin_memory().numpy()
            m1 = mod.PyRayMarchingGPU(mod.PyOMap(np.ascontiguousarray(occ.T.astype(bool))), MAX_RANGE)
            m1.set_sensor_model(table)
            m1.calc_range_repeat_angles_eval_sensor_model(hp, h_angles, h_obs, h_w)
            t0 = time.perf_counter()
            for _ in range(5):
                m1.calc_range_repeat_angles_eval_sensor_model(hp, h_angles, h_obs, h_w)
            base_e2e_ms = (time.perf_counter() - t0) / 5 * 1e3
            base = {"n_gpus": 1, "value": rays_per_step / (base_ms * 1e-3), "ms_per_step": base_ms,
                    "e2e_value": rays_per_step / (base_e2e_ms * 1e-3), "e2e_ms_per_step": base_e2e_ms,
                    "what": "the same 10^6 x 1080 update on rank 0's GPU alone, timed in this run after the sharded "
                            "measurement (5 steps, CUDA events / host clock) while the other ranks wait"}
        barrier()

    if rank == 0:
        peak, peak_src = load_peaks()
        n_local = hi - lo
        prof = profiled_update("c5_twostep", C5_KERNELS)
        # compulsory bytes of the fused call (12 N + 8 M in, 8 N out -- out once per peer when sharded)
        algo_bytes = 12 * n_local + 8 * C5_BEAMS + 8 * n_local * max(world, 1)
        kernel_s = ms * 1e-3 / K_
        achieved = algo_bytes / kernel_s / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None,
                "kernel": "rm_persist_kernel<ANGLES> (lane-re-queuing fan cast into a scratch array, ~90 % of the update) + "
                          "eval_overlap_kernel (table lookups + ordered products)",
                "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src,
                "note": "compulsory-only figure (0.0185 B/ray): a deep fused update runs as two kernels, the cast is bound "
                        "by instruction issue (issue slots 79 % busy, profiles/r02/ncu_c5_twostep.txt), not by HBM; the "
                        "ranges between the two kernels are the 8 B/ray of structure traffic counted in `traffic`"}
        if prof:
            roof["kernel_share_of_update"] = {k: v / prof["seconds"] for k, v in prof["per_kernel_seconds"].items()}
            roof["issue_slots_busy"] = dict(prof["per_kernel_issue_slots_busy"],
                                            what="fraction of the SM issue slots in use (ncu smsp__issue_active, warm capture): "
                                                 "the binding resource of the cast, which HBM is not")
        if prof:
            # the capture is a 200 000-particle launch of the same kernel on the same map: scale per particle
            per_particle = prof["dram_bytes"] / 200000.0
            roof["traffic"] = per_particle * n_local
            roof["traffic_source"] = prof["source"] + ", 200 000-particle launch scaled to this rank's shard"
            roof["with_structure_traffic"] = {"achieved": per_particle * n_local / kernel_s / 1e9,
                                              "frac": per_particle * n_local / kernel_s / 1e9 / peak,
                                              "what": "measured DRAM bytes of both kernels (distance-transform sectors that miss "
                                                      "L2 + ranges written and read back + poses + weights) over the update time: "
                                                      "tile-ordered processing keeps the 268 MB distance transform's working set in "
                                                      "L2 (hit rate 91 %)"}
        cfg = workload_config("c5")
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": K_, "warmup": W_,
            "ms_per_step": ms / K_, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic", "config": cfg,
            "gpu_launches": int(launches), "launch_mode": "eager", "weight_gather": gather_mode,
            "gather_verified": verified, "e2e_gather_verified": e2e_ok,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": 12 * n_local + 8 * C5_BEAMS,
                    "d2h_bytes_per_step": 8 * C5_PART, "ms_per_step": e2e_s / ke * 1e3, "steps": ke, "api": e2e_api,
                    "note": "bytes per rank; max over ranks of the host wall clock"},
            "roofline": roof,
            "strong_scaling_base": base,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if verified is False or e2e_ok is False:
        sys.exit(1)


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


def _roof(bytes_per_unit, rate, peak, structure=None):
    """HBM roofline fractions of a rate (units/s): compulsory-only, and with measured structure traffic per unit."""
    out = {"algorithmic_bytes_per_ray": bytes_per_unit, "hbm_frac_compulsory": bytes_per_unit * rate / 1e9 / peak}
    if structure is not None:
        out["dram_bytes_per_ray_measured"] = structure
        out["hbm_frac_with_structure_traffic"] = structure * rate / 1e9 / peak
    return out


def grid_sample_queries(W, H, step=10, rays=40):
    """Benchmark::grid_sample (RangeLib.h:2006-2023): a `step` lattice of positions x `rays` headings (main.cpp:59-61)."""
    xs, ys = np.arange(0, W, step), np.arange(0, H, step)
    th = (np.arange(rays) * (2.0 * np.pi / rays)).astype(np.float32)
    g = np.stack(np.meshgrid(xs, ys, indexing="ij"), -1).reshape(-1, 2).astype(np.float32)
    q = np.empty((len(g) * rays, 3), np.float32)
    q[:, :2] = np.repeat(g, rays, axis=0)
    q[:, 2] = np.tile(th, len(g))
    return q


def extra_throughput(rl, wl, occ, omap, dev, stream, peak, threads):
    """Secondary numbers (not the judged line), one block per BASELINE configuration, all device resident, each with
    its HBM-roofline fractions (compulsory bytes, and measured DRAM bytes per ray where the structure
    exceeds L2) and the reference's CPU implementation timed in the same run on `threads` host threads."""
    import torch
    out = {"cpu_threads": threads}

    def methods(o):
        def pcddt():
            m = rl.PyCDDTCast(o, MAX_RANGE, 108)
            m.prune()
            return m
        return (("rm", lambda: rl.PyRayMarchingGPU(o, MAX_RANGE)), ("cddt", lambda: rl.PyCDDTCast(o, MAX_RANGE, 108)),
                ("pcddt", pcddt), ("bl", lambda: rl.PyBresenhamsLine(o, MAX_RANGE)))

    # C1: the reference's main.cpp benchmark map (basement_hallways_10cm), random and grid-sample distributions
    try:
        occ1 = wl.load_map("basement_hallways_10cm")
        o1 = rl.PyOMap(np.ascontiguousarray(occ1.T.astype(bool)))
        W1, H1 = occ1.shape
        N = 1 << 24
        q_h = wl.random_queries(W1, H1, N, seed=12345)
        q = torch.from_numpy(q_h).to(dev)
        gq_h = grid_sample_queries(W1, H1)
        gq_h = gq_h[(gq_h[:, 0] >= 1) & (gq_h[:, 1] >= 1)]
        reps = (1 << 22) // len(gq_h) + 1
        gq = torch.from_numpy(np.tile(gq_h, (reps, 1))).to(dev)
        r = torch.empty(max(N, len(gq)), dtype=torch.float32, device=dev)
        c1 = {"map": "basement_hallways_10cm 600x600", "random": "2^24 rays, x~U(1,W-1), y~U(1,H-1), theta~U(0,2pi) (RangeLib.h:2057-2059)",
              "grid": "GRID_STEP 10 x GRID_RAYS 40 lattice (RangeLib.h:2006-2023), %d rays tiled %d times" % (len(gq_h), reps)}
        for nm, ctor in methods(o1):
            m = ctor()
            m.set_stream(stream.cuda_stream)
            n = N if nm != "bl" else N // 4
            t = _time_launches(lambda: m.calc_range_many_grid(q[:n], r[:n]), stream)
            tg = _time_launches(lambda: m.calc_range_many_grid(gq, r[:len(gq)]), stream)
            cpu = cpu_cast_rate("cddt" if nm == "pcddt" else nm, occ1, q_h[:200000], threads, pruned=nm == "pcddt", budget_s=2.0)
            c1[nm] = {"random_rays_per_s": n / t, "grid_rays_per_s": len(gq) / tg, "cpu": cpu, **_roof(16.0, n / t, peak)}
            del m
        out["c1_basement_10cm"] = c1
        del q, gq, r
    except Exception as ex:  # noqa: BLE001
        out["c1_error"] = str(ex)[:200]

    # the 5 cm map (north_star's 50 G rays/s RM target): random rays per method
    W, H = occ.shape
    N = 1 << 24
    q = torch.from_numpy(wl.random_queries(W, H, N, seed=1)).to(dev)
    r = torch.empty(N, dtype=torch.float32, device=dev)
    for nm, ctor in methods(omap):
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
    # lidar-scan batches through calc_range_repeat_angles (numpy_calc_range_angles): 4.2 algorithmic B/ray
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
            bpr = (12.0 * n_p + 4.0 * n_b + 4.0 * n_p * n_b) / (n_p * n_b)
            scans[label] = {"rays_per_s": n_p * n_b / t, **_roof(bpr, n_p * n_b / t, peak)}
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
        qb_h = wl.random_queries(big.shape[0], big.shape[1], n, seed=2)
        qb = torch.from_numpy(qb_h).to(dev)
        rb = torch.empty(n, dtype=torch.float32, device=dev)
        t_q = _time_launches(lambda: cd.calc_range_many_grid(qb, rb), stream, iters=5)
        cd.set_spatial_sort(False)
        t_q0 = _time_launches(lambda: cd.calc_range_many_grid(qb, rb), stream, iters=3)
        cd.set_spatial_sort(True)
        t0 = time.perf_counter()
        cd.prune()
        t_prune = time.perf_counter() - t0
        t_qp = _time_launches(lambda: cd.calc_range_many_grid(qb, rb), stream, iters=5)
        p_c = profiled("c3_cddt", "cddt_")
        p_p = profiled("c3_pcddt", "cddt_")
        cpu = cpu_cast_rate("cddt", big, qb_h[:1 << 20], threads, budget_s=3.0)
        out["c3_gigantic_map"] = {
            "cddt_build_s": t_build, "pcddt_prune_s": t_prune, "cddt_rays_per_s": n / t_q, "pcddt_rays_per_s": n / t_qp,
            "cddt_rays_per_s_direct_search": n / t_q0, "table_bytes_after_prune": cd.memory(),
            "cddt_roofline": _roof(16.0, n / t_q, peak, p_c["dram_bytes"] / n if p_c else None),
            "pcddt_roofline": _roof(16.0, n / t_qp, peak, p_p["dram_bytes"] / n if p_p else None),
            "cpu_cddt": cpu,
            "note": "2^24 random queries in caller order, one kernel; tables larger than L2 are searched through an L2-resident "
                    "index (16-byte record per bin + one skip entry per 64-byte block of zero points), so a query costs one "
                    "DRAM access to the table instead of ~9; cddt_rays_per_s_direct_search is the same batch without the "
                    "index; build / prune include the upload of the 120 MB grid; the CPU reference's prune of this map takes "
                    "~24 min on one thread (SURVEY.md section 6) and is not re-timed here"}
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
        q4_h = wl.random_queries(4096, 4096, n, seed=3)
        q4 = torch.from_numpy(q4_h).to(dev)
        r4 = torch.empty(n, dtype=torch.float32, device=dev)
        frames = []
        for f in range(8):
            blocks = wl.flip_blocks(occ4, f, seed=2026)
            rects = np.array([[x0, y0, p.shape[0], p.shape[1]] for x0, y0, p in blocks], np.int32)
            frames.append((torch.from_numpy(np.concatenate([p.ravel() for _, _, p in blocks])).to(dev), rects))

        def frame(f):
            patches, rects = frames[f % 8]
            bl.update_map_batch(patches, rects)  # all 64 patches of the frame
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
        t_cast = _time_launches(lambda: bl.calc_range_many_grid(q4, r4), stream)
        # CPU: the reference has no map update; a frame = a new BresenhamsLine (O(W*H) map copy) + the rays
        t0 = time.perf_counter()
        cpu = cpu_cast_rate("bl", occ4, q4_h, threads, budget_s=2.0)
        out["c4_dynamic_bl_4096"] = {"ms_per_frame": t * 1e3, "rays_per_s": n / t, "patches_per_frame": 64,
                                     "cast_only_rays_per_s": n / t_cast, **_roof(16.0, n / t_cast, peak),
                                     "cpu": cpu, "cpu_ms_per_frame_estimate": (cpu["construction_s"] + n / cpu["rays_per_s"]) * 1e3,
                                     "note": "update + query per frame; CPU frame = BresenhamsLine construction (map copy) + "
                                             "2^20 rays at the measured rate"}
        del bl, q4, r4, m4
    except Exception as ex:  # noqa: BLE001
        out["c4_error"] = str(ex)[:200]
    # C5: fused RM + sensor model, 1M particles x 1080 beams on a synthetic 8192^2 grid (one GPU)
    try:
        occ5, p5_h, a5_h, o5_h = c5_inputs()
        m5 = rl.PyOMap(np.ascontiguousarray(occ5.T.astype(bool)))
        t0 = time.perf_counter()
        rm5 = rl.PyRayMarchingGPU(m5, MAX_RANGE)
        t_dt = time.perf_counter() - t0
        table = wl.sensor_table(K_TABLE)
        rm5.set_sensor_model(table)
        rm5.set_stream(stream.cuda_stream)
        p5, a5, o5 = (torch.from_numpy(x).to(dev) for x in (p5_h, a5_h, o5_h))
        w5 = torch.empty(C5_PART, dtype=torch.float64, device=dev)
        t = _time_launches(lambda: rm5.calc_range_repeat_angles_eval_sensor_model(p5, a5, o5, w5), stream, iters=3, warm=1)
        rate = C5_PART * C5_BEAMS / t
        prof = profiled_update("c5_twostep", C5_KERNELS)
        sets = [np.ascontiguousarray(p5_h[i * 10000:(i + 1) * 10000]) for i in range(4)]
        v, kind, done, dtc = cpu_fused_run(occ5, sets, a5_h, o5_h, table, 8, 0, threads, budget_s=6.0)
        out["c5_rm_fused_8192"] = {"particles": C5_PART, "beams": C5_BEAMS, "s_per_update": t, "rays_per_s": rate,
                                   "dt_build_s": t_dt,
                                   **_roof((12.0 * C5_PART + 8.0 * C5_BEAMS + 8.0 * C5_PART) / (C5_PART * C5_BEAMS), rate, peak,
                                           prof["dram_bytes"] / (200000.0 * C5_BEAMS) if prof else None),
                                   "cpu": {"rays_per_s": v, "kind": kind, "threads": threads,
                                           "sample": "%d steps of 10 000 particles x 1080 beams" % done},
                                   "note": "268 MB distance transform (> L2); big clouds are processed in tile order so its working "
                                           "set stays in L2; a deep update runs as two kernels (lane-re-queuing cast into a scratch "
                                           "array in chunks of <= 1 GB of ranges, then the streaming evaluation)"}
        del rm5, p5, w5, m5
    except Exception as ex:  # noqa: BLE001
        out["c5_error"] = str(ex)[:200]
    return out


if __name__ == "__main__":
    main()
