"""How close are the reference's own CUDA kernels (oracle/_ref/libref_cuda.so: kernels.cu for sm_100a) to the
reference's CPU RayMarching (STRICT build) and to this library?  Prints agreement statistics and timings.
Run on a GPU box: python tests/probes/rmgpu_probe.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import range_libc_b200 as rl  # noqa: E402
from oracle import ref  # noqa: E402
from range_libc_b200 import workloads as wl  # noqa: E402


def stats(name, a, b):
    d = np.abs(a - b)
    rel = d / np.maximum(np.abs(b), 1e-6)
    print("%-28s bit-equal %.6f  within 1e-4 rel %.6f  >1px %.2e  max abs %.3f" % (
        name, np.mean(a == b), np.mean(rel <= 1e-4), np.mean(d > 1.0), d.max()), flush=True)


def main():
    occ = wl.load_map("basement_hallways_5cm")
    n = 1 << 20  # 4 chunks of 262144
    q = wl.random_queries(occ.shape[0], occ.shape[1], n, seed=12345)
    cpu = ref.RefMethod(ref.RM, ref.RefMap(occ=occ, flavor="strict"), 500.0, threads=os.cpu_count())
    want = cpu.calc_range_many(q)
    gmap = ref.RefMap(occ=occ, flavor="cuda")
    gpu = ref.RefMethod(ref.RMGPU, gmap, 500.0)
    got_ref = gpu.calc_range_many(q)
    omap = rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool)))
    ours = rl.PyRayMarchingGPU(omap, 500.0)
    got = np.empty(n, np.float32)
    ours.calc_range_many_grid(q, got)
    stats("reference CUDA vs ref CPU", got_ref, want)
    stats("this library vs ref CPU", got, want)
    stats("reference CUDA vs this lib", got_ref, got)
    for name, fn in (("reference CUDA (host bufs)", lambda: gpu.calc_range_many(q)),
                     ("this library  (host bufs)", lambda: ours.calc_range_many_grid(q, got))):
        fn()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        dt = (time.perf_counter() - t0) / 10
        print("%-28s %.3f ms per 2^20 rays = %.2f G rays/s" % (name, dt * 1e3, n / dt / 1e9), flush=True)


if __name__ == "__main__":
    main()
