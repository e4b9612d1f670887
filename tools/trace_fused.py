"""Per-phase timeline of the small fused launch (development aid).

    python tools/trace_fused.py --build     # here: nvcc -DRL_TRACE -> tools/_trace/librangelib_b200_trace.so
    python tools/trace_fused.py             # on a GPU box: runs the traced library and prints the timeline

The traced library stamps %globaltimer per CTA at: 0 kernel entry, 1 march entry (set-up done), 2 end of the first
own-ray burst, 3 end of the last burst (hand-off decided), 5 cooperative tail done, 6 table values stored,
7 product done; slots 8-10 hold the steps at hand-off and the rays alive after the first / last burst."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TRACE_DIR = os.path.join(ROOT, "tools", "_trace")
TRACE_LIB = os.environ.get("RL_TRACE_LIB") or os.path.join(TRACE_DIR, "librangelib_b200_trace.so")


def build():
    from range_libc_b200 import build as b
    os.makedirs(TRACE_DIR, exist_ok=True)
    objs, procs = [], []
    for s in b.SOURCES:
        obj = os.path.join(TRACE_DIR, s.replace(".cu", ".o"))
        objs.append(obj)
        procs.append(subprocess.Popen([b.nvcc()] + b.NVCC_FLAGS + ["-DRL_TRACE", "-c", os.path.join(b.CSRC, s), "-o", obj]))
    assert all(p.wait() == 0 for p in procs)
    subprocess.check_call([b.nvcc(), "-shared", "-o", TRACE_LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                           "-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    print(TRACE_LIB)


def run():
    import torch
    from range_libc_b200 import cabi
    cabi.LIB_PATH = TRACE_LIB  # before the first lib() call
    import bench
    import range_libc_b200 as rl
    from range_libc_b200 import workloads as wl
    L = cabi.lib()
    L.rl_debug_set_trace.argtypes = [C.c_void_p]
    occ = wl.load_map(bench.MAP)
    sets_h, angles_h, obs_h = bench.make_inputs(occ, 16)
    rm = rl.PyRayMarchingGPU(rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool))), bench.MAX_RANGE)
    rm.set_sensor_model(wl.sensor_table(bench.K_TABLE))
    st = torch.cuda.current_stream()
    rm.set_stream(st.cuda_stream)
    sets = torch.from_numpy(sets_h).cuda()
    angles, obs = torch.from_numpy(angles_h).cuda(), torch.from_numpy(obs_h).cuda()
    w = torch.empty(bench.N_PART, dtype=torch.float64, device="cuda")
    trace = torch.zeros(1024 * 16, dtype=torch.int64, device="cuda")
    assert L.rl_debug_set_trace(C.c_void_p(trace.data_ptr())) == 0
    names = {1: "march entry (set-up done)", 2: "first burst done", 3: "last burst done", 5: "cooperative tail done",
             6: "table values stored", 7: "product done"}
    for label, idx in (("global", (0, 2, 4, 6)), ("tracking", (1, 3, 5, 7))):
        rows = []
        for i in idx:
            for _ in range(3):
                rm.calc_range_repeat_angles_eval_sensor_model(sets[i], angles, obs, w)
            torch.cuda.synchronize()
            trace.zero_()
            rm.calc_range_repeat_angles_eval_sensor_model(sets[i], angles, obs, w)
            torch.cuda.synchronize()
            t = trace.cpu().numpy().reshape(1024, 16)[:1000].astype(np.float64)
            t0 = t[:, 0].min()
            t[:, 5] = np.maximum(t[:, 5], t[:, 3])  # CTAs that never handed rays off
            rows.append((t, t0))
        print("== %s clouds: ns since the first CTA's entry, over %d launches x 1000 CTAs" % (label, len(rows)))
        ent = np.concatenate([t[:, 0] - t0 for t, t0 in rows])
        print("  %-28s median %7.0f  p90 %7.0f  max %7.0f" % ("kernel entry", np.median(ent), np.percentile(ent, 90), ent.max()))
        for slot in (1, 2, 3, 5, 6, 7):
            v = np.concatenate([t[:, slot] - t0 for t, t0 in rows])
            print("  %-28s median %7.0f  p90 %7.0f  max %7.0f" % (names[slot], np.median(v), np.percentile(v, 90), v.max()))
        steps = np.concatenate([t[:, 8] for t, _ in rows])
        a1 = np.concatenate([t[:, 9] for t, _ in rows])
        a2 = np.concatenate([t[:, 10] for t, _ in rows])
        print("  steps at hand-off: median %.0f max %.0f | alive after first burst: median %.0f max %.0f | at hand-off: median %.0f max %.0f" % (
            np.median(steps), steps.max(), np.median(a1), a1.max(), np.median(a2), a2.max()))
        nb, nf, ne = (np.concatenate([t[:, c] for t, _ in rows]) for c in (11, 12, 13))
        print("  cooperative tail, per CTA: batches median %.0f max %.0f | replay steps median %.0f max %.0f | of them through "
              "the exact test: %.1f %% overall (max CTA %.0f)" % (np.median(nb), nb.max(), np.median(nf), nf.max(),
                                                                 100.0 * ne.sum() / max(nf.sum(), 1), ne.max()))
        # the slowest CTA of each launch
        for t, t0 in rows:
            k = int(np.argmax(t[:, 7]))
            print("  slowest CTA %4d: entry %5.0f march %5.0f burst1 %5.0f bursts %5.0f coop %5.0f stored %5.0f product %5.0f (steps %d, alive %d -> %d)" % (
                k, t[k, 0] - t0, t[k, 1] - t0, t[k, 2] - t0, t[k, 3] - t0, t[k, 5] - t0, t[k, 6] - t0, t[k, 7] - t0, t[k, 8], t[k, 9], t[k, 10]))
            print("                    tail of that CTA: %d batches, %d replay steps, %d through the exact test" % (t[k, 11], t[k, 12], t[k, 13]))
            c14, c15 = int(t[k, 14]), int(t[k, 15])
            print("                    its longest ray: %d steps in %d batches, %d cycles in the tail of which %d inside replay loops" % (
                c15 >> 32, c15 & 0xffffffff, c14 >> 32, c14 & 0xffffffff))


if __name__ == "__main__":
    build() if "--build" in sys.argv else run()
