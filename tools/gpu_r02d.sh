#!/bin/bash
O=gpurun_out/r02d
mkdir -p $O
python -m pytest tests/test_gpu_scale_parity.py -m gpu -x -q -k "c3_bin" > $O/c3test.log 2>&1; echo "c3 test rc=$?" > $O/status.txt
tail -15 $O/c3test.log
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rm or fused or golden" > $O/rmtest.log 2>&1; echo "rm tests rc=$?" >> $O/status.txt
tail -3 $O/rmtest.log
RL_BLOCK_BURST_PAIRS=6 python tools/tune_fused.py 4 8 12 16 24 32 > $O/tune_fused.log 2>&1
RL_BLOCK_BURST_PAIRS=4 python tools/tune_fused.py 8 16 >> $O/tune_fused.log 2>&1
cat $O/tune_fused.log
python - > $O/c3_time.log 2>&1 <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import range_libc_b200 as rl
from range_libc_b200 import workloads as wl
big = wl.load_map("gigantic_map")
cd = rl.PyCDDTCast(rl.PyOMap(np.ascontiguousarray(big.T.astype(bool))), 500.0, 108)
st = torch.cuda.current_stream(); cd.set_stream(st.cuda_stream)
n = 1 << 24
q = torch.from_numpy(wl.random_queries(big.shape[0], big.shape[1], n, seed=2)).cuda()
out = torch.empty(n, dtype=torch.float32, device="cuda")
def t(fn, it=5):
    for _ in range(2): fn()
    ts=[]
    for _ in range(it):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(st); fn(); b.record(st); b.synchronize(); ts.append(a.elapsed_time(b))
    return np.median(ts)
for pr in (False, True):
    if pr: cd.prune()
    for sort in (True, False):
        cd.set_spatial_sort(sort)
        ms = t(lambda: cd.calc_range_many_grid(q, out))
        print("pruned=%s sorted=%s  %.3f ms  %.2f G rays/s" % (pr, sort, ms, n / ms / 1e6), flush=True)
PY
cat $O/c3_time.log; cat $O/status.txt
