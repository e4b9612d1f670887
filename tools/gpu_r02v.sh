#!/bin/bash
O=gpurun_out/r02v
mkdir -p $O
python -m pytest tests/test_gpu_scale_parity.py -m gpu -x -q -k "c3_indexed or checkpoint" > $O/test.log 2>&1; echo "c3 tests rc=$?" > $O/status.txt
tail -3 $O/test.log
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fresh or fuzz or golden or radial or structures" >> $O/test.log 2>&1; echo "cddt tests rc=$?" >> $O/status.txt
tail -3 $O/test.log
python - > $O/time.log 2>&1 <<'PY'
import sys, os, time, numpy as np, torch
sys.path.insert(0, '.')
import range_libc_b200 as rl
from range_libc_b200 import workloads as wl
st = torch.cuda.current_stream()
def t(fn, it=5):
    for _ in range(2): fn()
    ts=[]
    for _ in range(it):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(st); fn(); b.record(st); b.synchronize(); ts.append(a.elapsed_time(b))
    return np.median(ts)
big = wl.load_map("gigantic_map")
cd = rl.PyCDDTCast(rl.PyOMap(np.ascontiguousarray(big.T.astype(bool))), 500.0, 108)
cd.set_stream(st.cuda_stream)
n = 1 << 24
q = torch.from_numpy(wl.random_queries(big.shape[0], big.shape[1], n, seed=2)).cuda()
out = torch.empty(n, dtype=torch.float32, device="cuda")
for pr in (False, True):
    if pr: cd.prune()
    ms = t(lambda: cd.calc_range_many_grid(q, out))
    print("C3 pruned=%s indexed (16-bit skip codes)  %.3f ms  %.2f G rays/s (memory %.1f MB)" % (pr, ms, n / ms / 1e6, cd.memory()/1e6), flush=True)
PY
cat $O/time.log
for w in c3_cddt c3_pcddt; do
  ncu --set full --cache-control none --clock-control none -k regex:cddt_batch -s 2 -c 1 -o $O/$w -f python tools/prof_r02.py $w 3 > $O/$w.log 2>&1
  python tools/ncu_summary.py $O/$w.ncu-rep > $O/$w.txt 2>&1
  grep -E "duration|registers|warps_active|dram__bytes|lts__t_sectors.sum|hit_rate|sectors per|long_scoreboard|issue_active" $O/$w.txt
done
cat $O/status.txt
