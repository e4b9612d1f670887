#!/bin/bash
O=gpurun_out/r02m
mkdir -p $O
python -m pytest tests/test_gpu_scale_parity.py -m gpu -x -q -k "c3_indexed or checkpoint" > $O/test.log 2>&1; echo "c3 tests rc=$?" > $O/status.txt
tail -15 $O/test.log
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fresh or fuzz or golden or structures or big_maps or radial" >> $O/test.log 2>&1; echo "parity subset rc=$?" >> $O/status.txt
tail -3 $O/test.log
python - > $O/time.log 2>&1 <<'PY'
import sys, time, numpy as np, torch
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
    for idx in (True, False):
        cd.set_spatial_sort(idx)
        ms = t(lambda: cd.calc_range_many_grid(q, out))
        print("C3 pruned=%s indexed=%s  %.3f ms  %.2f G rays/s  (memory %.1f MB)" % (pr, idx, ms, n / ms / 1e6, cd.memory() / 1e6), flush=True)
# small table: index forced vs direct
occ = wl.load_map("basement_hallways_5cm")
cs = rl.PyCDDTCast(rl.PyOMap(np.ascontiguousarray(occ.T.astype(bool))), 500.0, 108)
cs.set_stream(st.cuda_stream)
q2 = torch.from_numpy(wl.random_queries(occ.shape[0], occ.shape[1], n, seed=2)).cuda()
for idx in (1, 2):
    cs.set_spatial_sort(idx)
    ms = t(lambda: cs.calc_range_many_grid(q2, out))
    print("5cm CDDT index=%s  %.3f ms  %.2f G rays/s" % (idx == 2, ms, n / ms / 1e6), flush=True)
PY
cat $O/time.log
for w in c3_cddt c3_pcddt; do
  ncu --set full --cache-control none --clock-control none -k regex:cast_kernel -s 2 -c 1 -o $O/$w -f python tools/prof_r02.py $w 3 > $O/$w.log 2>&1
  python tools/ncu_summary.py $O/$w.ncu-rep > $O/$w.txt 2>&1
  grep -E "duration|dram__bytes|lts__t_sectors.sum|hit_rate|sectors per" $O/$w.txt
done
cat $O/status.txt
