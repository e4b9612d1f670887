#!/bin/bash
# round 2, session 4: programmatic dependent launch of back-to-back small fused updates; bench with the spin-kernel bracket
O=gpurun_out/r02z3
mkdir -p $O
echo "== RL_PDL=0" | tee -a $O/tune.log
RL_PDL=0 timeout 300 python tools/tune_fused.py 8 2>&1 | tee -a $O/tune.log
echo "== RL_PDL=1" | tee -a $O/tune.log
timeout 300 python tools/tune_fused.py 8 2>&1 | tee -a $O/tune.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/status.txt
tail -3 $O/pytest.log
RL_PDL=0 timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_nopdl.json 2> $O/bench_nopdl.err; echo "bench nopdl rc=$?" | tee -a $O/status.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/status.txt
python - <<'PY'
import json
for f in ("bench_nopdl", "bench"):
    try:
        d = json.loads(open("gpurun_out/r02z3/%s.json" % f).read().strip().split("\n")[-1])
        print(f, "value %.3f G  ms %.5f eager %.5f cold %.3f G e2e %.3f G (%.2f us)" % (d["value"] / 1e9, d["ms_per_step"], d["ms_per_step_eager"], d["value_cold_l2"] / 1e9, d["e2e"]["value"] / 1e9, d["e2e"]["ms_per_step"] * 1e3))
    except Exception as e:
        print(f, "unreadable", e)
PY
