#!/bin/bash
# N-GPU validation: all gather paths under pytest (when N >= 2), multi_gpu_check on N ranks, both bench arms at N
N=${1:-2}
O=gpurun_out/r02n$N
mkdir -p $O
nvidia-smi -L > $O/box.txt
if [ "$N" = "2" ]; then
  python -m pytest tests/test_gpu_scale_parity.py -m gpu -x -q -k "multi_gpu" > $O/multi_gpu_pytest.log 2>&1; echo "multi-gpu pytest rc=$?" > $O/status.txt
  tail -3 $O/multi_gpu_pytest.log
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu_check.py > $O/multi_gpu_check.log 2>&1; echo "multi_gpu_check rc=$?" >> $O/status.txt
grep "rank" $O/multi_gpu_check.log | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench n$N rc=$?" >> $O/status.txt
tail -3 $O/bench_n$N.err
python - $O/bench_n$N.json <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        b = d.get('strong_scaling_base', {})
        print("N=%d value %.4g rays/s (%.3f ms/step) e2e %.4g (%.3f ms) verified %s/%s | N=1 base in the same run: %.4g (%.3f ms), e2e %.4g" % (
            d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d.get('gather_verified'), d.get('e2e_gather_verified'),
            b.get('value', 0), b.get('ms_per_step', 0), b.get('e2e_value', 0)))
        if b: print("strong-scaling efficiency: value %.3f  e2e %.3f" % (d['value'] / (d['n_gpus'] * b['value']), d['e2e']['value'] / (d['n_gpus'] * b['e2e_value'])))
PY
if [ "$N" = "2" ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus $N --steps 5 --warmup 1 > $O/bench_ref_n$N.json 2> $O/bench_ref_n$N.err; echo "ref n$N rc=$?" >> $O/status.txt
  cut -c1-300 $O/bench_ref_n$N.json
fi
cat $O/status.txt
