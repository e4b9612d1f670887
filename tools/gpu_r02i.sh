#!/bin/bash
# 2-GPU validation: all gather paths under pytest, then both bench arms at N=2
O=gpurun_out/r02i
mkdir -p $O
nvidia-smi -L > $O/box.txt
python -m pytest tests/test_gpu_scale_parity.py -m gpu -x -q -k "multi_gpu" > $O/multi_gpu_pytest.log 2>&1; echo "multi-gpu pytest rc=$?" > $O/status.txt
tail -5 $O/multi_gpu_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu_check.py > $O/multi_gpu_check.log 2>&1; echo "multi_gpu_check rc=$?" >> $O/status.txt
grep "rank" $O/multi_gpu_check.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 ) > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?" >> $O/status.txt
tail -4 $O/bench_n2.err; cat $O/bench_n2.json | cut -c1-3000
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "ref n2 rc=$?" >> $O/status.txt
cut -c1-400 $O/bench_ref_n2.json
cat $O/status.txt
