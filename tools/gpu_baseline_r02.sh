#!/bin/bash
# Round-2 "before" captures: GPU tests, a bench line, and ncu --set full (warm: --cache-control none) of every
# kernel VERDICT r01 listed as unprofiled.  Run under gpurun from the repo root.
O=gpurun_out/r02a
mkdir -p $O
nvidia-smi -L > $O/box.txt; nproc >> $O/box.txt
python -m pytest tests -m gpu -x -q > $O/gputest.log 2>&1; echo "gputest rc=$?" >> $O/box.txt
python bench.py --steps 200 --warmup 20 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/box.txt
NCU="ncu --set full --clock-control none --cache-control none"
prof() {  # name kernel-regex skip count [extra ncu flags]
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 $NCU "$@" -k regex:$rx -s $skip -c $cnt -o $O/$name -f python tools/prof_r02.py $name > $O/$name.log 2>&1
  echo "$name rc=$?" >> $O/box.txt
}
prof fan_tracking rm_persist_kernel 2 1 --import-source on
prof fan_uniform rm_persist_kernel 2 1
prof rm_random rm_persist_kernel 2 1
prof fused_deep fused_rm_persist_kernel 2 1 --import-source on
prof c2 fused_kernel 16 4 --import-source on
prof c3_cddt cast_kernel 2 1 --import-source on
prof c3_pcddt cast_kernel 2 1
prof bl bl_persist_kernel 2 1 --import-source on
prof c4_bl bl_persist_kernel 2 1
prof c5 fused_kernel 1 1
prof edt_1200 edt_pass_kernel 2 2
prof edt_8192 edt_pass_kernel 0 2
ls -la $O >> $O/box.txt
