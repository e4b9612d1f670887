#!/bin/bash
# Round-2 final captures: ncu --set full (warm: --cache-control none) of the kernels as they are at the end of the
# round, summarised by tools/ncu_summary.py.  Run under gpurun from the repo root.
O=gpurun_out/r02prof
mkdir -p $O
NCU="ncu --set full --clock-control none --cache-control none"
prof() {  # name kernel-regex skip count
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 $NCU "$@" -k regex:$rx -s $skip -c $cnt -o $O/$name -f python tools/prof_r02.py $name > $O/$name.log 2>&1
  echo "$name rc=$?" >> $O/status.txt
  python tools/ncu_summary.py $O/$name.ncu-rep > $O/ncu_$name.txt 2>&1
}
prof fan_tracking rm_persist_kernel 2 1
prof fan_uniform rm_persist_kernel 2 1
prof rm_random rm_persist_kernel 2 1
prof fused_deep fused_rm_persist_kernel 2 1
prof c2 fused_kernel 16 4
prof c3_cddt cddt_batch_kernel 2 1
prof c3_pcddt cddt_batch_kernel 2 1
prof bl bl_persist_kernel 2 1
prof c4_bl bl_persist_kernel 2 1
prof c5 fused_kernel 1 1
prof edt_1200 edt_ 4 4
prof edt_8192 edt_ 0 4
rm -f $O/*.ncu-rep
cat $O/status.txt
grep -h -E "^## |duration" $O/ncu_*.txt | head -80
