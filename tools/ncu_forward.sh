#!/bin/bash
# ncu --set full over ONE B=32 UNet forward (the 4th of tools/time_forward.py), per kernel family; only the CSV
# summaries (tools/ncu_extract.py) and one small report with source come back.  Run under gpurun:
#   bash tools/ncu_forward.sh <tag>
set -u
tag=${1:-r02}
out=gpurun_out
export TIME_FORWARD_ITERS=1
python tools/time_forward.py 32 > $out/${tag}_ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 $out/${tag}_ncu_plain.log; exit 1; }
cap() {  # name regex skip count
  ncu --set full --clock-control none -k "regex:$2" -s $3 -c $4 -f -o /tmp/${tag}_$1 python tools/time_forward.py 32 > $out/${tag}_ncu_$1.log 2>&1
  python tools/ncu_extract.py /tmp/${tag}_$1.ncu-rep > $out/${tag}_ncu_$1.csv 2>> $out/${tag}_ncu_$1.log
  wc -l $out/${tag}_ncu_$1.csv
}
cap conv conv_igemm2 168 56
cap gn gn_apply 165 55
cap misc "attention_tc|conv_in2|conv_out|conv_igemm_kernel" 0 40
# a small report with source for the dominant kernel (first four convolutions of the forward)
ncu --set full --clock-control none --import-source on -k regex:conv_igemm2 -s 168 -c 4 -f -o $out/${tag}_conv_top python tools/time_forward.py 32 > $out/${tag}_ncu_top.log 2>&1
ls -la $out/${tag}_conv_top.ncu-rep
