#!/bin/bash
# ncu --set full over the kernels that are new in round 2 outside the UNet forward: the hand-written training encoder
# (weight-gradient GEMM: all 53 launches of one step; a sample of the BatchNorm / layout kernels), the 5x5x5 median.
# Only the CSV summaries (tools/ncu_extract.py) come back.  Keep the launch counts small: ncu saves and restores device
# memory around each of ~40 replay passes per kernel (318 BatchNorm launches did not finish in 25 minutes).
#   bash tools/ncu_train.sh <tag> [wgrad]
set -u
tag=${1:-r02}
out=gpurun_out
export TIME_ITERS=1 CDDPM_ENC_GRAPH=0
cap() {  # name regex skip count command...
  local name=$1 re=$2 skip=$3 cnt=$4
  shift 4
  timeout 300 ncu --set full --clock-control none -k "regex:$re" -s $skip -c $cnt -f -o /tmp/${tag}_$name "$@" > $out/${tag}_ncu_$name.log 2>&1
  python tools/ncu_extract.py /tmp/${tag}_$name.ncu-rep > $out/${tag}_ncu_$name.csv 2>> $out/${tag}_ncu_$name.log
  wc -l $out/${tag}_ncu_$name.csv
}
cap median "median5_pair|median3d" 5 2 python tools/time_median.py
# first layers of the 5th step (4 warm-up steps): stem + layer1.0
cap enc_bn "bn_stats|bn_finalize|bn_apply|bn_bwd" 1272 12 python tools/time_encoder_train.py 64 b200
cap enc_misc "pack_panels|im2col|col2im3|mask_relu|maxpool|unpack_grad" 284 8 python tools/time_encoder_train.py 64 b200
if [ "${2:-}" = "wgrad" ]; then cap enc_wgrad flat_wgrad_tc 212 53 python tools/time_encoder_train.py 64 b200; fi
