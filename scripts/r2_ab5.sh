#!/bin/bash
# A/B: fused vs classic step per organizations-per-GPU regime, decoder grid sweep in fused mode
set -u
mkdir -p gpurun_out
out=gpurun_out/r2_ab5.jsonl; : > $out
run() { echo "# $*" >> $out; env "$@" python scripts/ab_round.py 3 20 $W >> $out 2>> gpurun_out/r2_ab5.err; }
W=1; run DMT_STEP=fused; run DMT_STEP=classic
W=1; run DMT_STEP=fused DMT_DEC_BLOCKS=74; run DMT_STEP=fused DMT_DEC_BLOCKS=148; run DMT_STEP=fused DMT_DEC_BLOCKS=296
W=2; run DMT_STEP=fused; run DMT_STEP=classic
W=4; run DMT_STEP=fused; run DMT_STEP=classic
W=8; run DMT_STEP=fused; run DMT_STEP=classic
cat $out
