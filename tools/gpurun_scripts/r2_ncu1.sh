#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
CMD="python tools/probe.py --goals 1 --roots 2220 --days 1"
$CMD > $OUT/r2n1_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sssp_group -c 1 -o $OUT/r2n1_group -f $CMD > $OUT/r2n1_ncu.log 2>&1
echo "exit $?"; tail -3 $OUT/r2n1_plain.log; tail -5 $OUT/r2n1_ncu.log
