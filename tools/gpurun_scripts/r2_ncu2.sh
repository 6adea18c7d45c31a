#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
./tools/micro/randrow > $OUT/r2_randrow.log 2>&1; cat $OUT/r2_randrow.log
CMD="python tools/probe.py --goals 1 --roots 8880 --days 1"
$CMD > $OUT/r2n2_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sssp_group -c 1 -o $OUT/r2n2_group -f $CMD > $OUT/r2n2_ncu.log 2>&1
echo "exit $?"; tail -3 $OUT/r2n2_plain.log; tail -5 $OUT/r2n2_ncu.log
