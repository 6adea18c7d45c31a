#!/bin/bash
# A/B run of the SSSP kernel variants in one gpurun call: parity suite per variant, then the cfg3 probe.
# usage: tools/ab_variants.sh "<variant options>;<variant options>;..."   (S4MPI_OPTIONS syntax, empty = defaults)
set -u
OUT=gpurun_out
mkdir -p $OUT
IFS=';' read -ra VARS <<< "${1:-}"
for v in "${VARS[@]}"; do
    tag=$(echo "${v:-default}" | tr '=,' '__')
    echo "=== parity: $v" | tee -a $OUT/ab.log
    S4MPI_OPTIONS="$v" timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q -m gpu > $OUT/parity_$tag.log 2>&1
    echo "exit $?" | tee -a $OUT/ab.log
    tail -3 $OUT/parity_$tag.log | tee -a $OUT/ab.log
done
