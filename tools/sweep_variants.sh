#!/bin/bash
# usage: tools/sweep_variants.sh N B "cfg cfg ..." variant...   (variants from build/variants; "main" = shipped lib)
N=$1; B=$2; CFGS=$3; shift 3
for v in "$@"; do
  for bt in ${BLOCKS:-256}; do
    if [ "$v" = main ]; then unset SB200_LIB; else export SB200_LIB=build/variants/libstencil_b200_$v.so; fi
    echo "== variant $v block_threads $bt"
    SB200_BLOCK_THREADS=$bt python tools/sweep.py $N $B $CFGS 2>&1 | grep -v 'measured DFMA'
  done
done
