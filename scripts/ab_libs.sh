#!/bin/bash
# same-box A/B of two builds of the library (CIDNET_LIB): step time of the replayed graph, alternating, 3 rounds
# usage: bash scripts/ab_libs.sh libA.so libB.so [B H W]
A=$1; Bl=$2; shift 2; DIMS=${@:-1 640 1120}
for r in 1 2 3; do
  for L in $A $Bl; do
    echo -n "$(basename $L): "; CIDNET_LIB=$PWD/$L python scripts/time_kernels.py $DIMS zzz | head -1
  done
done
