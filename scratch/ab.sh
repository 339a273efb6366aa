#!/bin/bash
# A/B of library variants in one box: scratch/ab.sh "<band settings>" libA.so libB.so ...   (variants live in scratch/; "cur" = the built library)
bands=$1; shift
cp kiwi_b200/libkiwi_b200.so scratch/_cur.so
for rep in 1 2; do
  for v in "$@"; do
    if [ "$v" = cur ]; then cp scratch/_cur.so kiwi_b200/libkiwi_b200.so; else cp scratch/$v kiwi_b200/libkiwi_b200.so; fi
    echo "== $v"; python scratch/band_sweep.py c5 $bands 2>&1 | tail -n +1
  done
done
cp scratch/_cur.so kiwi_b200/libkiwi_b200.so
