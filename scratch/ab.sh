#!/bin/bash
# A/B bench of library variants in one box: scratch/ab.sh libA.so libB.so ...
# each variant is copied over kiwi_b200/libkiwi_b200.so and benched twice, interleaved
cp kiwi_b200/libkiwi_b200.so scratch/_orig.so
for rep in 1 2; do
  for v in "$@"; do
    cp scratch/$v kiwi_b200/libkiwi_b200.so
    python bench.py --steps 6 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['value'],1), 'evals/s  synth ms', round(d['stage_ms_per_step']['synthesis'],2), 'clk', d['clocks']['sm_mhz'])"
  done
done
cp scratch/_orig.so kiwi_b200/libkiwi_b200.so
