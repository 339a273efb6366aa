#!/bin/bash
# ncu captures of a round, run on the GPU box:   gpurun -- 'bash profiles/capture.sh r02'
# (launch lists: per-launch times are cold-cache and serialised, compare shares; --set full: one capture per kernel of interest)
r=${1:-r02}; out=gpurun_out/$r; mkdir -p $out
NCU="ncu --clock-control none"
for wl in c5 c2 c4; do
  $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $out/launches_$wl.csv python bench.py --workload $wl --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > $out/launches_$wl.log 2>&1
done
# the eight depth-band launches of one C5 step
$NCU --set full --import-source on -k regex:k_synth -s 24 -c 8 -o $out/prof_synth_c5 -f python bench.py --workload c5 --steps 1 --warmup 3 --no-secondary --no-cpu-baseline > $out/ncu_synth_c5.log 2>&1
$NCU --set full --import-source on -k regex:k_geometry -s 3 -c 1 -o $out/prof_geometry_c5 -f python bench.py --workload c5 --steps 1 --warmup 3 --no-secondary --no-cpu-baseline > $out/ncu_geometry_c5.log 2>&1
$NCU --set full --import-source on -k regex:k_mt_fused -s 3 -c 1 -o $out/prof_mt_fused_c2 -f python bench.py --workload c2 --steps 1 --warmup 3 --no-secondary --no-cpu-baseline > $out/ncu_mt_fused_c2.log 2>&1
# the reference-order synthesis (bulk-copy staging of node blocks) on one C3 candidate and the eikonal kernels on a C4 step
$NCU --set full --import-source on -k regex:k_synth_exact -c 1 -o $out/prof_synth_exact_c3 -f python scratch/time_exact.py > $out/ncu_synth_exact.log 2>&1
$NCU --metrics gpu__time_duration.sum -k regex:k_eik -c 40 --csv --log-file $out/launches_eik.csv python bench.py --workload c4 --steps 1 --warmup 3 --no-secondary --no-cpu-baseline > $out/launches_eik.log 2>&1
ls -la $out
