#!/bin/bash
# Run ON THE GPU BOX (via gpurun): ncu launch list + one full capture per kernel of the bench command.
#   gpurun --timeout 1500 -- 'bash tools/profile.sh r1 dna_100x100k'
# Outputs land in gpurun_out/; tools/summarize_ncu.py turns them into the tracked files in profiles/.
set -u
TAG=${1:-r1}
WL=${2:-dna_100x100k}
OUT=gpurun_out
mkdir -p $OUT
CMD="timeout 240 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-traffic --workload $WL"
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_${TAG}_${WL}.csv $CMD > $OUT/ncu_list_${TAG}_${WL}.log 2>&1
# full captures
ncu --set full --clock-control none --import-source on -k regex:'k_traverse|k_partial' -s 3 -c 1 -f -o $OUT/prof_${TAG}_${WL}_k1 $CMD > $OUT/ncu_k1_${TAG}_${WL}.log 2>&1
# (for 4-state / 4-category data the edge reduction is the epilogue of the traversal kernel: no separate K2 launch in the step)
ncu --set full --clock-control none --import-source on -k regex:'k_pmat' -s 3 -c 1 -f -o $OUT/prof_${TAG}_${WL}_k0 $CMD > $OUT/ncu_k0_${TAG}_${WL}.log 2>&1
ls -la $OUT | tail -20
