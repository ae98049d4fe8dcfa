#!/bin/bash
# Run ON THE GPU BOX (via gpurun): ncu captures of the kernels outside the bench step (parsimony traversal, batched
# SPR candidates, 20-state K3).  Outputs land in gpurun_out/; the numbers quoted in profiles/ncu_r2_aux.md come from them.
set -u
OUT=gpurun_out
mkdir -p $OUT
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,launch__grid_size
ncu --metrics $M --clock-control none -k regex:'k_pars_fitch' -s 2 -c 2 --csv --log-file $OUT/ncu_aux_pars.csv python tools/pars_bench.py > $OUT/ncu_aux_pars.log 2>&1
ncu --metrics $M --clock-control none -k regex:'k_spr_candidates' -s 1 -c 1 --csv --log-file $OUT/ncu_aux_spr_dna.csv python tools/latency_probe.py dna_100x50k > $OUT/ncu_aux_spr_dna.log 2>&1
ncu --metrics $M --clock-control none -k regex:'k_spr_candidates|k_eigen_lr_reg' -s 1 -c 2 --csv --log-file $OUT/ncu_aux_spr_aa.csv python tools/latency_probe.py aa_200x50k > $OUT/ncu_aux_spr_aa.log 2>&1
tail -4 $OUT/ncu_aux_pars.csv $OUT/ncu_aux_spr_dna.csv $OUT/ncu_aux_spr_aa.csv | cut -c1-300
