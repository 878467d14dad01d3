#!/bin/bash
# ncu captures of the shipping kernels (run under gpurun on ONE GPU).  The X-stream kernels are cooperative launches whose
# CTAs wait for each other, so they are captured with --replay-mode application (the whole program is re-run per metric
# pass) and a short metric list instead of `--set full` kernel replay.
set -u
OUT=${1:-gpurun_out/r2}
mkdir -p $OUT
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum"
run() {  # name, kernel regex, target args..., optional env via NCU_ENV
  name=$1; regex=$2; shift 2
  env $NCU_ENV python tools/ncu_target.py "$@" > $OUT/plain_$name.log 2>&1 || { echo "plain run of $name failed"; tail -3 $OUT/plain_$name.log; return; }
  env $NCU_ENV ncu --replay-mode application --metrics $M --clock-control none -k regex:"$regex" --csv --log-file $OUT/ncu_$name.csv \
      python tools/ncu_target.py "$@" > $OUT/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
}
NCU_ENV="PRMF_BLOCK=0" run c2 "skinny_tma_kernel|scores_kernel|objective_deferred_kernel" --config 2 --steps 3 --scores
NCU_ENV="PRMF_BLOCK=1" run c2_block "block_kernel" --config 2 --steps 3
NCU_ENV="PRMF_BLOCK=0" run c4 "skinny_tma_gen_kernel|scores_kernel|u_update_tiled|v_update_tiled" --config 4 --steps 2 --scores
NCU_ENV="PRMF_BLOCK=0" run c5 "tc_rowdot_kernel|u_update_tiled|v_update_tiled" --config 5 --steps 2
