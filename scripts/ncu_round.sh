#!/bin/bash
# Round evidence on the GPU box (one GPU): (1) launch list with per-launch duration and DRAM bytes of one eager training
# step of the default bench config, (2) `--set full` captures of the dominant kernels.  Outputs land in gpurun_out/;
# scripts/ncu_traffic.py turns (1) into profiles/rNN_gemm_traffic.json + a per-kernel share table.
# Usage: scripts/ncu_round.sh [tag]      (run under gpurun; never a bench value: everything here runs under a profiler)
set -u
TAG=${1:-r02}
OUT=gpurun_out
BENCH="python bench.py --no-graph --steps 1 --warmup 1 --no-cpu --no-eval --sustained-seconds 0 --instrument-steps 0"
# the eager step is warm-up (1 step) + timed (1 step) + e2e (1 step): the launch list covers all three, ncu_traffic.py
# keeps the middle one
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv \
    --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k "regex:gemm_tc_grouped_kernel|gemm_tc_kernel|attn_seq_bwd|gate_ln_bwd|nce_combine|layernorm_bwd|scatter_reduce_short|topk_from_candidates" \
    --launch-skip 170 --launch-count 170 -o $OUT/${TAG}_kernels $BENCH > $OUT/${TAG}_kernels.log 2>&1
ncu -i $OUT/${TAG}_kernels.ncu-rep --page details > $OUT/${TAG}_kernels_details.txt 2>/dev/null
ncu -i $OUT/${TAG}_kernels.ncu-rep --page raw --csv > $OUT/${TAG}_kernels_raw.csv 2>/dev/null
echo done
