#!/bin/bash
# ncu --set full of selected kernels of one eager training step of the default bench config.
# Usage (under gpurun): scripts/ncu_kernels.sh TAG 'regex' [launch-skip] [launch-count]
TAG=$1; RE=$2; SKIP=${3:-60}; CNT=${4:-40}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --no-graph --steps 1 --warmup 1 --no-cpu --no-eval --sustained-seconds 0 --instrument-steps 0"
ncu --set full --clock-control none --import-source on -k "regex:$RE" --launch-skip $SKIP --launch-count $CNT \
    -o $OUT/${TAG}_kernels $BENCH > $OUT/${TAG}_kernels.log 2>&1
ncu -i $OUT/${TAG}_kernels.ncu-rep --page details > $OUT/${TAG}_kernels_details.txt 2>/dev/null
ncu -i $OUT/${TAG}_kernels.ncu-rep --page raw --csv > $OUT/${TAG}_kernels_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_kernels.ncu-rep --page source --csv > $OUT/${TAG}_kernels_source.csv 2>/dev/null
rm -f $OUT/${TAG}_kernels.ncu-rep       # gpurun_out/ travels back only below 64 MiB
gzip -f $OUT/${TAG}_kernels_source.csv
tail -3 $OUT/${TAG}_kernels.log; du -sh $OUT
