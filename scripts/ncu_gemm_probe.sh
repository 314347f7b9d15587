#!/bin/bash
# ncu --set full of single GEMM shapes of config B through scripts/gemm_probe.py (one capture per shape).
# Usage (under gpurun): scripts/ncu_gemm_probe.sh tag "uvqk" "oproj" ...
TAG=$1; shift
mkdir -p gpurun_out
for name in "$@"; do
  f=gpurun_out/${TAG}_gemm_${name// /_}
  ncu --set full --clock-control none --import-source on -k "regex:gemm_tc" --launch-skip 3 --launch-count 1 \
      -o $f python scripts/gemm_probe.py "$name" > $f.log 2>&1
  ncu -i $f.ncu-rep --page details > $f.details.txt 2>/dev/null
  ncu -i $f.ncu-rep --page raw --csv > $f.raw.csv 2>/dev/null
done
python scripts/gemm_probe.py > gpurun_out/${TAG}_gemm_probe.txt 2>&1
cat gpurun_out/${TAG}_gemm_probe.txt
