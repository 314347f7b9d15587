#!/bin/bash
# Round evidence on one GPU: (1) ncu launch list (duration + DRAM bytes per launch) of one eager step of the default bench,
# (2) config D and C bench lines, (3) the 1-GPU eval sweep.  Outputs in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02f}
BENCH="python bench.py --no-graph --steps 1 --warmup 1 --no-cpu --no-eval --sustained-seconds 0 --instrument-steps 0"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches.log 2>&1; echo "ncu rc=$?"
timeout 300 python bench.py --config D --no-cpu --sustained-seconds 3 > gpurun_out/${TAG}_bench_D.json 2> gpurun_out/${TAG}_bench_D.err; echo "D rc=$?"
timeout 300 python bench.py --config C --no-cpu --sustained-seconds 3 > gpurun_out/${TAG}_bench_C.json 2> gpurun_out/${TAG}_bench_C.err; echo "C rc=$?"
timeout 900 python bench.py --eval-sweep > gpurun_out/${TAG}_eval_sweep_1gpu.jsonl 2> gpurun_out/${TAG}_eval_sweep.err; echo "sweep rc=$?"
tail -n 3 gpurun_out/${TAG}_eval_sweep_1gpu.jsonl | cut -c1-400
du -sh gpurun_out
