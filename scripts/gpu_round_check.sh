#!/bin/bash
# One GPU-box pass of the round: the whole `-m gpu` suite, then the default bench line with the per-shape GEMM table,
# then a torch.profiler kernel table of one eager step.  Logs go to gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
  tail -n 4 gpurun_out/pytest_gpu.log
fi
timeout 600 python bench.py --gemm-table gpurun_out/gemm_table.txt > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 300 python bench.py --profile --steps 3 --warmup 3 --no-cpu --no-eval --sustained-seconds 0 > gpurun_out/bench_profile.json 2> gpurun_out/bench_profile.txt; echo "profile rc=$?"
cat gpurun_out/gemm_table.txt
