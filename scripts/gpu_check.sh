#!/bin/bash
# One GPU-box pass: parity tests (fp32 verification path first, then the tcgen05 path), logs to gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
PT="python -m pytest -q --timeout=600 -p no:cacheprovider"
timeout 1200 $PT tests/test_gpu_ops.py -m gpu -k "not bf16" > gpurun_out/ops_fp32.log 2>&1; echo "ops_fp32 rc=$?"
timeout 1200 $PT tests/test_gpu_model.py -m gpu -k "not bf16 and not medium" > gpurun_out/model_fp32.log 2>&1; echo "model_fp32 rc=$?"
timeout 900 $PT tests/test_gpu_ops.py -m gpu -k "bf16" > gpurun_out/ops_bf16.log 2>&1; echo "ops_bf16 rc=$?"
timeout 900 $PT tests/test_gpu_model.py -m gpu -k "bf16 or medium" > gpurun_out/model_bf16.log 2>&1; echo "model_bf16 rc=$?"
tail -n 5 gpurun_out/ops_fp32.log gpurun_out/model_fp32.log gpurun_out/ops_bf16.log gpurun_out/model_bf16.log
