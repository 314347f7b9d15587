#!/bin/bash
# 2-GPU pass: the world-size-2 parity tests, then the default bench on 2 GPUs.  Run under `gpurun --gpus 2`.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_2gpu.json").read().strip().splitlines()[-1])
print("2gpu:", d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("sustained",{}).get("value"), d.get("host_enqueue_ms_per_step"), d.get("eval",{}).get("value"))
PY
